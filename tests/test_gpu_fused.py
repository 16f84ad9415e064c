"""GPU: the fused on-chip kernels (every class: cube16 / small / cube32 / R1 / cube64 / R8) forced with WC_OPT_PATH=2, against the
oracle.  Covers the geometries the chunking / padding logic branches on, the +M/-M tie slow path,
NaN at f = 0, both input dtypes, and many units per launch (persistent loop, producer run-ahead)."""
import numpy as np
import pytest

from conftest import same_bits, smooth_box

pytestmark = pytest.mark.gpu
F999 = float(np.float32(0.999))

FUSED1 = [(32, 32, 32), (16, 16, 16), (16, 32, 64), (64, 16, 32), (8, 8, 8), (4, 4, 4), (2, 2, 4), (8, 4, 4), (24, 40, 12),
          (48, 16, 16), (32, 16, 64), (64, 64, 8), (4, 64, 128), (12, 20, 28), (2, 2, 4096), (64, 2, 4),
          (16, 16, 24), (8, 8, 40)]
FUSED8 = [(64, 64, 64), (32, 64, 64), (64, 32, 64), (64, 64, 32), (48, 48, 48), (16, 128, 64), (40, 48, 56),
          (40, 40, 40), (36, 36, 36), (56, 56, 40), (48, 40, 64)]      # the last four: clusters of 2 / 4


def check(ctx, oracle, boxes, dims, keep, mode=0):
    packed = ctx.compress_batch(boxes, keep, thresh_mode=mode, dims=dims)
    for b, d, p in zip(boxes, dims, packed):
        runs, vals, _ = oracle.compress_unit(b, d, keep)
        assert p.npairs == runs.size, (d, p.npairs, runs.size)
        assert same_bits(p.runs, runs) and same_bits(p.vals, vals), d
    return packed


@pytest.fixture()
def fused_ctx(ctx):
    ctx.set_path(2)
    yield ctx
    ctx.set_path(0)


@pytest.mark.parametrize("dims", FUSED1 + FUSED8)
@pytest.mark.parametrize("dt", [np.float64, np.float32])
def test_fused_single_unit_shapes(fused_ctx, oracle, dims, dt):
    # float32 rows that are not 16-byte multiples are refused by the y-slab classes and taken by the x-slab ones
    rng = np.random.default_rng(abs(hash((dims, str(dt)))) % 2**32)
    for sym in (False, True):
        b = smooth_box(dims, rng, dtype=dt, sym=sym)
        for keep in (F999, float(np.float32(0.99))):
            check(fused_ctx, oracle, [b], [dims], keep)


def test_fused_many_units_mixed(fused_ctx, oracle):
    rng = np.random.default_rng(2024)
    dims = ([(32, 32, 32)] * 40 + [(64, 64, 64)] * 5 + [(16, 32, 64)] * 7 + [(8, 8, 8)] * 20 + [(24, 40, 12)] * 3) * 2
    boxes = [smooth_box(d, rng, dtype=np.float64 if i % 3 else np.float32, sym=(i % 2 == 0), noise=10.0 ** -(i % 5))
             for i, d in enumerate(dims)]
    for keep in (F999, float(np.float32(0.9999))):
        check(fused_ctx, oracle, boxes, dims, keep)


def test_fused_tie_and_special_values(fused_ctx, oracle):
    cases = []
    for dims in [(4, 4, 4), (32, 32, 32), (64, 64, 64)]:
        n = dims[0] * dims[1] * dims[2]
        rng = np.random.default_rng(n)
        z = np.zeros(n, np.float32); z[5] = 8; z[n - 3] = -8            # +M first
        cases.append((dims, z))
        z = np.zeros(n, np.float32); z[5] = -8; z[n - 3] = 8            # -M first -> everything kept
        cases.append((dims, z))
        z = np.zeros(n, np.float32); z[n // 2 + 1] = -8                 # negative max
        cases.append((dims, z))
        cases.append((dims, np.zeros(n, np.float32)))                   # all zero -> K = 0
        z = rng.standard_normal(n).astype(np.float32); z[:8] = np.nan   # NaN at f = 0 -> nothing kept
        cases.append((dims, z))
        z = rng.standard_normal(n).astype(np.float32); z[n // 3] = np.nan
        cases.append((dims, z))
        z = rng.standard_normal(n).astype(np.float32); z[7] = np.inf; z[n - 9] = -np.inf
        cases.append((dims, z))
        z = np.full(n, 3902.4, np.float32)                              # constant box (the bundled fixtures)
        cases.append((dims, z))
        z = np.full(n, -16.0, np.float32)
        cases.append((dims, z))
    boxes = [c[1].reshape(c[0][2], c[0][1], c[0][0]) for c in cases]
    dims = [c[0] for c in cases]
    check(fused_ctx, oracle, boxes, dims, F999)
    check(fused_ctx, oracle, boxes, dims, 0.5)


def test_fused_global_threshold(fused_ctx, oracle, wc):
    rng = np.random.default_rng(8)
    dims = [(32, 32, 32)] * 6 + [(64, 64, 64)] * 2
    boxes = [smooth_box(d, rng, sym=bool(i % 2)) * (1 + i) for i, d in enumerate(dims)]
    keep = float(np.float32(0.99))
    packed = fused_ctx.compress_batch(boxes, keep, thresh_mode=wc.WC_THRESH_GLOBAL)
    flats = [oracle.haar_forward(b, d) for b, d in zip(boxes, dims)]
    t = oracle.select_threshold_global(flats, keep)
    for f, p in zip(flats, packed):
        runs, vals = oracle.threshold_pack(f, t)
        assert same_bits(p.runs, runs) and same_bits(p.vals, vals)


def test_fused_rejects_what_it_cannot_hold(fused_ctx, wc):
    # odd dims and nz % 4 != 0 are held by the x-slab classes (tests/test_gpu_xslab.py); these are not: too large a plane,
    # too many planes, 128^3 (two passes by y-slabs)
    for shape in [(256, 256, 2), (4, 4, 300), (128, 128, 128)]:
        with pytest.raises(wc.WcError) as e:
            fused_ctx.compress_batch([np.zeros(shape, np.float32)], 0.9)
        assert e.value.status == 2


def test_fused_persistent_loop_many_units_per_cta(fused_ctx, oracle):
    """More units than CTAs / clusters: every CTA runs its persistent loop several times, the producer
    warp prefetches across unit boundaries and the stage ring wraps many times."""
    rng = np.random.default_rng(77)
    dims = [(32, 32, 32)] * 620 + [(64, 64, 64)] * 50
    base32 = [smooth_box((32, 32, 32), rng, dtype=np.float64, sym=bool(i % 2), noise=10.0 ** -(i % 4)) for i in range(8)]
    base64 = [smooth_box((64, 64, 64), rng, dtype=np.float64, sym=bool(i % 2)) for i in range(3)]
    boxes = [base32[i % 8] * (1.0 + 0.001 * i) for i in range(620)] + [base64[i % 3] * (1.0 + 0.01 * i) for i in range(50)]
    packed = fused_ctx.compress_batch(boxes, F999, dims=dims)
    for i in list(range(0, 620, 37)) + list(range(620, 670, 7)) + [619, 669]:
        runs, vals, _ = oracle.compress_unit(boxes[i], dims[i], F999)
        assert same_bits(packed[i].runs, runs) and same_bits(packed[i].vals, vals), i


# ---- fused decompress (k_fused_decompress<1>, <8>) ---------------------------------------------------
@pytest.mark.parametrize("dims", FUSED1 + FUSED8)
@pytest.mark.parametrize("out_dt", [np.float32, np.float64])
def test_fused_decompress_shapes(fused_ctx, oracle, wc, dims, out_dt):
    rng = np.random.default_rng(abs(hash((dims, "dec"))) % 2**32)
    boxes = [smooth_box(dims, rng, sym=s) for s in (False, True)]
    for keep in (F999, float(np.float32(0.99)), 1.0):
        packed = []
        for b in boxes:
            runs, vals, _ = oracle.compress_unit(b, dims, keep)
            packed.append(wc.PackedUnit(dims, dims[0] * dims[1] * dims[2], runs, vals))
        recon = fused_ctx.decompress_batch(packed, out_dtype=out_dt)
        for p, r in zip(packed, recon):
            ob = oracle.decompress_unit(p.runs, p.vals, dims)
            assert r.dtype == out_dt and same_bits(r.astype(np.float32), ob), (dims, keep)


def test_fused_decompress_many_units_and_odd_streams(fused_ctx, oracle, wc):
    rng = np.random.default_rng(4242)
    dims = [(32, 32, 32)] * 400 + [(64, 64, 64)] * 30 + [(16, 32, 64)] * 20
    packed = []
    for i, d in enumerate(dims):
        n = d[0] * d[1] * d[2]
        k = int(rng.integers(0, n // 2))
        if i % 50 == 0:
            k = 0                                   # empty stream -> all zeros
        runs = rng.integers(0, 3, k).astype(np.int32)
        if i % 7 == 0 and k > 10:
            runs[k // 2] = n                        # jumps past the end: this pair and all later ones are dropped
        vals = rng.standard_normal(k).astype(np.float32)
        packed.append(wc.PackedUnit(d, n, runs, vals))
    recon = fused_ctx.decompress_batch(packed)
    for i in list(range(0, 450, 13)) + [0, 399, 400, 429, 449]:
        ob = oracle.decompress_unit(packed[i].runs, packed[i].vals, dims[i])
        assert same_bits(recon[i], ob), i


# ---- plan round trip: the decoder runs from the segment tables the compress kernels wrote ---------------
def test_plan_roundtrip_segment_tables_all_fused_classes(fused_ctx, oracle, wc):
    """compress -> decompress -> rmse inside one plan for every fused class (literal-geometry cubes and
    runtime-geometry shapes, one CTA or 8 slabs per unit), both ingest dtypes, float32 and float64 outputs,
    sparse / dense / all-kept / empty units: the decompress kernels take the per-segment (first pair, last
    kept index) tables written by the compress kernels instead of scanning the pair lists."""
    import torch
    rng = np.random.default_rng(90210)
    shapes = [(32, 32, 32)] * 150 + [(64, 64, 64)] * 20 + [(16, 32, 64)] * 12 + [(24, 40, 12)] * 6 + \
             [(48, 48, 48)] * 6 + [(32, 64, 64)] * 5 + [(8, 8, 8)] * 9 + [(2, 2, 4)] * 3 + [(16, 16, 16)] * 700 + \
             [(8, 16, 8)] * 650 + [(8, 8, 8)] * 5000 + [(40, 40, 40)] * 9 + \
             [(56, 56, 40)] * 5
    host, dts = [], []
    for i, d in enumerate(shapes):
        dt = np.float32 if (i % 4 == 1 and d[0] % 4 == 0) else np.float64
        b = smooth_box(d, rng, dtype=dt, sym=(i % 2 == 0), noise=10.0 ** -(i % 6))
        if i % 17 == 3:
            b = -np.abs(b) - 1.0                       # negative max: every coefficient kept (SURVEY.md D3')
        if i % 19 == 5:
            b = np.zeros_like(b)                       # nothing kept
        host.append(b)
        dts.append(dt)
    dev = [torch.from_numpy(h).cuda() for h in host]
    code = lambda dt: wc.WC_F64 if dt == np.float64 else wc.WC_F32
    descs = wc.capi.box_descs([t.data_ptr() for t in dev], [code(dt) for dt in dts], shapes)
    plan = fused_ctx.plan(descs, wc.WC_DEVICE)
    for out_dt, tdt in ((np.float32, torch.float32), (np.float64, torch.float64)):
        outs = [torch.full(h.shape, 7.0, dtype=tdt, device="cuda") for h in host]
        torch.cuda.synchronize()
        odescs = wc.capi.box_descs([t.data_ptr() for t in outs], [code(out_dt)] * len(outs), shapes)
        for keep in (F999, float(np.float32(0.9)), 1.0):
            plan.compress(keep)
            plan.decompress(odescs, wc.WC_DEVICE)
            rm = plan.rmse(odescs)
            fused_ctx.sync()
            packed = plan.fetch_host()
            for i in list(range(0, 211, 7)) + list(range(211, len(shapes), 53)) + [149, 150, 169, 170, len(shapes) - 1]:
                runs, vals, _ = oracle.compress_unit(host[i], shapes[i], keep)
                assert same_bits(packed[i].runs, runs) and same_bits(packed[i].vals, vals), (i, shapes[i], keep)
                ob = oracle.decompress_unit(runs, vals, shapes[i])
                got = outs[i].cpu().numpy().astype(np.float32)
                assert same_bits(got.reshape(ob.shape), ob), (i, shapes[i], keep, out_dt)
                oe = oracle.rmse(host[i].astype(np.float32), ob, shapes[i])
                assert abs(rm[i] - oe) <= 1e-12 * max(abs(oe), 1e-300), (i, rm[i], oe)
    # global threshold (extension): one key for the whole batch, same tables
    plan.compress(F999, thresh_mode=wc.WC_THRESH_GLOBAL)
    outs = [torch.empty(h.shape, dtype=torch.float32, device="cuda") for h in host]
    odescs = wc.capi.box_descs([t.data_ptr() for t in outs], [wc.WC_F32] * len(outs), shapes)
    plan.decompress(odescs, wc.WC_DEVICE)
    fused_ctx.sync()
    packed = plan.fetch_host()
    for i in (0, 150, 170, 185, 300, 1000, 3000, 6000, len(shapes) - 1):
        ob = oracle.decompress_unit(packed[i].runs, packed[i].vals, shapes[i])
        assert same_bits(outs[i].cpu().numpy().reshape(ob.shape), ob), i
    plan.close()
