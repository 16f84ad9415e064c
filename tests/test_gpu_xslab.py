"""GPU: the x-slab classes (wc_xslab.cu, FUSED_CLS_XS1 / 2 / 4 / 8) — boxes of ANY shape on the fused path: odd dimensions
(the trailing element passes through the forward stage, src/compressor.cpp:98-175, and comes back as 0 from the inverse,
src/decompressor.cpp:99-108), nz % 4 != 0, rows that are not 16-byte multiples.  Forced with WC_OPT_PATH = 2 (an error
if a unit would fall back to the generic kernels), everything bit for bit against the oracle: one-shot batch API, plan
round trip (the compress kernel's plane tables), dense-stream decode plans with both index kernels, the batch-wide
threshold, ingest statistics, special values, overflowing and empty streams."""
import numpy as np
import pytest

from conftest import same_bits, smooth_box

pytestmark = pytest.mark.gpu
F999 = float(np.float32(0.999))

# one CTA (XS1): up to ~56000 cells; clusters of 2 / 4 / 8 beyond (xs_slabs in wc_xslab.cu)
XS1 = [(3, 5, 7), (7, 5, 3), (1, 1, 1), (1, 9, 4), (9, 1, 1), (2, 3, 1), (31, 17, 9), (33, 31, 29), (2, 4, 8), (6, 10, 14),
       (17, 2, 2), (255, 3, 5), (256, 5, 3), (5, 64, 64), (30, 30, 34), (15, 47, 44), (33, 33, 33), (6, 100, 90)]
XS2 = [(40, 40, 42), (41, 39, 37)]
XS4 = [(48, 50, 54), (63, 47, 41), (50, 50, 50), (47, 64, 42)]
XS8 = [(63, 63, 63), (65, 61, 57), (62, 66, 62), (130, 34, 50), (64, 64, 62), (61, 64, 64), (130, 66, 34), (127, 65, 33)]
ALL = XS1 + XS2 + XS4 + XS8


@pytest.fixture()
def fused_ctx(ctx):
    ctx.set_path(2)
    yield ctx
    ctx.set_path(0)


def dense_stream(wc, packed):
    k = np.array([p.npairs for p in packed], np.int32)
    pr = np.empty(max(int(k.sum()), 1), wc.capi.PAIR)
    o = 0
    for p in packed:
        pr["run"][o:o + p.npairs] = p.runs
        pr["val"][o:o + p.npairs] = p.vals
        o += p.npairs
    return pr, k


@pytest.mark.parametrize("dt", [np.float64, np.float32])
def test_xslab_compress_decompress_every_class(fused_ctx, oracle, wc, dt):
    rng = np.random.default_rng(77 if dt == np.float64 else 78)
    boxes = [smooth_box(d, rng, dtype=dt, sym=bool(i % 2), noise=10.0 ** -(i % 5)) for i, d in enumerate(ALL)]
    for keep in (F999, float(np.float32(0.9)), 1.0):
        packed = fused_ctx.compress_batch(boxes, keep, dims=ALL)
        for out_dt in (np.float32, np.float64):
            recon = fused_ctx.decompress_batch(packed, out_dtype=out_dt)
            for b, d, p, r in zip(boxes, ALL, packed, recon):
                runs, vals, _ = oracle.compress_unit(b, d, keep)
                assert p.npairs == runs.size, (d, keep, p.npairs, runs.size)
                assert same_bits(p.runs, runs) and same_bits(p.vals, vals), (d, keep)
                ob = oracle.decompress_unit(runs, vals, d)
                assert r.dtype == out_dt and same_bits(r.astype(np.float32).reshape(ob.shape), ob), (d, keep, out_dt)


def test_xslab_special_values(fused_ctx, oracle):
    """+M / -M ties decided by the first occurrence, a negative max (everything kept), NaN at f = 0 (nothing kept), NaNs and
    infinities elsewhere, all-zero boxes — on one-CTA and cluster shapes, with the special cells placed in the first and in
    the last x-slab and in the pass-through planes."""
    cases, dims = [], []
    for d in [(3, 5, 7), (31, 17, 9), (40, 40, 42), (63, 63, 63)]:
        n = d[0] * d[1] * d[2]
        rng = np.random.default_rng(n)
        base = smooth_box(d, rng, sym=True).copy()          # (nz, ny, nx)
        t1 = base.copy(); t1[0, 0, 0] = 900.0; t1[0, 0, 1] = 0.0; t1[-1, -1, -1] = -900.0     # pass-through corner vs block 0
        t2 = base.copy(); t2[-1, -1, -1] = 1e6; t2[-1, -1, -2] = -1e6                        # tie inside the last slab
        t3 = -np.abs(base) - 1.0                                                            # negative max
        t4 = base.copy(); t4[0, 0, 0] = np.nan                                              # NaN at f = 0
        t5 = base.copy(); t5[d[2] // 2, d[1] // 2, d[0] // 2] = np.nan; t5[0, 1, 0] = np.inf
        t6 = np.zeros_like(base)
        t7 = base.copy(); t7[:, :, -1] = 5e4                                                 # the trailing x plane dominates
        t8 = base.copy(); t8[-1, :, :] = -7e4                                                # the trailing z plane dominates
        for t in (t1, t2, t3, t4, t5, t6, t7, t8):
            cases.append(np.ascontiguousarray(t, np.float32)); dims.append(d)
    for keep in (F999, float(np.float32(0.5))):
        packed = fused_ctx.compress_batch(cases, keep, dims=dims)
        recon = fused_ctx.decompress_batch(packed)
        for i, (b, d, p, r) in enumerate(zip(cases, dims, packed, recon)):
            runs, vals, _ = oracle.compress_unit(b, d, keep)
            assert same_bits(p.runs, runs) and same_bits(p.vals, vals), (i, d, keep)
            assert same_bits(r.reshape(-1), oracle.decompress_unit(runs, vals, d).reshape(-1)), (i, d, keep)


@pytest.mark.parametrize("seg_index", [0, 1])
def test_xslab_streams_plan_roundtrip_and_decode_plans(fused_ctx, oracle, wc, seg_index):
    """Many units per launch (persistent loops, both parities of the exchange buffers), the plan round trip (tables from the
    compress kernel), and table-less dense streams through a decode plan with either index kernel; random streams with
    overflowing runs and empty lists."""
    import torch
    rng = np.random.default_rng(31 + seg_index)
    shapes = (XS1[:8] * 40) + (XS2 * 9) + (XS4 * 5) + (XS8 * 4) + [(32, 32, 32)] * 5 + [(64, 64, 64)] * 2
    host, dts = [], []
    for i, d in enumerate(shapes):
        dt = np.float32 if i % 3 == 1 else np.float64
        b = smooth_box(d, rng, dtype=dt, sym=(i % 2 == 0), noise=10.0 ** -(i % 6))
        if i % 17 == 3:
            b = -np.abs(b) - 1.0
        if i % 19 == 5:
            b = np.zeros_like(b)
        host.append(b); dts.append(dt)
    code = lambda dt: wc.WC_F64 if dt == np.float64 else wc.WC_F32
    fused_ctx.set_option(wc.capi.WC_OPT_SEG_INDEX, seg_index)
    try:
        dev = [torch.from_numpy(np.ascontiguousarray(h)).cuda() for h in host]
        descs = wc.capi.box_descs([t.data_ptr() for t in dev], [code(dt) for dt in dts], shapes)
        plan = fused_ctx.plan(descs, wc.WC_DEVICE)
        outs = [torch.full((int(np.prod(d)),), 7.0, dtype=torch.float32, device="cuda") for d in shapes]
        torch.cuda.synchronize()
        odescs = wc.capi.box_descs([t.data_ptr() for t in outs], [wc.WC_F32] * len(outs), shapes)
        want = {}
        for keep in (F999, float(np.float32(0.9))):
            plan.compress(keep)
            plan.decompress(odescs, wc.WC_DEVICE)
            rm = plan.rmse(odescs)
            fused_ctx.sync()
            packed = plan.fetch_host()
            for i in list(range(0, len(shapes), 3)) + [len(shapes) - 1]:
                runs, vals, _ = oracle.compress_unit(host[i], shapes[i], keep)
                assert same_bits(packed[i].runs, runs) and same_bits(packed[i].vals, vals), (i, shapes[i], keep)
                ob = oracle.decompress_unit(runs, vals, shapes[i])
                assert same_bits(outs[i].cpu().numpy().reshape(ob.shape), ob), (i, shapes[i], keep)
                oe = oracle.rmse(host[i].astype(np.float32), ob, shapes[i])
                n = int(np.prod(shapes[i]))
                tol = 1e-12 if n <= (1 << 18) else max(1e-12, n * 2.0 ** -54)      # test_gpu_parity.rmse_tol
                assert abs(rm[i] - oe) <= tol * max(abs(oe), 1e-300), (i, rm[i], oe)
                want[i] = ob
        # the batch-wide threshold (extension): keys-only pass, then packing with the given threshold
        plan.compress(F999, thresh_mode=wc.WC_THRESH_GLOBAL)
        fused_ctx.sync()
        got = plan.fetch_host()
        flats = [oracle.haar_forward(h, d) for h, d in zip(host, shapes)]
        t = oracle.select_threshold_global(flats, F999)
        for i in range(0, len(shapes), 5):
            runs, vals = oracle.threshold_pack(flats[i], t)
            assert same_bits(got[i].runs, runs) and same_bits(got[i].vals, vals), (i, shapes[i])
        plan.compress(float(np.float32(0.9)))
        fused_ctx.sync()
        packed = plan.fetch_host()
        plan.close()
        # the `-d` path: one dense stream + counts, float64 boxes on the host
        pr, kk = dense_stream(wc, packed)
        houts = [np.full((d[2], d[1], d[0]), 7.0, np.float64) for d in shapes]
        od = wc.capi.box_descs([o.ctypes.data for o in houts], [wc.WC_F64] * len(houts), shapes)
        dp = fused_ctx.decode_plan(od, wc.WC_HOST)
        for _ in range(2):
            dp.decode(pr.ctypes.data, kk.ctypes.data, wc.WC_HOST)
            dp.finish()
        dp.close()
        for i, ob in want.items():
            assert same_bits(houts[i].astype(np.float32), ob), (i, shapes[i])
        # hostile / odd streams: runs that jump past the end (that pair and all later ones are dropped), empty lists
        odd = []
        for i, d in enumerate(ALL * 2):
            n = d[0] * d[1] * d[2]
            k = int(rng.integers(0, max(n // 2, 1)))
            if i % 9 == 0:
                k = 0
            runs = rng.integers(0, 3, k).astype(np.int32)
            if i % 4 == 0 and k > 10:
                runs[k // 2] = n
            odd.append(wc.PackedUnit(d, n, runs, rng.standard_normal(k).astype(np.float32)))
        recon = fused_ctx.decompress_batch(odd)
        for p, r in zip(odd, recon):
            assert same_bits(r.reshape(-1), oracle.decompress_unit(p.runs, p.vals, tuple(p.dims)).reshape(-1)), tuple(p.dims)
    finally:
        fused_ctx.set_option(wc.capi.WC_OPT_SEG_INDEX, 0)


def test_xslab_ingest_stats(fused_ctx, oracle, wc):
    rng = np.random.default_rng(5)
    dims = [(3, 5, 7), (31, 17, 9), (40, 40, 42), (63, 47, 41), (63, 63, 63), (7, 5, 3)]
    boxes = []
    for i, d in enumerate(dims):
        b = smooth_box(d, rng, dtype=np.float64 if i % 2 == 0 else np.float32, sym=(i % 3 == 0)).copy()
        if i == 1:
            b *= 300.0                                   # need32
        if i == 2:
            b.reshape(-1)[::7] = np.nan                  # NaNs are skipped by min / max
        if i == 5:
            b[:] = np.nan
        boxes.append(b)
    fused_ctx.set_option(wc.capi.WC_OPT_INGEST_STATS, 1)
    try:
        plan = fused_ctx.plan_host(boxes, dims)
        plan.compress(F999)
        lo, hi, n32 = plan.unit_stats()
        got = plan.fetch_host()
        for i, (b, d) in enumerate(zip(boxes, dims)):
            f = b.astype(np.float32)
            with np.errstate(all="ignore"):
                elo = np.float32(np.inf) if np.isnan(f).all() else np.nanmin(f)
                ehi = np.float32(-np.inf) if np.isnan(f).all() else np.nanmax(f)
            assert lo[i] == elo and hi[i] == ehi, (i, d, lo[i], elo, hi[i], ehi)
            runs, vals, _ = oracle.compress_unit(b, d, F999)
            assert same_bits(got[i].runs, runs) and same_bits(got[i].vals, vals), (i, d)
            want32 = bool(vals.size and (np.abs(vals.astype(np.float64)) > 32767).any())
            assert bool(n32[i]) == want32, (i, d)
        plan.close()
    finally:
        fused_ctx.set_option(wc.capi.WC_OPT_INGEST_STATS, 0)


def test_xslab_unaligned_device_pointers(fused_ctx, oracle, wc):
    """Even-dimension boxes whose device pointers are only element-aligned (a FAB payload behind an odd-length text header):
    the 16-byte vector kernels refuse them, the x-slab kernels take them."""
    import torch
    rng = np.random.default_rng(9)
    dims = [(32, 32, 32), (16, 32, 64), (40, 40, 40)]
    boxes = [smooth_box(d, rng, dtype=np.float64) for d in dims]
    bufs, ptrs = [], []
    for b in boxes:
        t = torch.empty(b.size + 1, dtype=torch.float64, device="cuda")
        t[1:] = torch.from_numpy(b.reshape(-1)).cuda()
        bufs.append(t); ptrs.append(t.data_ptr() + 8)
    outs = [torch.full((b.size + 1,), 7.0, dtype=torch.float32, device="cuda") for b in boxes]
    torch.cuda.synchronize()
    descs = wc.capi.box_descs(ptrs, [wc.WC_F64] * 3, dims)
    odescs = wc.capi.box_descs([t.data_ptr() + 4 for t in outs], [wc.WC_F32] * 3, dims)
    plan = fused_ctx.plan(descs, wc.WC_DEVICE)
    plan.compress(F999)
    plan.decompress(odescs, wc.WC_DEVICE)
    fused_ctx.sync()
    got = plan.fetch_host()
    for i, (b, d) in enumerate(zip(boxes, dims)):
        runs, vals, _ = oracle.compress_unit(b, d, F999)
        assert same_bits(got[i].runs, runs) and same_bits(got[i].vals, vals), d
        ob = oracle.decompress_unit(runs, vals, d)
        assert same_bits(outs[i][1:].cpu().numpy().reshape(ob.shape), ob), d
    plan.close()


def test_xslab_rejects_what_it_cannot_hold(fused_ctx, wc):
    # a plane that no CTA holds, more than 256 planes, a slab above the capacity of 8 CTAs, 128^3 (two passes by y-slabs)
    for shape in [(2, 256, 256), (300, 4, 4), (96, 96, 50), (128, 128, 128)]:
        with pytest.raises(wc.WcError) as e:
            fused_ctx.compress_batch([np.zeros((shape[2], shape[1], shape[0]), np.float32)], 0.9, dims=[shape])
        assert e.value.status == 2
