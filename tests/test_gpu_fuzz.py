"""GPU: randomised shapes across every kernel class boundary (<= 512, <= 4096, <= 32768, <= 262144 cells; even
and odd dims -> fused and generic paths), random keep values, both dtypes: packed pairs, reconstructions and
RMSE against the oracle, for the one-shot batch API and for the plan round trip (segment tables)."""
import numpy as np
import pytest

from conftest import same_bits, smooth_box

pytestmark = pytest.mark.gpu


def _shapes(rng, n):
    out = []
    pools = [[2, 4, 6, 8, 10, 12, 16, 20, 24, 32, 40, 48, 64], [1, 3, 5, 7, 9, 15, 17, 31, 33]]
    while len(out) < n:
        odd = rng.random() < 0.25
        d = tuple(int(rng.choice(pools[1] if (odd and rng.random() < 0.5) else pools[0])) for _ in range(3))
        if d[0] * d[1] * d[2] <= 262144:
            out.append(d)
    return out


@pytest.mark.parametrize("seed", [11, 12, 13, 14, 15, 16, 17, 18, 19, 20])
def test_fuzz_batch_and_plan_roundtrip(wc, ctx, oracle, seed):
    import torch
    rng = np.random.default_rng(seed)
    shapes = _shapes(rng, 70)
    # make sure every literal-geometry class and its neighbours are present
    shapes += [(8, 8, 8)] * 3 + [(16, 16, 16)] * 3 + [(32, 32, 32)] * 2 + [(64, 64, 64)] + [(8, 8, 4), (16, 16, 12), (32, 32, 28), (64, 64, 60), (48, 4, 8), (64, 2, 4), (64, 4, 16), (2, 64, 64), (64, 64, 2), (40, 40, 40), (56, 56, 40), (36, 36, 36)]
    boxes, dts = [], []
    for i, d in enumerate(shapes):
        dt = np.float32 if rng.random() < 0.3 else np.float64
        b = smooth_box(d, rng, dtype=dt, sym=bool(i % 2), noise=10.0 ** -int(rng.integers(0, 6)))
        if rng.random() < 0.08:
            b = -np.abs(b) - 0.5                     # negative max: everything kept
        if rng.random() < 0.05:
            b = np.zeros_like(b)
        boxes.append(b); dts.append(dt)
    keep = float(np.float32(rng.choice([0.9, 0.99, 0.999, 0.9999])))
    # one-shot batch API (host boxes in, host pairs out; foreign-stream decode on the way back)
    packed = ctx.compress_batch(boxes, keep, dims=shapes)
    recon = ctx.decompress_batch(packed)
    for b, d, p, r in zip(boxes, shapes, packed, recon):
        runs, vals, _ = oracle.compress_unit(b, d, keep)
        assert same_bits(p.runs, runs) and same_bits(p.vals, vals), (seed, d, b.dtype)
        ob = oracle.decompress_unit(runs, vals, d)
        assert same_bits(r.reshape(ob.shape), ob), (seed, d)
    # plan round trip on the device
    dev = [torch.from_numpy(np.ascontiguousarray(b)).cuda() for b in boxes]
    outs = [torch.full((int(np.prod(d)),), 3.0, dtype=torch.float32, device="cuda") for d in shapes]
    torch.cuda.synchronize()
    code = lambda dt: wc.WC_F64 if dt == np.float64 else wc.WC_F32
    descs = wc.capi.box_descs([t.data_ptr() for t in dev], [code(dt) for dt in dts], shapes)
    odescs = wc.capi.box_descs([t.data_ptr() for t in outs], [wc.WC_F32] * len(outs), shapes)
    plan = ctx.plan(descs, wc.WC_DEVICE)
    for k2 in (keep, float(np.float32(0.95))):
        plan.compress(k2)
        plan.decompress(odescs, wc.WC_DEVICE)
        rm = plan.rmse(odescs)
        ctx.sync()
        got = plan.fetch_host()
        for i, (b, d) in enumerate(zip(boxes, shapes)):
            runs, vals, _ = oracle.compress_unit(b, d, k2)
            assert same_bits(got[i].runs, runs) and same_bits(got[i].vals, vals), (seed, i, d, k2)
            ob = oracle.decompress_unit(runs, vals, d)
            assert same_bits(outs[i].cpu().numpy().reshape(ob.shape), ob), (seed, i, d, k2)
            oe = oracle.rmse(b.astype(np.float32), ob, d)
            n = d[0] * d[1] * d[2]
            tol = 1e-12 if n <= (1 << 18) else max(1e-12, n * 2.0 ** -54)   # strict at every BASELINE size (test_gpu_parity.rmse_tol)
            assert abs(rm[i] - oe) <= tol * max(abs(oe), 1e-300), (seed, i, d, rm[i], oe)
    plan.close()


BIG_SHAPES = [(128, 128, 128), (44, 44, 44), (52, 52, 52), (60, 60, 60), (96, 80, 64), (128, 64, 64), (64, 128, 96),
              (130, 66, 34), (40, 40, 42), (31, 17, 9), (128, 2, 128), (72, 72, 72)]


@pytest.mark.parametrize("seg_index", [0, 1])
def test_big_and_irregular_boxes(wc, ctx, oracle, seg_index):
    """Boxes outside the cluster classes: more than 262144 cells (128^3), half-heights no cluster of 2/4/8 divides (44^3,
    52^3, 60^3) -> decoded by independent y-slab items with a run-time slab count (FUSED_CLS_RBIG: 64 slabs for 128^3,
    11 for 44^3 ...) after the streamed segment index; odd dimensions and nz % 4 != 0 stay on the generic kernels.  Both
    index kernels; one-shot batch API, dense-stream decode plan, plan round trip."""
    rng = np.random.default_rng(4242 + seg_index)
    shapes = BIG_SHAPES
    boxes = []
    for i, d in enumerate(shapes):
        b = smooth_box(d, rng, dtype=np.float64 if i % 3 else np.float32, sym=bool(i % 2), noise=10.0 ** -(i % 4))
        if i == 4:
            b = -np.abs(b) - 0.5                     # negative max: everything kept (K = N)
        boxes.append(b)
    keep = float(np.float32(0.999))
    ctx.set_option(wc.capi.WC_OPT_SEG_INDEX, seg_index)
    try:
        packed = ctx.compress_batch(boxes, keep, dims=shapes)
        recon = ctx.decompress_batch(packed)
        want = []
        for b, d, p, r in zip(boxes, shapes, packed, recon):
            runs, vals, _ = oracle.compress_unit(b, d, keep)
            assert same_bits(p.runs, runs) and same_bits(p.vals, vals), (d, b.dtype)
            ob = oracle.decompress_unit(runs, vals, d)
            assert same_bits(r.reshape(ob.shape), ob), d
            want.append(ob)
        # the `-d` path: dense stream + counts through a decode plan (float64 boxes out)
        kk = np.array([p.npairs for p in packed], np.int32)
        pr = np.empty(max(int(kk.sum()), 1), wc.capi.PAIR)
        o = 0
        for p in packed:
            pr["run"][o:o + p.npairs] = p.runs
            pr["val"][o:o + p.npairs] = p.vals
            o += p.npairs
        outs = [np.full((d[2], d[1], d[0]), 7.0, np.float64) for d in shapes]
        od = wc.capi.box_descs([o.ctypes.data for o in outs], [wc.WC_F64] * len(outs), shapes)
        dp = ctx.decode_plan(od, wc.WC_HOST)
        dp.decode(pr.ctypes.data, kk.ctypes.data, wc.WC_HOST)
        dp.finish()
        dp.close()
        for o, ob, d in zip(outs, want, shapes):
            assert same_bits(o.astype(np.float32), ob), d
    finally:
        ctx.set_option(wc.capi.WC_OPT_SEG_INDEX, 0)


@pytest.mark.parametrize("seg_index", [0, 1])
def test_big_box_runs_longer_than_half_a_million(wc, ctx, oracle, seg_index):
    """A 128^3 unit (2^21 coefficients) whose pair list holds zero runs above 2^19: the streamed index kernel clamps run + 1
    per pair so that its 32-bit sums cannot wrap, and that clamp has to sit above the unit's size (it sat at 2^19, the
    bound of the cluster classes).  Also a run that jumps past the end in the middle of the list."""
    d = (128, 128, 128)
    n = d[0] * d[1] * d[2]
    rng = np.random.default_rng(5 + seg_index)
    units = []
    for gaps in ([0, 600000, 5, 1400000 - 8, 3], [524287, 524288, 524289, 17], [7, 2 ** 20, 2 ** 21, 4, 4], [n - 1], [n], []):
        head = rng.integers(0, 4, 3000).astype(np.int32)                 # an ordinary stretch first, then the long runs
        runs = np.concatenate([head, np.array(gaps, np.int32), rng.integers(0, 3, 500).astype(np.int32)])
        units.append(wc.PackedUnit(d, n, runs, rng.standard_normal(runs.size).astype(np.float32)))
    ctx.set_option(wc.capi.WC_OPT_SEG_INDEX, seg_index)
    try:
        recon = ctx.decompress_batch(units)
        for i, (p, r) in enumerate(zip(units, recon)):
            assert same_bits(r.reshape(-1), oracle.decompress_unit(p.runs, p.vals, d).reshape(-1)), i
    finally:
        ctx.set_option(wc.capi.WC_OPT_SEG_INDEX, 0)
