"""GPU: the CUDA path (through the C ABI) against the oracle and the reference-generated golden
vectors.  Bar: coefficients, pairs, packed bytes and reconstructions bit-exact; RMSE <= 1e-12
relative (SURVEY.md §8a "Parity targets per row")."""
import os
import tempfile

import numpy as np
import pytest

from conftest import same_bits, smooth_box

pytestmark = pytest.mark.gpu

RMSE_RTOL = 1e-12
F999 = float(np.float32(0.999))


def rmse_tol(n):
    """RMSE parity bound (relative).  STRICT 1e-12 (north_star) for every unit of at most 2^18 cells —
    which covers every BASELINE box (64^3 = 2^18, 32^3, 16x32x64, 8x4x2).  The reference sums N squared
    float differences SEQUENTIALLY in float64 (src/calc-loss.cpp:30-35), itself only accurate to
    (N-1)*2^-53 relative on the sum; the GPU uses a fixed pairwise tree (~log2(N)*2^-53).  The two may
    differ by N*2^-53 on the sum = N*2^-54 on the root.  N*2^-54 reaches 1e-12 at N = 18 014 cells already,
    so it is NOT used as a bound for the BASELINE sizes: only beyond 2^18 cells (96^3 was observed at
    1.1e-12) does the scaled bound N*2^-54 apply (1.46e-11 * N/2^18)."""
    return RMSE_RTOL if n <= (1 << 18) else max(RMSE_RTOL, n * 2.0 ** -54)


def rmse_close(a, b, n=0):
    if np.isnan(a) or np.isnan(b):
        return np.isnan(a) and np.isnan(b)
    if np.isinf(a) or np.isinf(b):
        return a == b
    return abs(a - b) <= rmse_tol(n) * max(abs(b), np.finfo(np.float64).tiny)


def test_library_is_the_cuda_one(wc, ctx):
    import ctypes
    assert os.path.exists(wc.capi.LIB_PATH)
    n = ctypes.c_int(0)
    assert wc.capi.load().wc_device_count(ctypes.byref(n)) == 0 and n.value >= 1


@pytest.mark.parametrize("path", [1, 0])
def test_golden_cases_through_batch_api(wc, ctx, golden, path):
    """Every golden case (reference doctest vectors, plotfile boxes, edge cases) in ONE batch call."""
    ctx.set_path(path)
    try:
        boxes = [golden.arrays(i)["in"] for i in range(len(golden.cases))]
        dims = [tuple(c["dims"]) for c in golden.cases]
        by_keep = {}
        for i, c in enumerate(golden.cases):
            by_keep.setdefault(c["keep"], []).append(i)
        for keep, idx in by_keep.items():
            packed = ctx.compress_batch([boxes[i] for i in idx], keep, dims=[dims[i] for i in idx])
            recon = ctx.decompress_batch(packed)
            b32 = [boxes[i].astype(np.float32) for i in idx]
            rm = ctx.rmse_batch(b32, recon)
            for n, i in enumerate(idx):
                a, name = golden.arrays(i), golden.cases[i]["name"]
                assert same_bits(packed[n].runs, a["runs"]), name
                assert same_bits(packed[n].vals, a["vals"]), name
                assert packed[n].serialize() == a["ser"].tobytes(), name
                assert same_bits(recon[n], a["recon"]), name
                assert rmse_close(rm[n], a["rmse"][0], boxes[i].size), (name, rm[n], a["rmse"][0])
    finally:
        ctx.set_path(0)


def test_golden_cases_through_primitives(wc, ctx, golden):
    for i, c in enumerate(golden.cases):
        a = golden.arrays(i)
        dims = tuple(c["dims"])
        box = a["in"].reshape(dims[2], dims[1], dims[0])
        coef = ctx.haar_forward(box)
        assert same_bits(coef, a["coef"]), c["name"]                       # F (+A1)
        runs, vals = ctx.threshold_pack(a["coef"], c["keep"])              # T, M, P
        assert same_bits(runs, a["runs"]) and same_bits(vals, a["vals"]), c["name"]
        flat = ctx.rle_decode(a["runs"], a["vals"], a["coef"].size)        # U
        assert same_bits(ctx.haar_inverse(flat, dims), a["recon"]), c["name"]  # I


@pytest.mark.parametrize("dims", [(64, 64, 64), (32, 32, 32), (16, 32, 64), (128, 32, 16), (24, 40, 12),
                                  (66, 34, 18), (31, 17, 9), (2, 2, 2048), (2048, 2, 2), (1, 1, 1), (8, 4, 2),
                                  (96, 96, 96)])
@pytest.mark.parametrize("in_dtype", [np.float32, np.float64])
def test_random_boxes_vs_oracle(wc, ctx, oracle, dims, in_dtype):
    rng = np.random.default_rng(abs(hash((dims, str(in_dtype)))) % 2**32)
    boxes = [smooth_box(dims, rng, dtype=in_dtype), smooth_box(dims, rng, dtype=in_dtype, sym=True),
             (rng.standard_normal(dims[::-1]) * 3).astype(in_dtype)]
    for keep in (F999, float(np.float32(0.99)), float(np.float32(0.9999))):
        packed = ctx.compress_batch(boxes, keep)
        recon = ctx.decompress_batch(packed)
        rm = ctx.rmse_batch([b.astype(np.float32) for b in boxes], recon)
        for b, p, r, e in zip(boxes, packed, recon, rm):
            runs, vals, _ = oracle.compress_unit(b, dims, keep)
            assert p.npairs == runs.size
            assert same_bits(p.runs, runs) and same_bits(p.vals, vals)
            assert p.serialize() == oracle.packed_bytes(b, dims, keep).tobytes()
            ob = oracle.decompress_unit(runs, vals, dims)
            assert same_bits(r, ob)
            assert rmse_close(e, oracle.rmse(b.astype(np.float32), ob, dims), b.size)


def test_mixed_batch_many_units(wc, ctx, oracle):
    """A ragged batch (different dims and dtypes, empty box included) in one call, unit order kept."""
    rng = np.random.default_rng(99)
    shapes = [(32, 32, 32)] * 5 + [(64, 64, 64)] * 2 + [(8, 4, 2), (3, 5, 7), (0, 4, 4), (16, 32, 64), (10, 6, 2)] * 3
    boxes, dims = [], []
    for n, d in enumerate(shapes):
        dt = np.float64 if n % 2 else np.float32
        boxes.append(smooth_box(d, rng, dtype=dt, sym=(n % 3 == 0)) if d[0] else np.zeros((d[2], d[1], 0), dt))
        dims.append(d)
    packed = ctx.compress_batch(boxes, F999, dims=dims)
    assert [p.dims for p in packed] == dims
    for b, d, p in zip(boxes, dims, packed):
        if d[0] == 0:
            assert p.npairs == 0 and p.ncoef == 0
            continue
        runs, vals, _ = oracle.compress_unit(b, d, F999)
        assert same_bits(p.runs, runs) and same_bits(p.vals, vals), d
    recon = ctx.decompress_batch(packed, out_dtype=np.float64)
    for b, d, p, r in zip(boxes, dims, packed, recon):
        if d[0]:
            ob = oracle.decompress_unit(p.runs, p.vals, d)
            assert r.dtype == np.float64 and same_bits(r.astype(np.float32), ob) and np.array_equal(r, ob.astype(np.float64))


def test_decode_drops_out_of_range_pairs_like_the_reference(wc, ctx, oracle):
    runs = np.array([1, 5, 0, 2], np.int32)
    vals = np.array([1, 2, 3, 4], np.float32)
    assert same_bits(ctx.rle_decode(runs, vals, 4), oracle.rle_decode(runs, vals, 4))
    assert same_bits(ctx.rle_decode(runs, vals, 64), oracle.rle_decode(runs, vals, 64))
    big_runs = np.zeros(5000, np.int32); big_runs[::7] = 3
    big_vals = np.arange(5000, dtype=np.float32)
    assert same_bits(ctx.rle_decode(big_runs, big_vals, 6000), oracle.rle_decode(big_runs, big_vals, 6000))


def test_reference_style_file_roundtrip(wc, ctx):
    """The reference's 'File writing/compression' doctest (src/compressor.cpp:387-406) through the
    Python mirror of its interface, plus 'Calc RMSE' (src/calc-loss.cpp:68-86)."""
    box = np.full((16, 8, 4), 5.0, np.float32)  # Box3D box(4, 8, 16, 5.0f)
    with tempfile.TemporaryDirectory() as d:
        wc.compress([box], [0], 0.999, 0, 0, 0, d, ctx=ctx)
        result = wc.decompress(os.path.join(d, "compressed-wavelet-0-0-0-0.xz"), 0, 0, 0, 0, ctx=ctx)
    assert same_bits(result, box)
    t1 = [np.zeros((2, 2, 2), np.float32)] * 2
    t2 = [np.full((2, 2, 2), 3.5, np.float32)] * 2
    assert wc.calc_rmse_per_box(t1, t2, 2, ctx=ctx) == [3.5, 3.5]


def test_xz_files_are_byte_identical_to_the_reference(wc, ctx, ref):
    """Same liblzma preset => the .xz the host writes equals the file the reference writes."""
    rng = np.random.default_rng(1)
    box = smooth_box((16, 16, 16), rng)
    with tempfile.TemporaryDirectory() as d1, tempfile.TemporaryDirectory() as d2:
        wc.compress([box], [3], F999, 2, 1, 7, d1, ctx=ctx)
        ref.compress(box.reshape(1, -1), (16, 16, 16), F999, d2, t=2, lev=1, box_idx=7, comp_ids=[3])
        name = "compressed-wavelet-2-1-3-7.xz"
        assert open(os.path.join(d1, name), "rb").read() == open(os.path.join(d2, name), "rb").read()
        back, dims = ref.decompress(os.path.join(d1, name))   # the reference can read our file
        mine = wc.decompress(os.path.join(d2, name), ctx=ctx)  # and we can read the reference's
        assert same_bits(back, mine)


def test_global_threshold_extension(wc, ctx, oracle):
    rng = np.random.default_rng(5)
    dims = (16, 16, 16)
    boxes = [smooth_box(dims, rng, sym=bool(i % 2)) * (1 + i) for i in range(6)]
    keep = float(np.float32(0.99))
    packed = ctx.compress_batch(boxes, keep, thresh_mode=wc.WC_THRESH_GLOBAL)
    flats = [oracle.haar_forward(b, dims) for b in boxes]
    t = oracle.select_threshold_global(flats, keep)
    for f, p in zip(flats, packed):
        runs, vals = oracle.threshold_pack(f, t)
        assert same_bits(p.runs, runs) and same_bits(p.vals, vals)


def test_plan_device_roundtrip_with_torch_buffers(wc, ctx, oracle):
    """Device-resident path: inputs are torch CUDA tensors, compress -> decompress -> RMSE never
    leave the GPU (what estimate mode needs, src/modes.cpp:236-291)."""
    import torch
    rng = np.random.default_rng(17)
    dims_list = [(32, 32, 32)] * 4 + [(64, 64, 64)] + [(16, 32, 64)] * 2 + [(6, 10, 14)]
    host = [smooth_box(d, rng, dtype=np.float64, sym=(i % 2 == 1)) for i, d in enumerate(dims_list)]
    dev = [torch.from_numpy(h).cuda() for h in host]
    outs = [torch.empty(h.shape, dtype=torch.float32, device="cuda") for h in host]
    torch.cuda.synchronize()
    descs = wc.capi.box_descs([t.data_ptr() for t in dev], [wc.WC_F64] * len(dev), dims_list)
    odescs = wc.capi.box_descs([t.data_ptr() for t in outs], [wc.WC_F32] * len(dev), dims_list)
    plan = ctx.plan(descs, wc.WC_DEVICE)
    for keep in (F999, float(np.float32(0.9999))):
        plan.compress(keep)
        plan.decompress(odescs, wc.WC_DEVICE)
        rm = plan.rmse(odescs)
        packed = plan.fetch_host()
        assert plan.total_pairs() == sum(p.npairs for p in packed)
        for h, d, p, o, e in zip(host, dims_list, packed, outs, rm):
            runs, vals, _ = oracle.compress_unit(h, d, keep)
            assert same_bits(p.runs, runs) and same_bits(p.vals, vals)
            ob = oracle.decompress_unit(runs, vals, d)
            assert same_bits(o.cpu().numpy(), ob)
            assert rmse_close(e, oracle.rmse(h.astype(np.float32), ob, d))
    plan.close()


def test_error_codes(wc, ctx):
    descs = wc.capi.box_descs([0], [0], [(-1, 2, 2)])
    out = np.zeros(1, wc.capi.PACKED)
    assert ctx.lib.wc_compress_batch(ctx.h, descs.ctypes.data, 1, 0, 0.9, 0, out.ctypes.data, 0) == 2
    descs = wc.capi.box_descs([0], [7], [(2, 2, 2)])
    assert ctx.lib.wc_compress_batch(ctx.h, descs.ctypes.data, 1, 0, 0.9, 0, out.ctypes.data, 0) == 1
    bad = wc.PackedUnit((4, 4, 4), 63, np.zeros(0, np.int32), np.zeros(0, np.float32))
    with pytest.raises(wc.WcError) as e:
        ctx.decompress_batch([bad])
    assert e.value.status == 7
    neg = wc.PackedUnit((4, 4, 4), 64, np.array([-2], np.int32), np.array([1.0], np.float32))
    with pytest.raises(wc.WcError) as e:
        ctx.decompress_batch([neg])
    assert e.value.status == 7


def test_cpp_dropin_runs_the_reference_doctests_on_the_gpu():
    """oracle/_ref/dropin_test = wavelet-compression_b200/host/dropin_test.cpp compiled against the
    REFERENCE's own grid.h / box-structs.h (types + signatures of src/compressor.h, decompressor.h,
    calc-loss.h), linked to libwcgpu.so: the reference's doctest cases + a batched run."""
    import subprocess
    exe = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_ref", "dropin_test")
    if not os.path.exists(exe):
        pytest.skip("oracle/_ref/dropin_test not built (needs /root/reference headers at build time)")
    r = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and "dropin ok" in r.stdout, r.stdout + r.stderr


def test_multi_gpu_sharding_and_nccl_global_threshold():
    """One rank per GPU under torchrun (needs >= 2 GPUs; the 1-GPU box skips): tests/mgpu_check.py."""
    import subprocess
    import sys
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs at least 2 GPUs")
    n = 2 if n < 4 else 4
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}",
                        "--master-addr", "127.0.0.1", "--master-port", "29533",
                        os.path.join(root, "tests", "mgpu_check.py")], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "mgpu_check ok" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]
