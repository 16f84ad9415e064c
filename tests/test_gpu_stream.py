"""GPU: the stream-side additions of round 2, all through the C ABI and all against the oracle —
decode plans (wc_dplan_*: one dense pair stream + per-unit counts, the `-d` path of
src/decompressor.cpp:238-255 for a batch), the chunk-parallel segment index (k_seg_index2) against the
one-CTA-per-unit index, stream-ordered wc_plan_set_inputs over a timestep series, the chunk callback of
wc_plan_compress_to_host_chunked, and the per-unit by-products (min / max of src/preprocess.cpp:82-88,
need32 of src/compressor.cpp:224-229)."""
import ctypes as C

import numpy as np
import pytest

from conftest import same_bits, smooth_box

pytestmark = pytest.mark.gpu
F999 = float(np.float32(0.999))

MIXED = ([(64, 64, 64)] * 3 + [(32, 32, 32)] * 9 + [(40, 40, 40)] * 2 + [(36, 36, 36)] * 2 + [(16, 32, 64)] * 3 +
         [(16, 16, 16)] * 4 + [(8, 8, 8)] * 5 + [(8, 4, 4)] * 2 + [(64, 32, 64)] * 2 + [(48, 48, 48)] * 2)


def dense_stream(wc, packed):
    """The units' pairs back to back (= the files' bytes [20, 20+8K) concatenated) + the counts."""
    k = np.array([p.npairs for p in packed], np.int32)
    pr = np.empty(max(int(k.sum()), 1), wc.capi.PAIR)
    o = 0
    for p in packed:
        pr["run"][o:o + p.npairs] = p.runs
        pr["val"][o:o + p.npairs] = p.vals
        o += p.npairs
    return pr, k


def oracle_packed(wc, oracle, boxes, dims, keep):
    out = []
    for b, d in zip(boxes, dims):
        runs, vals, _ = oracle.compress_unit(b, d, keep)
        out.append(wc.PackedUnit(d, d[0] * d[1] * d[2], runs, vals))
    return out


@pytest.mark.parametrize("seg_index", [0, 1])
@pytest.mark.parametrize("out_dt", [np.float32, np.float64])
def test_dplan_host_stream_matches_oracle(wc, ctx, oracle, seg_index, out_dt):
    rng = np.random.default_rng(77 + seg_index)
    dims = MIXED
    boxes = [smooth_box(d, rng, sym=(i % 2 == 0), noise=10.0 ** -(i % 4)) for i, d in enumerate(dims)]
    ctx.set_option(wc.capi.WC_OPT_SEG_INDEX, seg_index)
    try:
        for keep in (F999, float(np.float32(0.9)), 1.0):
            packed = oracle_packed(wc, oracle, boxes, dims, keep)
            pr, k = dense_stream(wc, packed)
            outs = [np.full((d[2], d[1], d[0]), 7.0, out_dt) for d in dims]
            od = wc.capi.box_descs([o.ctypes.data for o in outs], [wc.WC_F64 if out_dt == np.float64 else wc.WC_F32] * len(outs), dims)
            dp = ctx.decode_plan(od, wc.WC_HOST)
            for _ in range(2):                                   # a plan is reusable
                dp.decode(pr.ctypes.data, k.ctypes.data, wc.WC_HOST)
                dp.finish()
            for p, o, d in zip(packed, outs, dims):
                ob = oracle.decompress_unit(p.runs, p.vals, d)
                assert same_bits(o.astype(np.float32), ob), (d, keep, seg_index)
            dp.close()
    finally:
        ctx.set_option(wc.capi.WC_OPT_SEG_INDEX, 0)


def test_dplan_device_stream_from_a_compress_plan(wc, ctx, oracle):
    """compress in one ctx/plan -> dense stream fetched -> decoded by a decode plan of ANOTHER ctx from device
    memory: no segment tables and no cached decode tables travel with the stream."""
    import torch
    rng = np.random.default_rng(5)
    dims = [(64, 64, 64)] * 4 + [(32, 32, 32)] * 24 + [(16, 32, 64)] * 3
    boxes = [smooth_box(d, rng, dtype=np.float64, sym=(i % 3 == 0)) for i, d in enumerate(dims)]
    plan = ctx.plan_host(boxes, dims)
    rec = plan.compress_to_host_records(F999)
    k = rec["npairs"].astype(np.int32).copy()
    total = int(k.sum())
    buf = (C.c_char * (8 * total)).from_address(int(rec[0]["pairs"]))
    pr = np.frombuffer(buf, wc.capi.PAIR, total).copy()        # dense: unit after unit
    plan.close()
    ctx2 = wc.Context(0)
    try:
        d_pairs = torch.from_numpy(pr.view(np.int64).copy()).cuda()
        d_k = torch.from_numpy(k).cuda()
        outs = [torch.full((d[0] * d[1] * d[2],), 3.0, dtype=torch.float32, device="cuda") for d in dims]
        od = wc.capi.box_descs([o.data_ptr() for o in outs], [wc.WC_F32] * len(outs), dims)
        dp = ctx2.decode_plan(od, wc.WC_DEVICE)
        dp.decode(d_pairs.data_ptr(), d_k.data_ptr(), wc.WC_DEVICE)
        dp.finish()
        o = 0
        for i, (b, d) in enumerate(zip(boxes, dims)):
            runs, vals, _ = oracle.compress_unit(b, d, F999)
            assert k[i] == runs.size and same_bits(pr["run"][o:o + k[i]], runs) and same_bits(pr["val"][o:o + k[i]], vals)
            o += int(k[i])
            ob = oracle.decompress_unit(runs, vals, d)
            assert same_bits(outs[i].cpu().numpy().reshape(ob.shape), ob), (i, d)
        dp.close()
    finally:
        ctx2.close()


@pytest.mark.parametrize("pipe", [1, 2, 0])
@pytest.mark.parametrize("seg_index", [0, 1])
def test_dplan_overflow_drop_and_empty_streams(wc, ctx, oracle, seg_index, pipe):
    """rle_decode's `if (idx < N)` (src/decompressor.cpp:24-27): a pair that jumps past the end is dropped together
    with everything after it; empty streams decode to zeros.  Large units so the index kernels see many chunks."""
    rng = np.random.default_rng(99)
    dims = [(64, 64, 64)] * 6 + [(32, 32, 32)] * 6 + [(40, 40, 40)] * 2
    packed = []
    for i, d in enumerate(dims):
        n = d[0] * d[1] * d[2]
        k = [0, n, n // 3, 4097, 4096, 1][i % 6]
        runs = (rng.integers(0, 3, k) if k < n else np.zeros(k)).astype(np.int32)
        if i % 6 == 2:
            runs[k // 2] = n                          # everything from here on is out of the box
        if i % 6 == 3:
            runs[:] = 0; runs[4096] = 2 * n           # the first pair of the second chunk leaves the box
        packed.append(wc.PackedUnit(d, n, runs, rng.standard_normal(k).astype(np.float32)))
    pr, k = dense_stream(wc, packed)
    outs = [np.full((d[2], d[1], d[0]), 7.0, np.float32) for d in dims]
    od = wc.capi.box_descs([o.ctypes.data for o in outs], [wc.WC_F32] * len(outs), dims)
    ctx.set_option(wc.capi.WC_OPT_SEG_INDEX, seg_index)
    ctx.set_option(wc.capi.WC_OPT_DECODE_PIPE, pipe)
    try:
        dp = ctx.decode_plan(od, wc.WC_HOST)
        dp.decode(pr.ctypes.data, k.ctypes.data, wc.WC_HOST)
        dp.finish()
        dp.close()
        recon = ctx.decompress_batch(packed)           # the blocking call takes the same index kernel
    finally:
        ctx.set_option(wc.capi.WC_OPT_SEG_INDEX, 0)
        ctx.set_option(wc.capi.WC_OPT_DECODE_PIPE, 2)
    for i, (p, o, d) in enumerate(zip(packed, outs, dims)):
        ob = oracle.decompress_unit(p.runs, p.vals, d)
        assert same_bits(o, ob), (i, d)
        assert same_bits(recon[i], ob), (i, d)


def test_dplan_rejects_corrupt_streams(wc, ctx):
    dims = [(64, 64, 64), (32, 32, 32)]
    outs = [np.zeros((d[2], d[1], d[0]), np.float32) for d in dims]
    od = wc.capi.box_descs([o.ctypes.data for o in outs], [wc.WC_F32] * 2, dims)
    dp = ctx.decode_plan(od, wc.WC_HOST)
    pr = np.zeros(100, wc.capi.PAIR)
    pr["run"][10] = -5                                              # negative run
    k = np.array([50, 50], np.int32)
    dp.decode(pr.ctypes.data, k.ctypes.data, wc.WC_HOST)
    with pytest.raises(wc.WcError) as e:
        dp.finish()
    assert e.value.status == 7
    k = np.array([50, 32 ** 3 + 1], np.int32)                        # K > ncoef: refused before any copy
    with pytest.raises(wc.WcError) as e:
        dp.decode(pr.ctypes.data, k.ctypes.data, wc.WC_HOST)
    assert e.value.status == 7
    dp.close()
    # the blocking call: K > ncoef is WC_ERR_CORRUPT on the host
    rec = np.zeros(1, wc.capi.PACKED)
    rec[0]["shape"] = (8, 8, 8); rec[0]["ncoef"] = 512; rec[0]["npairs"] = 513; rec[0]["pairs"] = pr.ctypes.data
    o = np.zeros(512, np.float32)
    od1 = wc.capi.box_descs([o.ctypes.data], [wc.WC_F32], [(8, 8, 8)])
    assert ctx.lib.wc_decompress_batch(ctx.h, rec.ctypes.data, 1, wc.WC_HOST, od1.ctypes.data, wc.WC_HOST) == 7


def test_dplan_generic_shapes_fall_back(wc, ctx, oracle):
    rng = np.random.default_rng(3)
    dims = [(5, 7, 3), (32, 32, 32), (31, 17, 9), (64, 64, 64)]
    boxes = [smooth_box(d, rng) for d in dims]
    packed = oracle_packed(wc, oracle, boxes, dims, F999)
    pr, k = dense_stream(wc, packed)
    outs = [np.full((d[2], d[1], d[0]), 7.0, np.float32) for d in dims]
    od = wc.capi.box_descs([o.ctypes.data for o in outs], [wc.WC_F32] * len(outs), dims)
    dp = ctx.decode_plan(od, wc.WC_HOST)
    dp.decode(pr.ctypes.data, k.ctypes.data, wc.WC_HOST)
    dp.finish()
    dp.close()
    for p, o, d in zip(packed, outs, dims):
        assert same_bits(o, oracle.decompress_unit(p.runs, p.vals, d)), d


def test_set_inputs_is_stream_ordered_over_a_series(wc, ctx, oracle):
    """BASELINE config 4 in miniature: one plan, a series of timesteps resident on the device, wc_plan_set_inputs
    between the compress calls with NO synchronisation in between; the slots are read back after each step."""
    import torch
    rng = np.random.default_rng(11)
    dims = [(64, 64, 64)] * 2 + [(32, 32, 32)] * 12 + [(16, 16, 16)] * 3
    T = 5
    series = [[smooth_box(d, rng, dtype=np.float64, sym=(t % 2 == 0), noise=10.0 ** -(t % 3)) for d in dims] for t in range(T)]
    dev = [[torch.from_numpy(b.reshape(-1)).cuda() for b in boxes] for boxes in series]
    descs = [wc.capi.box_descs([t.data_ptr() for t in ts], [wc.WC_F64] * len(dims), dims) for ts in dev]
    plan = ctx.plan(descs[0], wc.WC_DEVICE)
    order = [0, 1, 2, 3, 4, 2, 0, 4, 1, 3, 3, 0]                     # more steps than ring slots, repeats
    for step, t in enumerate(order):
        plan.set_inputs(descs[t])
        plan.compress(F999)
        if step % 3 == 2 or step == len(order) - 1:                  # only now and then: several steps in flight
            got = plan.fetch_host()
            for i, (b, d) in enumerate(zip(series[t], dims)):
                runs, vals, _ = oracle.compress_unit(b, d, F999)
                assert same_bits(got[i].runs, runs) and same_bits(got[i].vals, vals), (step, t, i)
    plan.close()


def test_compress_to_host_chunked_delivers_every_unit_once_in_order(wc, ctx, oracle):
    rng = np.random.default_rng(21)
    dims = [(32, 32, 32)] * 70 + [(64, 64, 64)] * 6 + [(16, 16, 16)] * 9
    boxes = [smooth_box(d, rng, dtype=np.float64 if i % 2 else np.float32, sym=(i % 5 == 0)) for i, d in enumerate(dims)]
    plan = ctx.plan_host(boxes, dims)
    seen, copies = [], {}

    def on_chunk(first, n, recs):
        seen.append((first, n))
        for j in range(n):                                            # the pairs are host-visible NOW
            kk = int(recs[j]["npairs"])
            buf = (C.c_char * (8 * max(kk, 1))).from_address(int(recs[j]["pairs"])) if kk else b""
            copies[first + j] = np.frombuffer(buf, wc.capi.PAIR, kk).copy() if kk else np.zeros(0, wc.capi.PAIR)
    rec = plan.compress_to_host_chunked(F999, on_chunk).copy()
    assert [f for f, _ in seen] == sorted(f for f, _ in seen)
    assert sum(n for _, n in seen) == len(dims) and seen[0][0] == 0
    assert all(seen[i][0] + seen[i][1] == seen[i + 1][0] for i in range(len(seen) - 1))
    rec2 = plan.compress_to_host_records(F999)
    assert np.array_equal(rec["npairs"], rec2["npairs"])
    for i, (b, d) in enumerate(zip(boxes, dims)):
        runs, vals, _ = oracle.compress_unit(b, d, F999)
        assert same_bits(copies[i]["run"], runs) and same_bits(copies[i]["val"], vals), i
    plan.close()


def test_unit_stats_minmax_and_need32(wc, ctx, oracle):
    """min / max of the narrowed values with NaNs skipped (src/preprocess.cpp:82-88) for every kernel class incl. the
    generic path, and need32 = "a kept |value| exceeds INT16_MAX" (src/compressor.cpp:224-229) against the oracle's
    kept values."""
    rng = np.random.default_rng(8)
    dims = [(64, 64, 64), (32, 32, 32), (32, 32, 32), (16, 16, 16), (8, 8, 8), (40, 40, 40), (16, 32, 64), (5, 7, 3),
            (32, 32, 32), (64, 64, 64), (8, 4, 4), (32, 32, 32)]
    boxes = []
    for i, d in enumerate(dims):
        b = smooth_box(d, rng, dtype=np.float64 if i % 2 == 0 else np.float32, sym=(i % 3 == 0)).copy()
        if i in (1, 9):
            b *= 300.0                                   # coefficients well above 32767 -> need32
        if i == 2:
            b.reshape(-1)[::7] = np.nan                  # NaNs are skipped by min / max
        if i == 8:
            b[:] = np.nan                                # nothing comparable: +inf / -inf
        if i == 11:
            b[:] = 40000.0; b.reshape(-1)[5] = -40001.0  # every value above INT16_MAX
        boxes.append(b)
    ctx.set_option(wc.capi.WC_OPT_INGEST_STATS, 1)
    try:
        for keep in (F999, 1.0):
            plan = ctx.plan_host(boxes, dims)
            plan.compress(keep)
            lo, hi, n32 = plan.unit_stats()
            got = plan.fetch_host()
            for i, (b, d) in enumerate(zip(boxes, dims)):
                f = b.astype(np.float32)
                with np.errstate(all="ignore"):
                    elo = np.float32(np.inf) if np.isnan(f).all() else np.nanmin(f)
                    ehi = np.float32(-np.inf) if np.isnan(f).all() else np.nanmax(f)
                assert lo[i] == elo and hi[i] == ehi, (i, d, lo[i], elo, hi[i], ehi)
                runs, vals, _ = oracle.compress_unit(b, d, keep)
                assert same_bits(got[i].runs, runs) and same_bits(got[i].vals, vals), (i, d)
                want32 = bool(vals.size and (np.abs(vals.astype(np.float64)) > 32767).any())
                assert bool(n32[i]) == want32, (i, d, keep, n32[i], want32)
            rec = plan.fetch_records(wc.WC_DEVICE)
            assert [int(r["flags"]) & 1 for r in rec] == [int(x) for x in n32]
            plan.close()
    finally:
        ctx.set_option(wc.capi.WC_OPT_INGEST_STATS, 0)
    # without the option the min / max request is a call-order error, need32 alone is fine
    plan = ctx.plan_host(boxes[:2], dims[:2])
    plan.compress(F999)
    with pytest.raises(wc.WcError):
        plan.unit_stats()
    plan.unit_stats(minmax=False)
    plan.close()


@pytest.mark.parametrize("pipe", [1, 2, 0])
def test_decode_kernels_many_units_all_densities(wc, ctx, oracle, pipe):
    """The cube decode kernels (pipelined and phase-by-phase) over many units per CTA — the pipeline's hand-over of the
    coefficient array between items (clean-as-you-go, full / empty barriers, descriptor ring) only shows with more
    units than SMs — at every density: empty, a single pair, sparse, the bench's ~40 %, all kept (K = N, 1024 chunks)."""
    rng = np.random.default_rng(1234 + pipe)
    dims = [(32, 32, 32)] * 700 + [(64, 64, 64)] * 45
    packed = []
    for i, d in enumerate(dims):
        n = d[0] * d[1] * d[2]
        mode = i % 7
        if mode == 0:
            k = 0
        elif mode == 1:
            k = 1
        elif mode == 2:
            k = n                                        # every coefficient kept: runs are all 0
        else:
            k = int(rng.integers(1, n // 2))
        if k == n:
            runs = np.zeros(k, np.int32)
        else:
            # runs with the right total: K kept positions drawn without replacement
            pos = np.sort(rng.choice(n, size=k, replace=False)) if k else np.zeros(0, np.int64)
            runs = np.diff(np.concatenate([[-1], pos])).astype(np.int32) - 1
        packed.append(wc.PackedUnit(d, n, runs, rng.standard_normal(k).astype(np.float32)))
    pr, kk = dense_stream(wc, packed)
    outs = [np.full((d[2], d[1], d[0]), 7.0, np.float32) for d in dims]
    od = wc.capi.box_descs([o.ctypes.data for o in outs], [wc.WC_F32] * len(outs), dims)
    ctx.set_option(wc.capi.WC_OPT_DECODE_PIPE, pipe)
    try:
        dp = ctx.decode_plan(od, wc.WC_HOST)
        for _ in range(2):
            dp.decode(pr.ctypes.data, kk.ctypes.data, wc.WC_HOST)
            dp.finish()
        dp.close()
    finally:
        ctx.set_option(wc.capi.WC_OPT_DECODE_PIPE, 2)
    for i in list(range(0, len(dims), 9)) + list(range(690, len(dims))):
        ob = oracle.decompress_unit(packed[i].runs, packed[i].vals, dims[i])
        assert same_bits(outs[i], ob), (i, dims[i], packed[i].npairs)


@pytest.mark.parametrize("pipe", [2, 0])
def test_all_kept_lists_with_and_without_a_stray_run(wc, ctx, oracle, pipe):
    """K == ncoef: the staged kernel copies such lists without a scan when every run is 0 (negative max, SURVEY D3');
    a single non-zero (or negative) run must send the unit through the general decode — pairs shifted, the tail dropped
    (src/decompressor.cpp:24-27) — and a negative run is WC_ERR_CORRUPT."""
    rng = np.random.default_rng(77)
    d = (32, 32, 32)
    n = 32768
    packed = []
    for i in range(6):
        runs = np.zeros(n, np.int32)
        if i == 1: runs[20000] = 3
        if i == 2: runs[0] = 1
        if i == 3: runs[n - 1] = 5
        if i == 4: runs[12286] = 7           # first pair past the staging area
        packed.append(wc.PackedUnit(d, n, runs, rng.standard_normal(n).astype(np.float32)))
    pr, k = dense_stream(wc, packed)
    outs = [np.full((32, 32, 32), 7.0, np.float32) for _ in packed]
    od = wc.capi.box_descs([o.ctypes.data for o in outs], [wc.WC_F32] * len(outs), [d] * len(outs))
    ctx.set_option(wc.capi.WC_OPT_DECODE_PIPE, pipe)
    try:
        dp = ctx.decode_plan(od, wc.WC_HOST)
        for _ in range(2):
            dp.decode(pr.ctypes.data, k.ctypes.data, wc.WC_HOST)
            dp.finish()
        for p, o in zip(packed, outs):
            assert same_bits(o, oracle.decompress_unit(p.runs, p.vals, d))
        pr["run"][5 * n + 100] = -2
        dp.decode(pr.ctypes.data, k.ctypes.data, wc.WC_HOST)
        with pytest.raises(wc.WcError) as e:
            dp.finish()
        assert e.value.status == 7
        dp.close()
    finally:
        ctx.set_option(wc.capi.WC_OPT_DECODE_PIPE, 2)


def test_quantile_thresholds_extension(wc, oracle):
    """EXTENSION (the north star's radix select; the reference thresholds at max * (1 - keep), so parity is defined by
    oracle.pyoracle.quantile_threshold): WC_THRESH_QUANTILE keeps the n - floor(keep * n) largest magnitudes of every
    unit (ties at the threshold dropped), WC_THRESH_QUANTILE_GLOBAL the same over the whole batch; pairs = the oracle's
    threshold_pack with that threshold.  Needs a plan created under WC_OPT_PATH = 1."""
    from oracle.pyoracle import quantile_threshold
    rng = np.random.default_rng(2025)
    dims = [(32, 32, 32), (64, 64, 64), (16, 16, 16), (5, 7, 3), (40, 40, 42), (8, 8, 8), (4, 2, 2)]
    boxes = [smooth_box(d, rng, dtype=np.float64 if i % 2 else np.float32, sym=bool(i % 3 == 0), noise=10.0 ** -(i % 3))
             for i, d in enumerate(dims)]
    boxes[5][:] = 3.0                                   # constant box: one non-zero coefficient, the rest ties at 0
    boxes[2].reshape(-1)[7] = np.nan
    ctx = wc.Context(0)
    try:
        plan = ctx.plan_host(boxes, dims)
        with pytest.raises(wc.WcError) as e:             # fused classes keep no coefficient scratch
            plan.compress(0.99, wc.WC_THRESH_QUANTILE)
        assert e.value.status == 8
        plan.close()
        ctx.set_path(1)
        plan = ctx.plan_host(boxes, dims)
        flats = [oracle.haar_forward(oracle.narrow(b) if b.dtype == np.float64 else b, d) for b, d in zip(boxes, dims)]
        for keep in (0.999, 0.9, 0.5, 0.0, 1.0):
            plan.compress(keep, wc.WC_THRESH_QUANTILE)
            got = plan.fetch_host()
            for i, (p, f) in enumerate(zip(got, flats)):
                t = quantile_threshold([f], keep)
                runs, vals = oracle.threshold_pack(f, t)
                assert same_bits(p.runs, runs) and same_bits(p.vals, vals), (keep, i, dims[i], t, p.npairs, runs.size)
                n = f.size
                assert p.npairs <= n - min(int(np.floor(keep * n)), n)
            plan.compress(keep, wc.WC_THRESH_QUANTILE_GLOBAL)
            got = plan.fetch_host()
            t = quantile_threshold(flats, keep)
            for i, (p, f) in enumerate(zip(got, flats)):
                runs, vals = oracle.threshold_pack(f, t)
                assert same_bits(p.runs, runs) and same_bits(p.vals, vals), ("global", keep, i, dims[i], t)
        # the split API a multi-GPU caller uses (histograms exposed between hist and pick) gives the same result
        plan.quantile_begin(0.9, True, 0)
        for ps in range(3):
            assert plan.quantile_hist(ps)
            plan.quantile_pick(ps)
        plan.quantile_pack()
        got = plan.fetch_host()
        t = quantile_threshold(flats, 0.9)
        for p, f in zip(got, flats):
            runs, vals = oracle.threshold_pack(f, t)
            assert same_bits(p.runs, runs) and same_bits(p.vals, vals)
        plan.close()
    finally:
        ctx.close()
