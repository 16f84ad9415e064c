"""CPU, world_size 2 over gloo: the host-side multi-GPU logic — unit sharding, the arg-max key all-reduce of
the global-threshold extension (checked against the oracle's concatenation rule), and the stats gather."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def np_key(flat):
    """numpy mirror of make_key + the per-unit max (test infrastructure)."""
    bits = flat.view(np.uint32).astype(np.uint64)
    ab = bits & np.uint64(0x7FFFFFFF)
    valid = ab <= np.uint64(0x7F800000)
    f = np.arange(flat.size, dtype=np.uint64)
    key = (ab << np.uint64(32)) | ((np.uint64(0x7FFFFFFF) - f) << np.uint64(1)) | (bits >> np.uint64(31))
    key = np.where(valid, key, np.uint64(0))
    return int(key.max()) if flat.size else 0


def np_plan_key(flats):
    """mirror of k_global_key: lowest unit index wins ties; bit 63 = first coefficient is NaN."""
    best = 0
    for i, fl in enumerate(flats):
        k = np_key(fl)
        if k == 0:
            continue
        g = (k & 0xFFFFFFFF00000000) | ((0x7FFFFFFF - i) << 1) | (k & 1)
        best = max(best, g)
    first = next((fl for fl in flats if fl.size), None)     # first NON-EMPTY unit (k_global_key)
    if first is not None and np.isnan(first[0]):
        best |= 1 << 63
    return best


def key_to_threshold(key, keep):
    if key >> 63:
        return float("nan")
    a = np.array([(key >> 32) & 0x7FFFFFFF | ((key & 1) << 31)], np.uint32).view(np.float32)[0]
    return float(np.float64(a) * (1.0 - keep))


def _worker(rank, world, port, tmp):
    sys.path.insert(0, ROOT)
    import importlib
    wc = importlib.import_module("wavelet-compression_b200")
    from oracle.pyoracle import Oracle
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    orc = Oracle()
    rng = np.random.default_rng(1234)           # same data on every rank
    dims = (8, 8, 8)
    for case in range(6):
        boxes = [rng.standard_normal(512).astype(np.float32) * (1 + (i % 3)) for i in range(11)]
        if case == 1:
            boxes[7][:] = 0; boxes[7][3] = -50.0       # the winner is negative and sits on rank 1
        if case == 2:
            boxes[2][5] = 40.0; boxes[9][5] = -40.0     # +M on rank 0 and -M on rank 1 tie: rank 0 wins
        if case == 3:
            boxes[0][:8] = np.nan                       # NaN at the very first coefficient
        flats = [orc.haar_forward(b, dims) for b in boxes]
        sizes = [b.size for b in boxes]
        lo, hi = wc.amr_synth.shard_units(sizes, world, rank)
        if case >= 4:
            # rank 0 owns only EMPTY units: the first coefficient of the concatenation is rank 1's; in case 5
            # it is NaN, and a stale NaN flag on rank 0 must not leak into the batch key
            empty = np.zeros(0, np.float32)
            lo, hi = (0, 5) if rank == 0 else (5, 16)
            flats = [empty] * 5 + flats
            if case == 5:
                flats[5] = flats[5].copy(); flats[5][0] = np.nan
        local_key = np_plan_key(flats[lo:hi])
        t = torch.tensor([local_key - (1 << 64) if local_key >> 63 else local_key], dtype=torch.int64)
        g = wc.distributed.allreduce_key(t, lo, has_nonempty=any(f.size for f in flats[lo:hi]))
        gk = int(g.item()) & 0xFFFFFFFFFFFFFFFF
        keep = float(np.float32(0.99))
        want = orc.select_threshold_global(flats, keep)
        got = key_to_threshold(gk, keep)
        assert (np.isnan(want) and np.isnan(got)) or want == got, (case, rank, want, got)
        # single-process key over the whole batch == all-reduced key (up to the unit-index field)
        whole = np_plan_key(flats)
        assert (whole >> 32) == (gk >> 32) and (whole & 1) == (gk & 1), (case, hex(whole), hex(gk))
        # stats gather
        stats = torch.tensor([float(i) for i in range(lo, hi)], dtype=torch.float64)
        allv = wc.distributed.gather_unit_stats(stats)
        if rank == 0:
            assert allv.tolist() == [float(i) for i in range(len(flats))]
    open(os.path.join(tmp, f"ok{rank}"), "w").write("ok")
    dist.destroy_process_group()


def test_two_rank_gloo(tmp_path):
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert (tmp_path / "ok0").exists() and (tmp_path / "ok1").exists()


def test_shard_units_balances_by_size():
    sys.path.insert(0, ROOT)
    import importlib
    wc = importlib.import_module("wavelet-compression_b200")
    sizes = [64 ** 3] * 64 + [32 ** 3] * 1536          # boxes of AMR-256-L4 (x 8 components each)
    for world in (1, 2, 4, 8):
        slices = [wc.amr_synth.shard_units(sizes, world, r) for r in range(world)]
        assert slices[0][0] == 0 and slices[-1][1] == len(sizes)
        assert all(slices[i][1] == slices[i + 1][0] for i in range(world - 1))
        loads = [sum(sizes[a:b]) for a, b in slices]
        assert max(loads) - min(loads) <= 64 ** 3


def _quantile_worker(rank, world, port, q):
    """Radix select across two ranks with all-reduced histograms (gloo): the host model of k_q_hist / k_q_pick."""
    import numpy as np
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        sys.path.insert(0, ROOT)
        import __graft_entry__ as g
        pkg = g.package()
        d = pkg.distributed
        rng = np.random.default_rng(5)
        allv = np.concatenate([rng.standard_normal(5000).astype(np.float32) * 10.0 ** rng.integers(-3, 4, 5000),
                               np.zeros(700, np.float32), np.full(300, 2.5, np.float32), np.full(5, np.nan, np.float32)]).astype(np.float32)
        rng.shuffle(allv)
        mine = np.ascontiguousarray(allv[rank::world])
        keys = mine.view(np.uint32) & np.uint32(0x7fffffff)
        keys = keys[keys <= np.uint32(0x7f800000)]
        out = []
        for keep in (0.0, 0.5, 0.9, 0.999, 1.0):
            n = torch.tensor([mine.size], dtype=torch.int64)
            dist.all_reduce(n)
            ntot = int(n.item())
            rank_ = ntot - min(int(np.floor(keep * ntot)), ntot)
            prefix, thresh = 0, None
            for p, (shift, bits) in enumerate(d.radix_select_passes()):
                h = torch.from_numpy(d.radix_histogram_host(keys, prefix, p))
                d.allreduce_histogram(h)
                b, rank_ = d.radix_pick_host(h.numpy(), rank_)
                if b is None:
                    thresh = -1.0
                    break
                prefix = (prefix << bits) | b
            if thresh is None:
                thresh = float(np.array([prefix], np.uint32).view(np.float32)[0])
            out.append(thresh)
        if rank == 0:
            from oracle.pyoracle import quantile_threshold
            q.put((out, [quantile_threshold([allv], k) for k in (0.0, 0.5, 0.9, 0.999, 1.0)]))
    finally:
        dist.destroy_process_group()


def test_global_quantile_select_across_two_ranks_gloo():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29650 + os.getpid() % 200
    procs = [ctx.Process(target=_quantile_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs: p.start()
    got, want = q.get(timeout=60)
    for p in procs: p.join(60)
    assert got == want, (got, want)
