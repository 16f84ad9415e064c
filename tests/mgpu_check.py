"""Launched by torchrun (one rank per GPU): sharded per-unit compression + the NCCL global-threshold extension (MAX of the
arg-max key) + the NCCL global-quantile extension (SUM of the radix-select histograms), checked against the oracle over
the WHOLE batch.  Exit code 0 = parity."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    import importlib
    wc = importlib.import_module("wavelet-compression_b200")
    from conftest import same_bits, smooth_box
    from oracle.pyoracle import Oracle
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    rng = np.random.default_rng(99)                       # identical batch on every rank
    dims = [(32, 32, 32)] * 24 + [(64, 64, 64)] * 3 + [(16, 32, 64)] * 5 + [(6, 10, 14)] * 4
    boxes = [smooth_box(d, rng, dtype=np.float64, sym=bool(i % 2)) * (1.0 + 0.37 * (i % 5)) for i, d in enumerate(dims)]
    sizes = [b.size for b in boxes]
    lo, hi = wc.amr_synth.shard_units(sizes, world, rank)
    mine = [torch.from_numpy(b).to(dev) for b in boxes[lo:hi]]
    descs = wc.capi.box_descs([t.data_ptr() for t in mine], [wc.WC_F64] * len(mine), dims[lo:hi])
    stream = torch.cuda.Stream(device=dev)
    ctx = wc.Context(local, stream=stream.cuda_stream)
    keep = float(np.float32(0.99))
    ok = True
    with torch.cuda.stream(stream):
        plan = ctx.plan(descs, wc.WC_DEVICE)
        # 1. reference semantics: per-unit thresholds, no collective on the data path
        plan.compress(keep)
        packed = plan.fetch_host()
        counts = wc.distributed.gather_unit_stats(torch.tensor([p.npairs for p in packed], dtype=torch.int64, device=dev))
        # 2. extension: one threshold for the whole batch (NCCL MAX all-reduce of the arg-max key)
        wc.distributed.compress_global_threshold(plan, keep, lo, dev)
        packed_g = plan.fetch_host()
        # 3. extension: one QUANTILE threshold for the whole batch (NCCL SUM all-reduce of the radix-select histograms);
        #    needs a coefficient scratch for every unit, i.e. a plan created under WC_OPT_PATH = 1
        ctx.set_path(1)
        plan_q = ctx.plan(descs, wc.WC_DEVICE)
        wc.distributed.compress_global_quantile(plan_q, keep, sum(sizes[lo:hi]), dev)
        packed_q = plan_q.fetch_host()
        plan_q.close()
        ctx.set_path(0)
    orc = Oracle()
    flats = [orc.haar_forward(orc.narrow(b), d) for b, d in zip(boxes, dims)]
    tg = orc.select_threshold_global(flats, keep)
    from oracle.pyoracle import quantile_threshold
    tq = quantile_threshold(flats, keep)
    for i, pq in enumerate(packed_q):
        rq, vq = orc.threshold_pack(flats[lo + i], tq)
        ok &= same_bits(pq.runs, rq) and same_bits(pq.vals, vq)
    for i, (p, pg) in enumerate(zip(packed, packed_g)):
        u = lo + i
        runs, vals, _ = orc.compress_unit(boxes[u], dims[u], keep)
        ok &= same_bits(p.runs, runs) and same_bits(p.vals, vals)
        rg, vg = orc.threshold_pack(flats[u], tg)
        ok &= same_bits(pg.runs, rg) and same_bits(pg.vals, vg)
    if rank == 0:
        want = [orc.compress_unit(b, d, keep)[0].size for b, d in zip(boxes, dims)]
        ok &= counts.cpu().tolist() == want
    flag = torch.tensor([1 if ok else 0], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("mgpu_check", "ok" if flag.item() else "FAILED", "world", world)
    plan.close()
    ctx.close()
    dist.destroy_process_group()
    return 0 if flag.item() else 1


if __name__ == "__main__":
    sys.exit(main())
