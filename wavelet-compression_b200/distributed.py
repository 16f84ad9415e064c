"""Multi-GPU host logic: one process per GPU (torch.distributed, NCCL over NVLink on the GPU box, gloo
in the CPU tests).

The reference's units (timestep, level, box, component) are independent: the data
path needs NO collective — ranks take contiguous, size-balanced slices of the unit list
(amr_synth.shard_units) and only per-unit statistics are gathered.  The one exchange step is the
EXTENSION of BASELINE config 5, a threshold shared by all boxes: every rank reduces its units to
one 64-bit arg-max key on the device (wc_plan_transform), the keys are all-reduced (MAX) and every
rank packs with the winner (wc_plan_pack_with_key), so the kept set is identical to a
single-process run over the concatenated batch.

Key layout (csrc/wc_common.cuh make_key, csrc/wc_generic.cu k_global_key):
    bit 63      the very first coefficient of the batch is NaN (meaningful on the lowest rank that owns a
                non-empty unit: empty units contribute no coefficient to the concatenation)
    bits 62:32  bits of |c_max|
    bits 31:1   0x7fffffff - unit index   (lower unit wins ties, as std::max_element would)
    bit 0       sign of c_max
"""
from __future__ import annotations

import torch
import torch.distributed as dist

_LOW31 = 0x7FFFFFFF


def rebase_key(key: torch.Tensor, unit_offset: int):
    """Local plan key (unit index local to this rank) -> key ordered by GLOBAL unit index.
    `key` is a 1-element int64 tensor holding the raw 64-bit pattern.  Returns (key without bit 63, bit 63)."""
    flag = (key < 0).to(torch.int64)                       # bit 63
    k = key & 0x7FFFFFFFFFFFFFFF
    has = (k != 0).to(torch.int64)
    hi = (k >> 32) << 32
    sign = k & 1
    local = _LOW31 - ((k >> 1) & _LOW31)
    rebased = hi | ((_LOW31 - (local + unit_offset)) << 1) | sign
    return rebased * has, flag


def allreduce_key(key: torch.Tensor, unit_offset: int, group=None, has_nonempty: bool = True) -> torch.Tensor:
    """Batch-wide arg-max key over all ranks: MAX of the rebased keys, plus the NaN-at-f=0 bit of the LOWEST
    rank that owns a non-empty unit (`has_nonempty`: word 1 of wc_plan_transform's key buffer) — a rank
    holding only empty units owns no coefficient of the concatenation, so its bit must not count.
    Returns a 1-element int64 tensor with the 64-bit pattern wc_plan_pack_with_key expects."""
    rank = dist.get_rank(group)
    world = dist.get_world_size(group)
    k, flag = rebase_key(key, unit_offset)
    owner = torch.full_like(k, rank if has_nonempty else world)
    dist.all_reduce(owner, op=dist.ReduceOp.MIN, group=group)
    flag0 = flag if int(owner.item()) == rank else torch.zeros_like(flag)
    dist.all_reduce(k, op=dist.ReduceOp.MAX, group=group)
    dist.all_reduce(flag0, op=dist.ReduceOp.MAX, group=group)
    # bit 63 as a sign bit of the int64 pattern
    return torch.where(flag0 > 0, k - (1 << 62) - (1 << 62), k)


def compress_global_threshold(plan, keep: float, unit_offset: int, device, group=None):
    """WC_THRESH_GLOBAL across ranks for an already created device-resident plan."""
    ctx = plan.ctx
    key_dev = plan.transform()
    t2 = torch.empty(2, dtype=torch.int64, device=device)    # [key, batch owns a non-empty unit]
    from .capi import check
    check(ctx.lib.wc_memcpy(ctx.h, t2.data_ptr(), key_dev, 16, 2), "wc_memcpy", ctx.h)
    g = allreduce_key(t2[:1].clone(), unit_offset, group, has_nonempty=bool(int(t2[1].item()))).contiguous()
    if g.is_cuda:
        torch.cuda.current_stream(device).synchronize()
    plan.pack_with_key(keep, g.data_ptr())
    plan._global_key_hold = g
    return g


Q_BINS = 2048


def allreduce_histogram(hist: torch.Tensor, group=None) -> torch.Tensor:
    """Sum of the ranks' radix-select histograms (int64[2048], exact: counts stay far below 2^63)."""
    dist.all_reduce(hist, op=dist.ReduceOp.SUM, group=group)
    return hist


def radix_select_passes():
    """(shift, bits) of the three radix-select passes over the 31 magnitude bits (as k_q_hist)."""
    return ((20, 11), (9, 11), (0, 9))


def radix_histogram_host(keys, prefix: int, pass_: int):
    """Host model of k_q_hist for one shard: histogram (int64[2048]) of pass `pass_` over the uint32 magnitude keys
    (NaNs removed) that match `prefix`.  Used by the CPU tests of the multi-rank select; the product path is the kernel."""
    import numpy as np
    shift, bits = radix_select_passes()[pass_]
    k = np.asarray(keys, np.uint32)
    if pass_ > 0:
        k = k[(k >> np.uint32(shift + bits)) == np.uint32(prefix)]
    return np.bincount(((k >> np.uint32(shift)) & np.uint32((1 << bits) - 1)).astype(np.int64), minlength=Q_BINS).astype(np.int64)


def radix_pick_host(hist, rank: int):
    """Host model of k_q_pick: bucket holding `rank` counted from the top, and the rank inside it (None: rank >= total)."""
    total = int(hist.sum())
    if rank >= total:
        return None, rank
    c = 0
    for b in range(len(hist) - 1, -1, -1):
        if c + int(hist[b]) > rank:
            return b, rank - c
        c += int(hist[b])
    raise AssertionError("unreachable")


def compress_global_quantile(plan, keep: float, n_local: int, device, group=None):
    """WC_THRESH_QUANTILE_GLOBAL across ranks (EXTENSION): every rank histograms its own coefficients, the histograms
    are summed with NCCL, every rank picks the same bucket — three passes — and packs with the common threshold, so
    the kept set is the one a single GPU holding all units would keep."""
    ctx = plan.ctx
    from .capi import check
    n = torch.tensor([int(n_local)], dtype=torch.int64, device=device)
    dist.all_reduce(n, op=dist.ReduceOp.SUM, group=group)
    plan.quantile_begin(keep, True, int(n.item()))
    h = torch.empty(Q_BINS, dtype=torch.int64, device=device)
    for p in range(3):
        hist_dev = plan.quantile_hist(p)
        check(ctx.lib.wc_memcpy(ctx.h, h.data_ptr(), hist_dev, 8 * Q_BINS, 2), "wc_memcpy", ctx.h)
        ctx.sync()
        allreduce_histogram(h, group)
        if h.is_cuda:
            torch.cuda.current_stream(device).synchronize()
        check(ctx.lib.wc_memcpy(ctx.h, hist_dev, h.data_ptr(), 8 * Q_BINS, 2), "wc_memcpy", ctx.h)
        plan.quantile_pick(p)
    plan.quantile_pack()


def gather_unit_stats(local: torch.Tensor, group=None):
    """Per-unit statistics (pair counts, RMSE) of every rank on rank 0, in unit order.  Ranks may
    hold different numbers of units."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    n = torch.tensor([local.numel()], dtype=torch.int64, device=local.device)
    sizes = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(sizes, n, group=group)
    sizes = [int(s.item()) for s in sizes]
    m = max(sizes) if sizes else 0
    pad = torch.zeros(m, dtype=local.dtype, device=local.device)
    pad[:local.numel()] = local.reshape(-1)
    outs = [torch.zeros_like(pad) for _ in range(world)]
    dist.all_gather(outs, pad, group=group)
    if rank != 0:
        return None
    return torch.cat([o[:s] for o, s in zip(outs, sizes)])
