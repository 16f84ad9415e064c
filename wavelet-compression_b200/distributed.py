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
