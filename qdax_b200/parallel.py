"""Collective plumbing for DistributedMAPElites: one process per GPU, torch.distributed (NCCL over
NVLink / NVSwitch on the GPU box; gloo in the CPU tests).  Device-agnostic on purpose: these helpers only move
and merge buffers -- all arithmetic on offspring happens in libqdx.so.

The reference gathers (genotypes, fitnesses, descriptors) with jax.lax.all_gather and concatenates along axis 0 in
device order (qdax/core/distributed_map_elites.py:134-141), so the global index of offspring i of rank r is
r * B_dev + i.  `all_gather_rows` reproduces exactly that layout."""

from __future__ import annotations

from typing import Optional

import torch
import torch.distributed as dist

_SIGN64 = -(1 << 63)


def world(group=None):
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(group), dist.get_world_size(group)
    return 0, 1


def all_gather_rows(x: torch.Tensor, group=None, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """concatenate(all_gather(x), axis=0): rank r's rows land at [r*B_dev, (r+1)*B_dev)."""
    rank, size = world(group)
    if size == 1:
        return x
    x = x.contiguous()
    if out is None:
        out = torch.empty((size * x.shape[0],) + tuple(x.shape[1:]), dtype=x.dtype, device=x.device)
    dist.all_gather_into_tensor(out, x, group=group)
    return out


def all_reduce_max_u64_(keys_i64: torch.Tensor, group=None) -> torch.Tensor:
    """In-place element-wise UNSIGNED 64-bit max across ranks of a buffer viewed as int64 (the packed insertion
    keys).  NCCL / gloo reduce signed integers, so the sign bit is flipped around the collective: x ^ 2^63 maps
    unsigned order onto signed order."""
    _, size = world(group)
    if size == 1:
        return keys_i64
    keys_i64.bitwise_xor_(_SIGN64)
    dist.all_reduce(keys_i64, op=dist.ReduceOp.MAX, group=group)
    keys_i64.bitwise_xor_(_SIGN64)
    return keys_i64


def all_reduce_max_i64_(keys_i64: torch.Tensor, group=None) -> torch.Tensor:
    """In-place element-wise max across ranks of the packed insertion keys (63-bit, never negative) and of the
    generation-key slots that ride behind them."""
    _, size = world(group)
    if size > 1:
        dist.all_reduce(keys_i64, op=dist.ReduceOp.MAX, group=group)
    return keys_i64


def all_reduce_disjoint_rows_(staging: torch.Tensor, group=None) -> torch.Tensor:
    """In-place merge of per-cell staging rows of which at most ONE rank holds a non-zero copy: an integer SUM over
    the raw 32-bit patterns is then an exact bitwise merge (a float sum would lose the sign of -0.0)."""
    _, size = world(group)
    if size == 1:
        return staging
    dist.all_reduce(staging.view(torch.int32), op=dist.ReduceOp.SUM, group=group)
    return staging


def all_equal(x: torch.Tensor, group=None) -> bool:
    """Replica-consistency check: do all ranks hold bit-identical `x`?"""
    _, size = world(group)
    if size == 1:
        return True
    v = x.contiguous().view(torch.uint8).reshape(-1).to(torch.int64)
    chk = torch.stack([v.sum(), (v * (torch.arange(v.numel(), device=v.device) % 65521 + 1)).sum()])
    lo, hi = chk.clone(), chk.clone()
    dist.all_reduce(lo, op=dist.ReduceOp.MIN, group=group)
    dist.all_reduce(hi, op=dist.ReduceOp.MAX, group=group)
    return bool(torch.equal(lo, hi))


def all_reduce_min_int(value: int, device, group=None) -> int:
    """min over ranks of a small host integer (QDX_ERR_* codes are negative: the minimum is the worst error any rank saw)."""
    _, size = world(group)
    if size == 1:
        return int(value)
    t = torch.tensor([int(value)], dtype=torch.int64, device=device if dist.get_backend(group) == "nccl" else "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.MIN, group=group)
    return int(t.item())
