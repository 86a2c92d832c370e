"""Data-parallel plumbing for the PDE layers (one process per GPU, torch.distributed).

The path shards over the batch with no exchange inside the layer: every
(b, c) plane is independent.  The only collective is the sum of the coefficient gradients over
ranks -- the same all-reduce DDP performs -- done here on one flat buffer so the <= 49 KB of
PDE gradients cost a single NCCL launch.
"""
from __future__ import annotations

from typing import Iterable, List

import torch
import torch.distributed as dist


def shard_bounds(n: int, rank: int, world: int):
    """Contiguous, balanced [lo, hi) slice of a batch of n for `rank` (ragged tails allowed)."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def allreduce_coefficient_grads(params: Iterable[torch.nn.Parameter], group=None, average: bool = False) -> List[torch.Tensor]:
    """Sum (or average, DDP-style) .grad of `params` across ranks in one flat all-reduce.
    Parameters without a gradient (tiny_imagenet's unused beta_base) are skipped."""
    grads = [p.grad for p in params if p.grad is not None]
    if not grads or not dist.is_available() or not dist.is_initialized():
        return grads
    flat = torch.cat([g.reshape(-1) for g in grads])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    if average:
        flat /= dist.get_world_size(group)
    off = 0
    for g in grads:
        n = g.numel()
        g.copy_(flat[off:off + n].view_as(g))
        off += n
    return grads
