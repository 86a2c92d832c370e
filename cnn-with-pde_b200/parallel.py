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

    The flat buffer is laid out from the rank-invariant list "every parameter that requires a
    gradient", with zeros where this rank has none (an empty shard, a branch skipped on one rank), so
    all ranks always exchange the same number of elements.  A presence flag per parameter rides along:
    a parameter no rank has a gradient for (tiny_imagenet's unused beta_base) keeps `.grad is None`
    everywhere, as in the reference; one that some rank has a gradient for gets the sum on every rank."""
    plist = [p for p in params if p.requires_grad]
    if not plist:
        return []
    if not dist.is_available() or not dist.is_initialized():
        return [p.grad for p in plist if p.grad is not None]
    dev = plist[0].device
    dtype = plist[0].dtype
    pieces = [(p.grad if p.grad is not None else torch.zeros_like(p)).reshape(-1).to(dtype) for p in plist]
    present = torch.tensor([0.0 if p.grad is None else 1.0 for p in plist], dtype=dtype, device=dev)
    flat = torch.cat(pieces + [present])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    n_total = flat.numel() - len(plist)
    seen = flat[n_total:].tolist() if any(p.grad is None for p in plist) else None
    if average:
        flat[:n_total] /= dist.get_world_size(group)
    out, off = [], 0
    for i, p in enumerate(plist):
        n = p.numel()
        piece = flat[off:off + n].view_as(p)
        off += n
        if p.grad is not None:
            p.grad.copy_(piece)
        elif seen is not None and seen[i] > 0:
            p.grad = piece.to(p.dtype).clone()
        if p.grad is not None:
            out.append(p.grad)
    return out
