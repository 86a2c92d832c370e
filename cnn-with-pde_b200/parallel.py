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


_ONES = {}   # (device, dtype, n) -> cached presence flags for the common case "every gradient is there"


def allreduce_coefficient_grads(params: Iterable[torch.nn.Parameter], group=None, average: bool = False) -> List[torch.Tensor]:
    """Sum (or average, DDP-style) .grad of `params` across ranks in one flat all-reduce.

    The flat buffer is laid out from the rank-invariant list "every parameter that requires a
    gradient", with zeros where this rank has none (an empty shard, a branch skipped on one rank), so
    all ranks always exchange the same number of elements.  A presence flag per parameter rides along:
    a parameter no rank has a gradient for (tiny_imagenet's unused beta_base) keeps `.grad is None`
    everywhere, as in the reference; one that some rank has a gradient for gets the sum on every rank.
    Three launches besides the collective (cat, scale if averaging, one multi-tensor copy back)."""
    plist = [p for p in params if p.requires_grad]
    if not plist:
        return []
    if not dist.is_available() or not dist.is_initialized():
        return [p.grad for p in plist if p.grad is not None]
    dev, dtype, n = plist[0].device, plist[0].dtype, len(plist)
    missing = [p.grad is None for p in plist]
    if any(missing):
        present = torch.tensor([0.0 if m else 1.0 for m in missing], dtype=dtype, device=dev)
    else:
        key = (dev, dtype, n)
        present = _ONES.get(key)
        if present is None:
            present = _ONES[key] = torch.ones(n, dtype=dtype, device=dev)
    pieces = [(torch.zeros_like(p) if m else p.grad).reshape(-1).to(dtype) for p, m in zip(plist, missing)]
    flat = torch.cat(pieces + [present])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    n_total = flat.numel() - n
    seen = flat[n_total:].tolist() if any(missing) else None
    if average:
        flat[:n_total] /= dist.get_world_size(group)
    out, dst, src, off = [], [], [], 0
    for i, p in enumerate(plist):
        k = p.numel()
        piece = flat[off:off + k].view_as(p)
        off += k
        if p.grad is not None:
            dst.append(p.grad)
            src.append(piece)
        elif seen is not None and seen[i] > 0:
            p.grad = piece.to(p.dtype).clone()
        if p.grad is not None:
            out.append(p.grad)
    if dst:
        torch._foreach_copy_(dst, src)
    return out
