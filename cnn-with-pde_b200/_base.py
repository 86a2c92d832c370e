"""Helpers shared by the drop-in modules (pure PyTorch, off the hot path)."""
from __future__ import annotations

import torch
import torch.nn.functional as F


def check_input(u: torch.Tensor, channels: int, size_h: int, size_w: int, who: str):
    if u.dim() != 4:
        raise ValueError(f"{who}: expected a 4-D NCHW tensor, got shape {tuple(u.shape)}")
    B, C, H, W = u.shape
    if C != channels:
        raise ValueError(f"{who}: expected {channels} channel(s), got {C}")
    if H != size_h or W != size_w:
        raise ValueError(f"{who}: expected {size_h}x{size_w} planes, got {H}x{W}")


def cached_config(module, key, build):
    """The layer's kernel configuration, rebuilt only when one of the plain attributes it is made of
    (`key`) has changed since the last call (they are ordinary, assignable attributes in the reference)."""
    hit = module.__dict__.get("_pde_cfg")
    if hit is None or hit[0] != key:
        hit = (key, build())
        module.__dict__["_pde_cfg"] = hit
    return hit[1]


def smooth_coefficients(coeffs: torch.Tensor, dim: int = 1, kernel_size: int = 3) -> torch.Tensor:
    """3-tap moving average with replicate padding along dim 1 of a (lines, N) tensor
    (what mnist_test.py:135-149 computes); kept as a helper, the kernels fuse it."""
    if kernel_size == 1:
        return coeffs
    if dim != 1:
        raise NotImplementedError("Only dim=1 smoothing implemented")
    pad = kernel_size // 2
    padded = F.pad(coeffs.unsqueeze(1), (pad, pad), mode="replicate")
    kernel = torch.ones(1, 1, kernel_size, device=coeffs.device, dtype=coeffs.dtype) / kernel_size
    return F.conv1d(padded, kernel, padding=0).squeeze(1)
