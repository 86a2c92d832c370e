#!/usr/bin/env python
"""tiny_imagenet layer at its roofline batch: fp32 I/O against bf16 I/O (forward, backward, Gcell-updates/s)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
from tests import cases as K, runners
kind, ctor, big_b, _ = bench.LAYERS["tiny"]
c = K.case("q", kind, B=big_b, perturb=False, **ctor)
layer = runners.make_cuda_layer(c)


def t(fn, n=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n


for dt in (torch.float32, torch.bfloat16):
    u = torch.randn(big_b, 3, 64, 64, device="cuda").to(dt)
    g = torch.randn_like(u)
    x = u.clone().requires_grad_(True)
    fwd = t(lambda: layer(x))
    y = layer(x)
    ps = [p for p in layer.parameters() if p.requires_grad]
    bwd = t(lambda: torch.autograd.grad(y, [x] + ps, g, retain_graph=True, allow_unused=True))
    cells = big_b * 3 * 64 * 64
    nb = 2 if dt == torch.bfloat16 else 4
    print(f"{str(dt):16s} fwd {fwd:.3f} ms ({2 * nb * cells / fwd / 1e6:.0f} GB/s)  bwd {bwd:.3f} ms ({3 * nb * cells / bwd / 1e6:.0f} GB/s)  "
          f"{cells / ((fwd + bwd) * 1e-3) / 1e9:.1f} Gcell-updates/s")
