#!/usr/bin/env python
"""Forward / backward time of the run-time-sized implicit kernels (csrc/adi_generic.cu) at a few plane sizes."""
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from tests import cases as K, runners  # noqa: E402


def timed(fn, n=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


for name, kind, B, ctor in (("mnist 1x48x48, 10 steps", "mnist", 32768, dict(size=48)),
                            ("svhn 3x64x64, 10 steps", "svhn", 4096, dict(size=64, channels=3)),
                            ("cifar10 3x36x36, 5 steps", "cifar10", 16384, dict(size=36, channels=3, dt=0.001, num_steps=5)),
                            ("cifar10 3x32x32, 5 steps (specialised kernels)", "cifar10", 16384, dict(size=32, channels=3, dt=0.001, num_steps=5)),
                            ("mnist 1x128x128, 10 steps", "mnist", 4096, dict(size=128)),
                            ("mnist 1x48x48, 10 steps (again)", "mnist", 32768, dict(size=48))):
    c = K.case("probe", kind, B=B, perturb=False, **ctor)
    layer = runners.make_cuda_layer(c)
    x = torch.randn(B, *c.shape, device="cuda", requires_grad=True)
    g = torch.randn(B, *c.shape, device="cuda")
    with torch.no_grad():
        t_inf = timed(lambda: layer(x))
    y = layer(x)
    t_fwd = timed(lambda: layer(x))
    t_bwd = timed(lambda: y.backward(g, retain_graph=True))
    each = [round(timed(lambda: y.backward(g, retain_graph=True), n=1), 1) for _ in range(6)]
    steps = layer.num_steps
    cu = B * x[0].numel() * steps / 1e9
    print(f"{name}: B {B}  inference {t_inf:.3f} ms  fwd {t_fwd:.3f} ms  bwd {t_bwd:.3f} ms  "
          f"fwd+bwd {cu / ((t_fwd + t_bwd) * 1e-3):.1f} Gcell-updates/s  single backward calls {each}", flush=True)
