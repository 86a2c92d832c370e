"""Run a few forward+backward passes of one layer (profiling target for ncu).

    python tools/prof_layer.py <layer> <batch> [iters] [--no-gin]
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

import bench  # noqa: E402
from tests import cases as K  # noqa: E402
from tests import runners  # noqa: E402


def main():
    name, B = sys.argv[1], int(sys.argv[2])
    iters = int(sys.argv[3]) if len(sys.argv) > 3 and not sys.argv[3].startswith("-") else 3
    need_gin = "--no-gin" not in sys.argv
    kind, ctor, _, _ = bench.LAYERS[name]
    c = K.case("prof", kind, B=B, perturb=False, **ctor)
    layer = runners.make_cuda_layer(c)
    u = torch.randn(B, *c.shape, device="cuda")
    g = torch.randn(B, *c.shape, device="cuda")
    x = u.clone().requires_grad_(need_gin)
    e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
    for it in range(iters):
        for p in layer.parameters():
            p.grad = None
        e0.record()
        y = layer(x)
        e1.record()
        y.backward(g)
        e2.record()
        torch.cuda.synchronize()
        print(f"iter {it}: fwd {e0.elapsed_time(e1):.3f} ms  bwd {e1.elapsed_time(e2):.3f} ms")


if __name__ == "__main__":
    main()
