#!/usr/bin/env python
"""Profiling target: a few forward + backward passes of a run-time-sized implicit layer (csrc/adi_generic.cu).

    python tools/generic_prof.py mnist 48 1 8192        # kind, size, channels, batch
"""
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from tests import cases as K, runners  # noqa: E402

kind, size, C, B = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
ctor = dict(size=size) if kind in ("mnist", "fashion") else dict(size=size, channels=C)
c = K.case("prof", kind, B=B, perturb=False, **ctor)
layer = runners.make_cuda_layer(c)
x = torch.randn(B, *c.shape, device="cuda", requires_grad=True)
g = torch.randn(B, *c.shape, device="cuda")
for it in range(3):
    e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
    e0.record(); y = layer(x); e1.record(); y.backward(g); e2.record(); torch.cuda.synchronize()
    print(f"iter {it}: fwd {e0.elapsed_time(e1):.3f} ms  bwd {e1.elapsed_time(e2):.3f} ms", flush=True)
