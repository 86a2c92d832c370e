#!/usr/bin/env python
"""cProfile of the host side of a forward+backward call at the script batch (GPU box)."""
import cProfile
import os
import pstats
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import bench  # noqa: E402
from tests import cases as K, runners  # noqa: E402
import cnn_with_pde_b200.functional as F  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "fashion"
kind, ctor, _, b = bench.LAYERS[name]
c = K.case("lat", kind, B=b, perturb=False, **ctor)
layer = runners.make_cuda_layer(c)
params = list(layer.parameters())
u = torch.randn(b, *c.shape, device="cuda")
g = torch.randn(b, *c.shape, device="cuda")
x = u.clone().requires_grad_(True)

# wall time inside the Function's forward / backward bodies
acc = {"f": 0.0, "b": 0.0, "n": 0}
Fn = F._AdiFunction if kind not in ("emotion", "tiny") else (F._EmotionFunction if kind == "emotion" else F._TinyFunction)
of, ob = Fn.forward, Fn.backward


def call():
    for p in params:
        p.grad = None
    y = layer(x)
    y.backward(g)


for _ in range(20):
    call()
torch.cuda.synchronize()
n = 500
t0 = time.perf_counter()
for _ in range(n):
    call()
t1 = time.perf_counter()
torch.cuda.synchronize()
print(f"{name}: {1e6 * (t1 - t0) / n:.1f} us per fwd+bwd (host)")
t0 = time.perf_counter()
for _ in range(n):
    y = layer(x)
t1 = time.perf_counter()
torch.cuda.synchronize()
print(f"{name}: {1e6 * (t1 - t0) / n:.1f} us per training forward (host)")
y = layer(x)
t0 = time.perf_counter()
for _ in range(n):
    torch.autograd.grad(y, [x] + params, g, retain_graph=True, allow_unused=True)
t1 = time.perf_counter()
torch.cuda.synchronize()
print(f"{name}: {1e6 * (t1 - t0) / n:.1f} us per backward via autograd.grad (host)")
t0 = time.perf_counter()
for _ in range(n):
    for p in params:
        p.grad = None
t1 = time.perf_counter()
print(f"{name}: {1e6 * (t1 - t0) / n:.1f} us per grad reset")
pr = cProfile.Profile()
pr.enable()
for _ in range(n):
    call()
pr.disable()
torch.cuda.synchronize()
st = pstats.Stats(pr)
st.sort_stats("cumulative").print_stats(28)
