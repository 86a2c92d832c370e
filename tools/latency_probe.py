#!/usr/bin/env python
"""Where a forward+backward call at the reference's own batch size spends its time.

    python tools/latency_probe.py [layer ...]

Per layer (bench.LAYERS, script batch): host time per eager call (no sync inside the loop), the
device time of the same work replayed from a CUDA graph (kernels only), and the eager wall time
per call with a sync -- the number bench.py reports as script_batch_latency_us.
"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

import bench  # noqa: E402
from tests import cases as K  # noqa: E402
from tests import runners  # noqa: E402


def probe(name, n=200):
    kind, ctor, _, b = bench.LAYERS[name]
    c = K.case("lat_" + name, kind, B=b, perturb=False, **ctor)
    layer = runners.make_cuda_layer(c)
    params = [p for p in layer.parameters()]
    gen = torch.Generator(device="cuda").manual_seed(1)
    u = torch.randn(b, *c.shape, device="cuda", generator=gen)
    g = torch.randn(b, *c.shape, device="cuda", generator=gen)
    x = u.clone().requires_grad_(True)

    def call():
        for p in params:
            p.grad = None
        layer(x).backward(g)

    def call_nogin():
        for p in params:
            p.grad = None
        layer(u).backward(g)

    for _ in range(10):
        call()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for _ in range(n):
        call()
    e1.record()
    host_us = (time.perf_counter() - t0) / n * 1e6
    torch.cuda.synchronize()
    eager_us = e0.elapsed_time(e1) / n * 1e3
    # forward only, host side
    t0 = time.perf_counter()
    with torch.no_grad():
        for _ in range(n):
            layer(u)
    fwd_host_us = (time.perf_counter() - t0) / n * 1e6
    torch.cuda.synchronize()
    # kernels only: the same call replayed from a CUDA graph
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for _ in range(3):
            call()
        s.synchronize()
        cg = torch.cuda.CUDAGraph()
        with torch.cuda.graph(cg, stream=s):
            call()
    torch.cuda.current_stream().wait_stream(s)
    cg.replay()
    torch.cuda.synchronize()
    e0.record()
    for _ in range(n):
        cg.replay()
    e1.record()
    torch.cuda.synchronize()
    graph_us = e0.elapsed_time(e1) / n * 1e3
    print(f"{name:18s} B={b:4d}  eager {eager_us:7.1f} us/call  host {host_us:7.1f} us/call  inference host {fwd_host_us:6.1f} us  "
          f"graph replay {graph_us:7.1f} us/call", flush=True)


if __name__ == "__main__":
    for name in (sys.argv[1:] or list(bench.LAYERS)):
        probe(name)
