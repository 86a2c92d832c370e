#!/usr/bin/env python
"""Inference forward (torch.no_grad) at the script batch: device time per call, CUDA-graph replay."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
from tests import cases as K, runners
for name in (sys.argv[1:] or ["fashion", "mnist", "cifar10_pde1", "cifar10_pde2", "svhn"]):
    kind, ctor, _, b = bench.LAYERS[name]
    c = K.case("inf_" + name, kind, B=b, perturb=False, **ctor)
    layer = runners.make_cuda_layer(c)
    u = torch.randn(b, *c.shape, device="cuda")
    res = []
    for env in ({}, {"PDE_B200_ADI_LEGACY": "1"}):
        os.environ.update(env)
        s = torch.cuda.Stream(); s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s), torch.no_grad():
            import cnn_with_pde_b200.functional as F
            F._table_cache.clear()
            for _ in range(3): layer(u)
            s.synchronize()
            cg = torch.cuda.CUDAGraph()
            with torch.cuda.graph(cg, stream=s, capture_error_mode="thread_local"):
                y = layer(u)
        torch.cuda.current_stream().wait_stream(s)
        cg.replay(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(200): cg.replay()
        e1.record(); torch.cuda.synchronize()
        res.append(e0.elapsed_time(e1) / 200 * 1e3)
        for k in env: os.environ.pop(k)
    print(f"{name:16s} B={b:4d} inference forward (prepare + kernel): default {res[0]:6.1f} us   whole-line {res[1]:6.1f} us")
