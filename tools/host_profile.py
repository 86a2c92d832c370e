#!/usr/bin/env python
"""Host-side cost of a forward+backward call at the script batch (run on the GPU box):
time inside the autograd.Function bodies vs the autograd engine around them, and the cost of the
individual C-ABI calls."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import bench  # noqa: E402
from tests import cases as K, runners  # noqa: E402
import cnn_with_pde_b200.functional as F  # noqa: E402
from cnn_with_pde_b200 import _cabi  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "fashion"
kind, ctor, _, b = bench.LAYERS[name]
c = K.case("lat", kind, B=b, perturb=False, **ctor)
layer = runners.make_cuda_layer(c)
params = list(layer.parameters())
u = torch.randn(b, *c.shape, device="cuda")
g = torch.randn(b, *c.shape, device="cuda")
x = u.clone().requires_grad_(True)
Fn = F._AdiFunction if kind not in ("emotion", "tiny") else (F._EmotionFunction if kind == "emotion" else F._TinyFunction)
acc = {"f": 0, "b": 0}
of, ob = Fn.forward, Fn.backward


def tf(*a, **k):
    t0 = time.perf_counter_ns()
    r = of(*a, **k)
    acc["f"] += time.perf_counter_ns() - t0
    return r


def tb(*a, **k):
    t0 = time.perf_counter_ns()
    r = ob(*a, **k)
    acc["b"] += time.perf_counter_ns() - t0
    return r


def call():
    for p in params:
        p.grad = None
    layer(x).backward(g)


def bench_loop(n=1000):
    for _ in range(20):
        call()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n):
        call()
    dt = (time.perf_counter() - t0) / n * 1e6
    torch.cuda.synchronize()
    return dt


print(f"{name}: {bench_loop():.1f} us per fwd+bwd (host), plain")
Fn.forward, Fn.backward = staticmethod(tf), staticmethod(tb)
acc["f"] = acc["b"] = 0
n = 1000
tot = bench_loop(n)
print(f"{name}: {tot:.1f} us per fwd+bwd (host), instrumented: inside forward() {acc['f'] / (n + 20) / 1e3:.1f} us, "
      f"inside backward() {acc['b'] / (n + 20) / 1e3:.1f} us, rest (module call, autograd engine, grad accumulation) "
      f"{tot - (acc['f'] + acc['b']) / (n + 20) / 1e3:.1f} us")
Fn.forward, Fn.backward = of, ob

# individual pieces
L = _cabi.lib()


def t(fn, n=3000):
    for _ in range(50):
        fn()
    t0 = time.perf_counter()
    for _ in range(n):
        fn()
    r = (time.perf_counter() - t0) / n * 1e6
    torch.cuda.synchronize()
    return r


if kind not in ("emotion", "tiny"):
    cfg = layer._config()
    plan = F._adi_plan(cfg, b, 0, 0)
    maps = [layer.alpha_base, layer.beta_base, layer.alpha_time_coeff, layer.beta_time_coeff]
    tables = F._bytes(plan.tables_bytes, u.device)
    st = F._stream(u.device)
    print("  pde_adi_prepare call            %.1f us" % t(lambda: L.pde_adi_prepare(plan.dref, plan.sref, *[p.data_ptr() for p in maps], tables.data_ptr(), st)))
    out = torch.empty_like(u)
    ck = F._bytes(plan.ckpt_bytes, u.device)
    chan = getattr(layer, "channel_mixing", getattr(layer, "channel_coupling", None))
    skw = getattr(layer, "skip_weight", None)
    print("  pde_adi_forward_train call      %.1f us" % t(lambda: L.pde_adi_forward_train(plan.dref, tables.data_ptr(), u.data_ptr(), F._ptr(chan), F._ptr(skw), out.data_ptr(), ck.data_ptr(), st)))
    ws = F._bytes(plan.ws_saved_bytes, u.device)
    gin = torch.empty_like(u)
    flat = torch.empty(plan.grad_numel + 16 + 1, device=u.device)
    pl = cfg.C * cfg.N * cfg.N
    gp = [flat.data_ptr() + 4 * k * pl for k in range(4)]
    gc = flat.data_ptr() + 4 * plan.grad_numel if chan is not None else None
    gs = gc + 64 if skw is not None else None
    print("  pde_adi_backward_saved call     %.1f us" % t(lambda: L.pde_adi_backward_saved(plan.dref, tables.data_ptr(), u.data_ptr(), g.data_ptr(), F._ptr(chan), F._ptr(skw), ck.data_ptr(), gin.data_ptr(), gp[0], gp[1], gp[2], gp[3], gc, gs, ws.data_ptr(), plan.ws_saved_bytes, st)))
    print("  env_tuning                      %.2f us" % t(F.env_tuning))
    print("  _adi_plan lookup                %.2f us" % t(lambda: F._adi_plan(cfg, b, 0, 0)))
    print("  layer._config()                 %.2f us" % t(layer._config))
    print("  _stream                         %.2f us" % t(lambda: F._stream(u.device)))
    print("  torch.empty_like                %.2f us" % t(lambda: torch.empty_like(u)))
    print("  _bytes(ws)                      %.2f us" % t(lambda: F._bytes(plan.ws_saved_bytes, u.device)))
    print("  4 views                         %.2f us" % t(lambda: [flat[k * pl:(k + 1) * pl].view(cfg.C, cfg.N, cfg.N) for k in range(4)]))
with torch.no_grad():
    print("  module call under no_grad       %.1f us" % t(lambda: layer(u), 1000))
