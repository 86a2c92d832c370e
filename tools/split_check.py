"""Parity + timing of the split (half-line) ADI kernels against the oracle and the legacy kernels.

    python tools/split_check.py [parity] [time]
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

import bench  # noqa: E402
from tests import cases as K  # noqa: E402
from tests import runners  # noqa: E402


def parity():
    cases = [
        K.case("fashion", "fashion", B=9),
        K.case("fashion_init", "fashion", B=4, perturb=False),
        K.case("mnist", "mnist", B=5),
        K.case("cifar10_pde1", "cifar10", B=5, **K.SCRIPT_INSTANCES["cifar10_pde1"]),
        K.case("cifar10_pde2", "cifar10", B=3, **K.SCRIPT_INSTANCES["cifar10_pde2"]),
        K.case("cifar2", "cifar2", B=5, **K.SCRIPT_INSTANCES.get("cifar2_diffusion1", {})),
        K.case("svhn", "svhn", B=5, **K.SCRIPT_INSTANCES["svhn"]),
        K.case("fashion_dt5", "fashion", B=8, perturb=False, dt=5.0),
    ]
    worst = 0.0
    for forced, qf, qb in (("2", "1", "1"), ("2", "2", "2"), ("4", "4", "2"), ("4", "2", "1")):
        os.environ["PDE_B200_SPLIT_P"] = forced
        os.environ["PDE_B200_SPLIT_QF"] = qf
        os.environ["PDE_B200_SPLIT_QB"] = qb
        forced = f"{forced} Q={qf}/{qb}"
        for c in cases:
            params, io = K.make_params(c), K.make_io(c)
            want = runners.run_oracle(c, params=params, io=io, dtype=np.float32)
            for need_gin in (True, False):
                try:
                    got = runners.run_cuda(c, params=params, io=io, need_gin=need_gin)
                except Exception as e:  # noqa: BLE001
                    print(f"P={forced} {c.name} gin={need_gin}: EXC {e}")
                    continue
                w = dict(want)
                if not need_gin:
                    w.pop("gin")
                    got.pop("gin")
                errs = runners.compare(got, w)
                m = max(errs.values())
                worst = max(worst, m if m == m else 1e9)
                flag = "ok " if m <= 1e-5 else "BAD"
                print(f"P={forced} {c.name:14s} gin={int(need_gin)} {flag} max={m:.2e} " +
                      " ".join(f"{k}={v:.1e}" for k, v in errs.items() if not v <= 1e-5))
    for k in ("PDE_B200_SPLIT_P", "PDE_B200_SPLIT_QF", "PDE_B200_SPLIT_QB"):
        os.environ.pop(k, None)
    print("worst", worst)


def timing():
    variants = [("1", "", ""), ("0", "", ""), ("0", "2", "")]
    for name, B in (("fashion", 262144), ("mnist", 131072), ("cifar10_pde1", 65536)):
        kind, ctor, _, _ = bench.LAYERS[name]
        c = K.case("t", kind, B=B, perturb=False, **ctor)
        for legacy, qf, qb in variants:
            os.environ["PDE_B200_ADI_LEGACY"] = legacy
            os.environ["PDE_B200_SPLIT_QF"] = qf
            os.environ["PDE_B200_SPLIT_QB"] = qb
            layer = runners.make_cuda_layer(c)
            u = torch.randn(B, *c.shape, device="cuda")
            g = torch.randn(B, *c.shape, device="cuda")
            x = u.clone().requires_grad_(True)
            ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
            best = [1e9, 1e9, 1e9]
            for it in range(4):
                for p in layer.parameters():
                    p.grad = None
                with torch.no_grad():
                    ev[0].record()
                    layer(u)
                    ev[1].record()
                y = layer(x)
                ev[2].record()
                y.backward(g)
                ev[3].record()
                torch.cuda.synchronize()
                ts = [ev[0].elapsed_time(ev[1]), ev[1].elapsed_time(ev[2]), ev[2].elapsed_time(ev[3])]
                best = [min(a, b) for a, b in zip(best, ts)]
            print(f"{name} B={B} legacy={legacy} qf={qf or '-'} qb={qb or '-'}: fwd(eval) {best[0]:.3f}  fwd(train) {best[1]:.3f}  bwd {best[2]:.3f} ms")
            del layer, u, g, x, y
            torch.cuda.empty_cache()
    for k in ("PDE_B200_ADI_LEGACY", "PDE_B200_SPLIT_QF", "PDE_B200_SPLIT_QB"):
        os.environ.pop(k, None)


if __name__ == "__main__":
    what = sys.argv[1:] or ["parity", "time"]
    if "parity" in what:
        parity()
    if "time" in what:
        timing()
