#!/usr/bin/env python
"""One small forward+backward of every kernel family and variant, checked against the oracle: the target
of the compute-sanitizer runs (tools/sanitize.sh).  Small batches: the sanitizer tools slow kernels 10-100x."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
from tests import cases as K, runners  # noqa: E402

CASES = [
    ({}, K.case("san_fashion", "fashion", B=19)),                                   # half-line, P = 2
    ({"PDE_B200_SPLIT_P": "4", "PDE_B200_SPLIT_QF": "4"}, K.case("san_fashion_p4", "fashion", B=70)),   # P = 4, Q = 4, ragged
    ({"PDE_B200_SPLIT_P": "4", "PDE_B200_SPLIT_QF": "2"}, K.case("san_mnist_p4", "mnist", B=21, num_steps=3)),
    ({}, K.case("san_cifar10", "cifar10", B=7, **K.SCRIPT_INSTANCES["cifar10_pde3"])),                 # pre-step mix
    ({"PDE_B200_SPLIT_QF": "2"}, K.case("san_cifar2", "cifar2", B=9, **K.SCRIPT_INSTANCES["cifar2_diffusion2"])),   # Lie
    ({}, K.case("san_svhn", "svhn", B=5, size=32, channels=3, num_steps=3)),                            # coupling + skip
    ({}, K.case("san_fashion_exact", "fashion", B=8, perturb=False, dt=5.0)),                           # exact mode
    ({"PDE_B200_ADI_LEGACY": "1"}, K.case("san_whole_fashion", "fashion", B=9)),                         # whole-line kernels
    ({"PDE_B200_ADI_LEGACY": "1"}, K.case("san_whole_cifar10", "cifar10", B=5, **K.SCRIPT_INSTANCES["cifar10_pde3"])),
    ({}, K.case("san_svhn16", "svhn", B=3, size=16, channels=3, num_steps=2)),                          # whole-line, N = 16
    ({}, K.case("san_emotion", "emotion", B=9)),                                                        # register-tiled
    ({"PDE_B200_EMO_TILED": "0"}, K.case("san_emotion_generic", "emotion", B=5, Nx=24, Ny=24)),
    ({}, K.case("san_tiny", "tiny", B=5, **K.SCRIPT_INSTANCES["tiny"])),                                # TMA ring
]


def main():
    only = sys.argv[1:] or None
    worst = 0.0
    for env, c in CASES:
        if only and not any(o in c.name for o in only):
            continue
        os.environ.update(env)
        try:
            params, io = K.make_params(c), K.make_io(c)
            got = runners.run_cuda(c, params=params, io=io)
            want = runners.run_oracle(c, params=params, io=io, dtype=np.float32)
            errs = runners.compare(got, want)
            worst = max(worst, max(errs.values()))
            print(f"{c.name:24s} worst rel-err {max(errs.values()):.2e}", flush=True)
            assert max(errs.values()) <= 1e-5, errs
        finally:
            for k in env:
                os.environ.pop(k, None)
    # the dormant tiny methods
    import torch
    import oracle as O
    from cnn_with_pde_b200.tiny_imagenet import ImprovedDiffusionLayer
    layer = ImprovedDiffusionLayer().cuda()
    u = torch.randn(5, 64, 64, device="cuda")
    y = layer.implicit_diffusion_step(u, 0.7, 1.9)
    want = O.tiny_split("implicit_diffusion_step", u.cpu().numpy(), coeff_x=0.7, coeff_y=1.9, dt=0.01)
    print(f"{'san_tiny_split':24s} worst rel-err {runners.rel_l2(y.cpu().numpy(), want):.2e}")
    print("all ok, worst", worst)


if __name__ == "__main__":
    main()
