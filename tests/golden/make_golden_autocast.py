"""Fixture: the reference's CIFAR layer run the way its training script runs it -- inside
torch.autocast (cifar10.py:459, cifar_2version.py:521) -- next to its own plain fp32 run.

    python tests/golden/make_golden_autocast.py          (build container: needs /root/reference)

There is no GPU here, so the reference runs under CPU autocast (bfloat16): the channel-mixing
matmul (cifar10.py:71) is then computed in reduced precision, the Thomas sweeps promote back to fp32.
Writes tests/golden/autocast_cifar10.npz with both runs (outputs, grad_input, parameter gradients).
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from tests import cases as K, refload  # noqa: E402

CASES = [K.case("autocast_cifar10_pde3", "cifar10", B=4, **K.SCRIPT_INSTANCES["cifar10_pde3"]),
         K.case("autocast_cifar2_diffusion2", "cifar2", B=4, **K.SCRIPT_INSTANCES["cifar2_diffusion2"])]


def run(c, amp):
    script, cls = K.REF_CLASS[c.kind]
    layer = refload.quiet(getattr(refload.load(script), cls), **c.ctor)
    params, (u, g) = K.make_params(c), K.make_io(c)
    sd = layer.state_dict()
    for k, v in params.items():
        sd[k] = torch.from_numpy(np.asarray(v)).reshape(sd[k].shape)
    layer.load_state_dict(sd)
    x = torch.from_numpy(u).requires_grad_(True)
    with torch.autocast("cpu", dtype=torch.bfloat16, enabled=amp):
        y = layer(x)
    y.float().backward(torch.from_numpy(g))
    out = {"y": y.detach().float().numpy(), "gin": x.grad.numpy()}
    for k, p in layer.named_parameters():
        out["g_" + k] = p.grad.float().numpy()
    return out


def main():
    z = {}
    for c in CASES:
        plain, amp = run(c, False), run(c, True)
        for k in plain:
            z[f"{c.name}/fp32/{k}"] = plain[k]
            z[f"{c.name}/amp/{k}"] = amp[k]
            den = np.linalg.norm(plain[k].astype(np.float64))
            print(f"{c.name:30s} {k:22s} autocast vs fp32 rel-L2 {np.linalg.norm(amp[k].astype(np.float64) - plain[k]) / den:.2e}  dtype {amp[k].dtype}")
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "autocast_cifar.npz"), **z)


if __name__ == "__main__":
    main()
