"""Generate tests/golden/model_*.npz from the UNMODIFIED reference classifiers (build container).

    python -m tests.golden.make_golden_models            # from the repo root

For each model: construct the reference class under a fixed seed, perturb the PDE coefficients
away from their constant init, run one forward + backward in eval mode (dropout off, batch norm
on its running statistics) with a cross-entropy loss on seeded inputs, and store the full
state_dict, the inputs, the logits, the loss and the gradients of every PDE-layer parameter.
"""
from __future__ import annotations

import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from tests import refload  # noqa: E402

# name -> (script, class, input shape, classes, batch)
MODELS = {
    "mnist": ("mnist_test", "PDEClassifier", (1, 28, 28), 10, 4),
    "fashion": ("fashion_mnist", "FashionPDEClassifier", (1, 28, 28), 10, 4),
    "cifar10": ("cifar10", "CIFAR10PDENoConv", (3, 32, 32), 10, 3),
}   # SVHN (9 M parameters = 33 MB fixture) and emotion (5 MB) are left out: their layers have fixtures of their own
PDE_PREFIXES = ("diff.", "pde.", "feature_extractor.pde")


def is_pde_param(name: str) -> bool:
    return name.startswith(PDE_PREFIXES)


def main():
    if not refload.available():
        raise SystemExit("reference not available; fixtures can only be generated in the build container")
    import torch
    torch.set_num_threads(1)
    for name, (script, cls, shape, classes, B) in MODELS.items():
        torch.manual_seed(77)
        model = refload.quiet(getattr(refload.load(script), cls))
        rng = np.random.RandomState(4242)
        with torch.no_grad():
            for n, p in model.named_parameters():
                if is_pde_param(n) and p.dim() >= 2 and "channel" not in n:
                    p.add_(torch.from_numpy((0.05 * rng.standard_normal(p.shape)).astype(np.float32)) * p.abs().mean().clamp(min=0.1))
        model.eval()
        x = torch.from_numpy(rng.standard_normal((B,) + shape).astype(np.float32))
        y = torch.from_numpy(rng.randint(0, classes, size=B).astype(np.int64))
        logits = model(x)
        loss = torch.nn.functional.cross_entropy(logits, y)
        loss.backward()
        blob = {"x": x.numpy(), "y": y.numpy(), "logits": logits.detach().numpy(), "loss": np.float32(loss.item())}
        for k, v in model.state_dict().items():
            blob["sd_" + k] = v.numpy()
        for n, p in model.named_parameters():
            if is_pde_param(n) and p.grad is not None:
                blob["g_" + n] = p.grad.numpy()
        path = os.path.join(HERE, f"model_{name}.npz")
        np.savez_compressed(path, **blob)
        print(f"{name:8s} {os.path.getsize(path)/1e6:6.2f} MB  loss {loss.item():.6f}")


if __name__ == "__main__":
    main()
