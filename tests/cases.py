"""Shared parity cases: layer kinds, constructor arguments, seeded inputs and weights.

Used by tests/golden/make_golden.py (runs the imported reference), by the oracle tests and
by the GPU parity tests, so that all three see the same numbers.  Inputs are produced with
numpy's RandomState so they do not depend on the torch RNG implementation.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, Tuple

import numpy as np

# kind -> (reference script, reference class name)
REF_CLASS = {
    "mnist": ("mnist_test", "DiffusionLayer"),
    "fashion": ("fashion_mnist", "DiffusionLayer"),
    "svhn": ("SVHN", "DiffusionLayer"),
    "cifar10": ("cifar10", "EnhancedDiffusionLayer"),
    "cifar2": ("cifar_2version", "LearnableDiffusionLayer"),
    "emotion": ("emotion_recognition", "PDELayer"),
    "tiny": ("tiny_imagenet", "ImprovedDiffusionLayer"),
}


@dataclass
class Case:
    name: str
    kind: str
    ctor: Dict = field(default_factory=dict)     # kwargs exactly as the reference ctor takes them
    B: int = 2
    perturb: bool = True                          # weights moved away from the constant init
    seed: int = 1234
    shape: Tuple[int, int, int] = (1, 28, 28)     # (C, H, W)


def _shape(kind, ctor):
    if kind in ("mnist", "fashion"):
        n = ctor.get("size", 28)
        return (1, n, n)
    if kind in ("svhn", "cifar10", "cifar2"):
        n = ctor.get("size", 32)
        return (ctor.get("channels", 3), n, n)
    if kind == "emotion":
        return (1, ctor.get("Nx", 48), ctor.get("Ny", 48))
    if kind == "tiny":
        n = ctor.get("size", 64)
        return (ctor.get("channels", 3), n, n)
    raise KeyError(kind)


def case(name, kind, B=2, perturb=True, seed=1234, **ctor) -> Case:
    return Case(name=name, kind=kind, ctor=ctor, B=B, perturb=perturb, seed=seed, shape=_shape(kind, ctor))


# The layer instances the reference scripts actually construct ...
SCRIPT_INSTANCES = {
    "mnist": dict(),                                                   # mnist_test.py:226
    "fashion": dict(),                                                 # fashion_mnist.py:203
    "svhn": dict(size=32, channels=3),                                 # SVHN.py:238
    "cifar10_pde1": dict(size=32, channels=3, dt=0.001, num_steps=5, dx=1.0, dy=1.0),   # cifar10.py:253
    "cifar10_pde2": dict(size=32, channels=3, dt=0.002, num_steps=8, dx=2.0, dy=2.0),   # cifar10.py:255
    "cifar10_pde3": dict(size=32, channels=3, dt=0.005, num_steps=4, dx=1.5, dy=1.5),   # cifar10.py:257
    "cifar2_diffusion1": dict(size=32, channels=3, dt=0.001, num_steps=8),              # cifar_2version.py:269
    "cifar2_diffusion2": dict(size=32, channels=3, dt=0.002, num_steps=5),              # cifar_2version.py:270
    "emotion": dict(Nx=48, Ny=48),                                     # emotion_recognition.py:173
    "tiny": dict(size=64, channels=3, num_steps=1, use_implicit=False),  # tiny_imagenet.py:243
}

# ... each at its default init and at a perturbed weight set, small batch (golden fixtures).
GOLDEN_CASES = [
    case("mnist_init", "mnist", B=3, perturb=False),
    case("mnist_pert", "mnist", B=3),
    case("mnist_dxdy", "mnist", B=1, dx=0.7, dy=1.3, dt=0.05, num_steps=3, size=12),
    case("fashion_init", "fashion", B=3, perturb=False),
    case("fashion_pert", "fashion", B=3),
    case("svhn_init", "svhn", B=2, perturb=False, **SCRIPT_INSTANCES["svhn"]),
    case("svhn_pert", "svhn", B=2, **SCRIPT_INSTANCES["svhn"]),
    case("cifar10_pde1_pert", "cifar10", B=2, **SCRIPT_INSTANCES["cifar10_pde1"]),
    case("cifar10_pde2_pert", "cifar10", B=2, **SCRIPT_INSTANCES["cifar10_pde2"]),
    case("cifar10_pde3_init", "cifar10", B=2, perturb=False, **SCRIPT_INSTANCES["cifar10_pde3"]),
    case("cifar2_diffusion1_pert", "cifar2", B=2, **SCRIPT_INSTANCES["cifar2_diffusion1"]),
    case("cifar2_diffusion2_init", "cifar2", B=2, perturb=False, **SCRIPT_INSTANCES["cifar2_diffusion2"]),
    case("emotion_init", "emotion", B=2, perturb=False, **SCRIPT_INSTANCES["emotion"]),
    case("emotion_pert", "emotion", B=3, **SCRIPT_INSTANCES["emotion"]),
    case("tiny_init", "tiny", B=1, perturb=False, **SCRIPT_INSTANCES["tiny"]),
    case("tiny_pert", "tiny", B=2, size=16, channels=3, num_steps=3, dt=0.02),
    # plane sizes no script uses (csrc/adi_generic.cu): appended, the indices above are referred to elsewhere
    case("svhn_7x7_pert", "svhn", B=2, size=7, channels=3, num_steps=2),
    case("cifar10_c4_30x30_pert", "cifar10", B=2, size=30, channels=4, dt=0.01, num_steps=2, dx=1.0, dy=1.5),
    case("cifar2_c2_36x36_pert", "cifar2", B=2, size=36, channels=2, dt=0.02, num_steps=3),
    case("mnist_48x48_pert", "mnist", B=2, size=48, num_steps=2, dt=0.05),
]

# BASELINE.json config shapes (full batch); run against the live reference / the oracle.
CONFIG_CASES = [
    case("C1_mnist", "mnist", B=64, perturb=True),
    case("C2_fashion", "fashion", B=256, perturb=True),
    case("C3_cifar10_pde1", "cifar10", B=512, **SCRIPT_INSTANCES["cifar10_pde1"]),
    case("C3_cifar10_pde2", "cifar10", B=512, **SCRIPT_INSTANCES["cifar10_pde2"]),
    case("C3_cifar10_pde3", "cifar10", B=512, **SCRIPT_INSTANCES["cifar10_pde3"]),
    case("C3_cifar2_diffusion1", "cifar2", B=512, **SCRIPT_INSTANCES["cifar2_diffusion1"]),
    case("C3_cifar2_diffusion2", "cifar2", B=512, **SCRIPT_INSTANCES["cifar2_diffusion2"]),
    case("C4_svhn", "svhn", B=256, **SCRIPT_INSTANCES["svhn"]),
    case("C4_emotion", "emotion", B=64, perturb=False, **SCRIPT_INSTANCES["emotion"]),
    case("C5_tiny_b32", "tiny", B=32, **SCRIPT_INSTANCES["tiny"]),
    case("C5_tiny_b512", "tiny", B=512, perturb=False, **SCRIPT_INSTANCES["tiny"]),
]


# tiny_imagenet.ImprovedDiffusionLayer's dormant methods (tiny_imagenet.py:88-233):
# (fixture name, method, planes (B, H, W), the layer's dt, the method's positional arguments after u)
TINY_SPLIT_CASES = [
    ("adi_64", "implicit_diffusion_step", (3, 64, 64), 0.01, (0.05, 0.05)),          # the model's own size / init values
    ("adi_64_strong", "implicit_diffusion_step", (2, 64, 64), 0.5, (3.0, 7.5)),      # r = 0.75, 1.9
    ("adi_48x40", "implicit_diffusion_step", (5, 48, 40), 0.2, (1.3, 0.4)),          # H != W, W not a multiple of 8
    ("solve_x_33", "solve_implicit_x", (2, 33, 33), 0.01, (0.9, 0.3)),               # odd edge, explicit dt argument
    ("solve_y_64", "solve_implicit_y", (2, 64, 64), 0.01, (2.5, 0.1)),
    ("solve_x_clamped", "solve_implicit_x", (2, 6, 6), 0.01, (-0.3, 1.0)),           # negative r: pivots hit the clamp
    ("explicit_x_64", "diffuse_x_explicit", (3, 64, 64), 0.01, (0.11,)),
    ("explicit_y_30x64", "diffuse_y_explicit", (3, 30, 64), 0.02, (0.13,)),
]


def default_params(c: Case) -> Dict[str, np.ndarray]:
    """Reference init values, deterministic parts only (SURVEY.md section 8b state_dict row)."""
    C, H, W = c.shape
    f = np.float32
    if c.kind == "mnist":
        return dict(alpha_base=np.full((H, W), 2.0, f), beta_base=np.full((H, W), 2.0, f),
                    alpha_time_coeff=np.zeros((H, W), f), beta_time_coeff=np.zeros((H, W), f))
    if c.kind == "fashion":
        return dict(alpha_base=np.full((H, W), 1.8, f), beta_base=np.full((H, W), 1.8, f),
                    alpha_time_coeff=np.zeros((H, W), f), beta_time_coeff=np.zeros((H, W), f))
    if c.kind == "svhn":
        rs = np.random.RandomState(c.seed + 7)
        return dict(alpha_base=np.full((C, H, W), 0.1, f), beta_base=np.full((C, H, W), 0.1, f),
                    alpha_time_coeff=(rs.randn(C, H, W) * 0.001).astype(f),
                    beta_time_coeff=(rs.randn(C, H, W) * 0.001).astype(f),
                    channel_coupling=(np.eye(C) * 0.01).astype(f), skip_weight=np.array(0.9, f))
    if c.kind in ("cifar10", "cifar2"):
        rs = np.random.RandomState(c.seed + 7)
        return dict(alpha_base=np.ones((C, H, W), f), beta_base=np.ones((C, H, W), f),
                    alpha_time_coeff=np.zeros((C, H, W), f), beta_time_coeff=np.zeros((C, H, W), f),
                    channel_mixing=(np.eye(C) + 0.01 * rs.randn(C, C)).astype(f))
    if c.kind == "emotion":
        return dict(alpha_w1=np.array(0.1, f), alpha_w2=np.array(0.1, f), alpha_w3=np.array(0.1, f),
                    beta_w1=np.array(0.3, f), beta_w2=np.array(0.2, f), beta_w3=np.array(0.2, f))
    if c.kind == "tiny":
        return dict(alpha_base=np.full((C,), 0.05, f), beta_base=np.full((C,), 0.05, f),
                    channel_scaling=np.ones((C,), f))
    raise KeyError(c.kind)


def make_params(c: Case) -> Dict[str, np.ndarray]:
    p = default_params(c)
    if not c.perturb:
        return p
    rs = np.random.RandomState(c.seed + 11)
    f = np.float32
    if c.kind in ("mnist", "fashion", "svhn", "cifar10", "cifar2"):
        dt = c.ctor.get("dt", {"mnist": 0.001, "fashion": 0.3, "svhn": 0.01}.get(c.kind, 0.001))
        steps = c.ctor.get("num_steps", {"fashion": 4}.get(c.kind, 10))
        T = dt * steps
        for k in ("alpha_base", "beta_base"):
            a = p[k].astype(np.float64) + 0.3 * rs.randn(*p[k].shape)
            flat = a.reshape(-1)
            n = 12 if flat.size >= 12 else 2               # (2 x 2 planes: one cell on either side)
            idx = rs.choice(flat.size, size=n, replace=False)
            flat[idx[:n // 2]] = -0.5 + 0.4 * rs.rand(n // 2)   # forced below clamp min
            flat[idx[n // 2:]] = 10.5 + rs.rand(n // 2)         # forced above the CIFAR clamp max
            p[k] = a.astype(f)
        for k in ("alpha_time_coeff", "beta_time_coeff"):
            # strong enough that some cells cross a clamp edge inside [0, T]
            p[k] = (rs.randn(*p[k].shape) * (0.5 / max(T, 1e-9))).astype(f)
        for k in ("channel_mixing", "channel_coupling"):
            if k in p:
                C = p[k].shape[0]
                p[k] = (np.eye(C) + 0.2 * rs.randn(C, C)).astype(f)
        if "skip_weight" in p:
            p["skip_weight"] = np.array(rs.randn() * 0.8, f)
    elif c.kind == "emotion":
        # keep the scheme stable-ish so values stay O(1): shrink the weights
        for k in list(p):
            p[k] = np.array(float(p[k]) * 0.05 * (1 + 0.5 * rs.randn()), f)
    elif c.kind == "tiny":
        C = p["alpha_base"].shape[0]
        ab = 0.05 + 0.04 * rs.randn(C)
        if C >= 3:
            ab[1] = 0.4      # above clamp max 0.15 -> zero gradient
            ab[2] = -1.0     # below clamp min
        p["alpha_base"] = ab.astype(f)
        p["channel_scaling"] = (1 + 0.3 * rs.randn(C)).astype(f)
    return p


def make_io(c: Case) -> Tuple[np.ndarray, np.ndarray]:
    rs = np.random.RandomState(c.seed)
    C, H, W = c.shape
    u = rs.randn(c.B, C, H, W).astype(np.float32)
    g = rs.randn(c.B, C, H, W).astype(np.float32)
    return u, g


# Parameters that receive a gradient (tiny's beta_base never does: tiny_imagenet.py:26,40).
def grad_names(c: Case):
    names = list(default_params(c).keys())
    if c.kind == "tiny":
        names.remove("beta_base")
    return names
