"""Run one parity case through the imported reference, the C oracle, or the CUDA product."""
from __future__ import annotations

from typing import Dict

import numpy as np

from . import cases as K


# ------------------------------------------------------------------------------ reference
def run_reference(c: K.Case, params=None, io=None, double: bool = False) -> Dict[str, np.ndarray]:
    """Forward + backward of the reference's own nn.Module on CPU (fp32, or its fp64 twin)."""
    import torch
    from . import refload

    script, cls = K.REF_CLASS[c.kind]
    mod = refload.load(script)
    params = params if params is not None else K.make_params(c)
    u, g = io if io is not None else K.make_io(c)
    old = torch.get_default_dtype()
    try:
        if double:
            # smooth_coefficients builds its kernel in the default dtype (mnist_test.py:144)
            torch.set_default_dtype(torch.float64)
        layer = refload.quiet(getattr(mod, cls), **c.ctor)
        if double:
            layer = layer.double()
        tdt = torch.float64 if double else torch.float32
        sd = layer.state_dict()
        for k, v in params.items():
            sd[k] = torch.from_numpy(np.asarray(v)).to(tdt).reshape(sd[k].shape)
        layer.load_state_dict(sd)
        x = torch.from_numpy(u).to(tdt).requires_grad_(True)
        y = layer(x)
        y.backward(torch.from_numpy(g).to(tdt))
        out = {"y": y.detach().numpy().copy(), "gin": x.grad.detach().numpy().copy()}
        for k, p in layer.named_parameters():
            if p.grad is not None:
                out["g_" + k] = p.grad.detach().numpy().copy()
        return out
    finally:
        torch.set_default_dtype(old)


# --------------------------------------------------------------------------------- oracle
def oracle_spec(c: K.Case):
    import oracle as O
    if c.kind == "mnist":
        return O.spec_mnist(**c.ctor)
    if c.kind == "fashion":
        return O.spec_fashion(**c.ctor)
    if c.kind == "svhn":
        return O.spec_svhn(**c.ctor)
    if c.kind == "cifar10":
        return O.spec_cifar10(**c.ctor)
    if c.kind == "cifar2":
        return O.spec_cifar2(**c.ctor)
    if c.kind == "emotion":
        return O.EmoSpec(**c.ctor)
    if c.kind == "tiny":
        kw = {k: v for k, v in c.ctor.items() if k != "use_implicit"}
        return O.TinySpec(**kw)
    raise KeyError(c.kind)


def emotion_buffers(c: K.Case, dtype=np.float32):
    """The registered buffers x, y (emotion_recognition.py:73-74): torch.linspace in the
    default dtype (fp32; fp64 for the double twin, which is built under a float64 default)."""
    import torch
    Nx, Ny = c.ctor.get("Nx", 48), c.ctor.get("Ny", 48)
    Lx, Ly = c.ctor.get("Lx", 1.0), c.ctor.get("Ly", 1.0)
    tdt = torch.float64 if np.dtype(dtype) == np.float64 else torch.float32
    return torch.linspace(0, Lx, Nx, dtype=tdt).numpy(), torch.linspace(0, Ly, Ny, dtype=tdt).numpy()


def run_oracle(c: K.Case, params=None, io=None, dtype=np.float32, need_gin=True, nthreads=0,
               forward_only=False) -> Dict[str, np.ndarray]:
    import oracle as O
    params = params if params is not None else K.make_params(c)
    u, g = io if io is not None else K.make_io(c)
    u, g = u.astype(dtype), g.astype(dtype)
    P = {k: np.asarray(v).astype(dtype) for k, v in params.items()}
    spec = oracle_spec(c)
    out: Dict[str, np.ndarray] = {}
    if c.kind in ("mnist", "fashion", "svhn", "cifar10", "cifar2"):
        chan = P.get("channel_mixing", P.get("channel_coupling"))
        skipw = P.get("skip_weight")
        maps = (P["alpha_base"], P["beta_base"], P["alpha_time_coeff"], P["beta_time_coeff"])
        out["y"] = O.adi_forward(spec, u, *maps, chan=chan, skipw=skipw, nthreads=nthreads)
        if forward_only:
            return out
        r = O.adi_backward(spec, u, g, *maps, chan=chan, skipw=skipw, need_gin=need_gin, nthreads=nthreads)
        out["gin"] = r["gin"]
        shp = params["alpha_base"].shape
        for k in ("alpha_base", "beta_base", "alpha_time_coeff", "beta_time_coeff"):
            out["g_" + k] = r[k].reshape(shp)
        if "chan" in r:
            out["g_channel_mixing" if "channel_mixing" in P else "g_channel_coupling"] = r["chan"]
        if "skip_weight" in r:
            out["g_skip_weight"] = r["skip_weight"]
    elif c.kind == "emotion":
        xs, ys = emotion_buffers(c, dtype)
        w = np.array([P[k] for k in ("alpha_w1", "alpha_w2", "alpha_w3", "beta_w1", "beta_w2", "beta_w3")], dtype)
        out["y"] = O.emotion_forward(spec, u, w, xs.astype(dtype), ys.astype(dtype), nthreads=nthreads)
        if forward_only:
            return out
        r = O.emotion_backward(spec, u, g, w, xs.astype(dtype), ys.astype(dtype), need_gin=need_gin, nthreads=nthreads)
        out["gin"] = r["gin"]
        for i, k in enumerate(("alpha_w1", "alpha_w2", "alpha_w3", "beta_w1", "beta_w2", "beta_w3")):
            out["g_" + k] = np.asarray(r["w"][i])
    elif c.kind == "tiny":
        out["y"] = O.tiny_forward(spec, u, P["alpha_base"], P["channel_scaling"], nthreads=nthreads)
        if forward_only:
            return out
        r = O.tiny_backward(spec, u, g, P["alpha_base"], P["channel_scaling"], need_gin=need_gin, nthreads=nthreads)
        out["gin"] = r["gin"]
        out["g_alpha_base"] = r["alpha_base"]
        out["g_channel_scaling"] = r["channel_scaling"]
    else:
        raise KeyError(c.kind)
    return out


# ------------------------------------------------------------------------------- compare
def rel_l2(a, b) -> float:
    a = np.asarray(a, np.float64).ravel()
    b = np.asarray(b, np.float64).ravel()
    den = np.linalg.norm(b)
    if den == 0.0:
        return float(np.linalg.norm(a))
    return float(np.linalg.norm(a - b) / den)


def rel_max(a, b) -> float:
    a = np.asarray(a, np.float64).ravel()
    b = np.asarray(b, np.float64).ravel()
    den = np.max(np.abs(b)) if b.size else 0.0
    if den == 0.0:
        return float(np.max(np.abs(a))) if a.size else 0.0
    return float(np.max(np.abs(a - b)) / den)


def compare(got: Dict[str, np.ndarray], want: Dict[str, np.ndarray], keys=None) -> Dict[str, float]:
    """rel-L2 and max-abs/max-ref error per key (the two measures SURVEY.md section 4 names)."""
    res = {}
    for k in (keys or want.keys()):
        if want.get(k) is None or got.get(k) is None:
            continue
        res[k] = max(rel_l2(got[k], want[k]), rel_max(got[k], want[k]))
    return res


# ---------------------------------------------------------------------------- CUDA product
OUR_CLASS = {
    "mnist": ("mnist_test", "DiffusionLayer"),
    "fashion": ("fashion_mnist", "DiffusionLayer"),
    "svhn": ("SVHN", "DiffusionLayer"),
    "cifar10": ("cifar10", "EnhancedDiffusionLayer"),
    "cifar2": ("cifar_2version", "LearnableDiffusionLayer"),
    "emotion": ("emotion_recognition", "PDELayer"),
    "tiny": ("tiny_imagenet", "ImprovedDiffusionLayer"),
}


def make_cuda_layer(c: K.Case, params=None, device="cuda"):
    import importlib
    import torch
    import cnn_with_pde_b200  # noqa: F401
    mod_name, cls = OUR_CLASS[c.kind]
    mod = importlib.import_module("cnn_with_pde_b200." + mod_name)
    layer = getattr(mod, cls)(**c.ctor)
    params = params if params is not None else K.make_params(c)
    sd = layer.state_dict()
    for k, v in params.items():
        sd[k] = torch.from_numpy(np.asarray(v)).reshape(sd[k].shape)
    layer.load_state_dict(sd)
    return layer.to(device)


def run_cuda(c: K.Case, params=None, io=None, need_gin=True) -> Dict[str, np.ndarray]:
    """Forward + backward of OUR module on cuda:0, through the C ABI."""
    import torch
    layer = make_cuda_layer(c, params)
    u, g = io if io is not None else K.make_io(c)
    x = torch.from_numpy(u).cuda().requires_grad_(need_gin)
    y = layer(x)
    y.backward(torch.from_numpy(g).cuda())
    torch.cuda.synchronize()
    out = {"y": y.detach().cpu().numpy(), "gin": x.grad.detach().cpu().numpy() if need_gin else None}
    for k, p in layer.named_parameters():
        if p.grad is not None:
            out["g_" + k] = p.grad.detach().cpu().numpy()
    return out
