"""Import the reference's training scripts as modules, unmodified.

Only usable where /root/reference exists (the build container); tests that need it skip
elsewhere.  The scripts import matplotlib / seaborn / kagglehub at top level, which are not
installed, so inert stub modules are registered for exactly the names that fail to import.
Nothing is copied out of the reference and `__name__ != "__main__"`, so no training starts.
"""
from __future__ import annotations

import contextlib
import importlib
import importlib.util
import io
import os
import sys
import types

REFERENCE_DIR = os.environ.get("PDE_REFERENCE_DIR", "/root/reference")
SCRIPTS = ("mnist_test", "fashion_mnist", "cifar10", "cifar_2version", "SVHN",
           "emotion_recognition", "tiny_imagenet")
_cache = {}


def available() -> bool:
    return all(os.path.isfile(os.path.join(REFERENCE_DIR, s + ".py")) for s in SCRIPTS)


class _Inert:
    def __call__(self, *a, **k):
        return _Inert()

    def __getattr__(self, name):
        if name.startswith("__") and name.endswith("__"):
            raise AttributeError(name)
        return _Inert()

    def __iter__(self):
        return iter(())


def _stub(name: str):
    mod = types.ModuleType(name)

    def _getattr(attr, _name=name):
        if attr.startswith("__") and attr.endswith("__"):
            raise AttributeError(attr)
        return _Inert()

    mod.__getattr__ = _getattr
    sys.modules[name] = mod
    return mod


def _ensure_stubs():
    import torch  # noqa: F401  (must come before stubbing, see appendix C)
    import torchvision  # noqa: F401
    for name in ("matplotlib", "matplotlib.pyplot", "seaborn", "kagglehub"):
        if name in sys.modules:
            continue
        try:
            importlib.import_module(name)
        except Exception:
            _stub(name)
            if "." in name:
                parent, child = name.rsplit(".", 1)
                setattr(sys.modules[parent], child, sys.modules[name])


def load(script: str):
    """Return the reference script `script` (e.g. "mnist_test") as a module."""
    if script in _cache:
        return _cache[script]
    if not available():
        raise FileNotFoundError(f"reference not found under {REFERENCE_DIR}")
    _ensure_stubs()
    import torch
    bench_flag = torch.backends.cudnn.benchmark
    spec = importlib.util.spec_from_file_location("ref_" + script, os.path.join(REFERENCE_DIR, script + ".py"))
    mod = importlib.util.module_from_spec(spec)
    with contextlib.redirect_stdout(io.StringIO()):
        spec.loader.exec_module(mod)
    torch.backends.cudnn.benchmark = bench_flag  # cifar10.py:16 flips it at import time
    _cache[script] = mod
    return mod


def quiet(fn, *a, **k):
    """Call fn with stdout silenced (several reference constructors print)."""
    with contextlib.redirect_stdout(io.StringIO()):
        return fn(*a, **k)
