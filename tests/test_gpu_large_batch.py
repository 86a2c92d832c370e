"""Parity of the kernels `bench.py` actually times: the DEFAULT dispatch at large batch.

At the reference's own batch sizes (64 ... 512) the implicit layers are served by the whole-line
kernels of adi.cu; from ~1200 samples on the library switches, on its own, to the half-line kernels
of adi_split.cu (four sample pairs per group for the single-channel layers from ~2400 samples), and
that is what every roofline-size measurement runs.  These tests therefore use no tuning override:
batches of 16384+ samples, weights perturbed so that cells sit outside both clamp edges and the time
coefficients cross a clamp inside [0, T], and compare the output, grad_input and EVERY parameter
gradient (sums over the whole batch: fp32 accumulators in tensor memory, finished in double) with
the oracle in fp32 and fp64 (OpenMP on all host cores: seconds).

Bar: rel-err <= 1e-5 (north_star), both rel-L2 and max-abs / max-ref, and no further from the fp64
oracle than the fp32 oracle is, plus 1e-5.
"""
import os

import numpy as np
import pytest

from . import cases as K
from . import runners

pytestmark = pytest.mark.gpu
TOL = 1e-5

_LARGE = [
    K.case("large_fashion", "fashion", B=16384 + 6),                                   # ragged last group
    K.case("large_fashion_init", "fashion", B=32768, perturb=False),                   # the bench's weights
    K.case("large_mnist", "mnist", B=16384),
    K.case("large_cifar10_pde1", "cifar10", B=16384 + 3, **K.SCRIPT_INSTANCES["cifar10_pde1"]),
    K.case("large_cifar10_pde2", "cifar10", B=16384, **K.SCRIPT_INSTANCES["cifar10_pde2"]),
    K.case("large_cifar2_diffusion1", "cifar2", B=16384, **K.SCRIPT_INSTANCES["cifar2_diffusion1"]),
    K.case("large_svhn", "svhn", B=16384 + 1, **K.SCRIPT_INSTANCES["svhn"]),
]


def _check(c, got, o32, o64):
    errs = runners.compare(got, o32)
    assert set(errs) == {k for k, v in o32.items() if v is not None}, (sorted(errs), sorted(o32))
    bad = {k: e for k, e in errs.items() if not e <= TOL}
    assert not bad, f"{c.name} vs oracle fp32: {bad}"
    e_cuda, e_ref = runners.compare(got, o64), runners.compare(o32, o64)
    bad = {k: (e_cuda[k], e_ref[k]) for k in e_cuda if not e_cuda[k] <= e_ref[k] + TOL}
    assert not bad, f"{c.name}: further from fp64 than the fp32 oracle + 1e-5: {bad}"
    return errs


@pytest.mark.parametrize("c", _LARGE, ids=lambda c: c.name)
def test_default_dispatch_at_large_batch_matches_oracle(c):
    import cnn_with_pde_b200.functional as F
    assert F.env_tuning() == 0, "a PDE_B200_* tuning switch is set: this test must run the default dispatch"
    import cnn_with_pde_b200 as P
    from ctypes import byref
    params, io = K.make_params(c), K.make_io(c)
    # the point of the test: this batch is served by the half-line kernels without being asked to
    layer = runners.make_cuda_layer(c, params)
    d = layer._config().desc(c.B, 0)
    assert P._cabi.lib().pde_adi_checkpoint_bytes(byref(d)) > 0, "expected the half-line (checkpointing) kernels"
    got = runners.run_cuda(c, params=params, io=io)
    nt = os.cpu_count() or 1
    o32 = runners.run_oracle(c, params=params, io=io, dtype=np.float32, nthreads=nt)
    o64 = runners.run_oracle(c, params=params, io=io, dtype=np.float64, nthreads=nt)
    errs = _check(c, got, o32, o64)
    print(f"{c.name}: worst rel-err {max(errs.values()):.2e} ({max(errs, key=errs.get)})")


def test_default_dispatch_without_grad_input_at_large_batch():
    """What training needs: coefficient gradients only (the layer is the first op of every model)."""
    c = K.case("large_cifar10_pde3_nogin", "cifar10", B=16384, **K.SCRIPT_INSTANCES["cifar10_pde3"])
    params, io = K.make_params(c), K.make_io(c)
    got = runners.run_cuda(c, params=params, io=io, need_gin=False)
    assert got["gin"] is None
    nt = os.cpu_count() or 1
    o32 = runners.run_oracle(c, params=params, io=io, dtype=np.float32, need_gin=False, nthreads=nt)
    errs = runners.compare({k: v for k, v in got.items() if k != "gin"}, {k: v for k, v in o32.items() if k != "gin"})
    bad = {k: e for k, e in errs.items() if not e <= TOL}
    assert not bad, bad


@pytest.mark.parametrize("kind,B", [("fashion", 9473), ("mnist", 9475)], ids=lambda v: str(v))
def test_whole_line_inference_forward_two_pairs_per_warp(kind, B):
    """Under no_grad a large single-channel batch takes fwd_kernel<N, 2> (two sample pairs per warp);
    ragged B % 4 != 0 exercises its tail.  Forced as well, so that the test does not depend on the SM
    count of the device."""
    import torch
    c = K.case(f"np2_{kind}", kind, B=B)
    params, io = K.make_params(c), K.make_io(c)
    want = runners.run_oracle(c, params=params, io=io, dtype=np.float32, nthreads=os.cpu_count() or 1, forward_only=True)
    layer = runners.make_cuda_layer(c, params)
    x = torch.from_numpy(io[0]).cuda()
    for forced in ("", "2", "1"):
        if forced:
            os.environ["PDE_B200_FWD_NP"] = forced
        try:
            with torch.no_grad():
                y = layer(x)
        finally:
            os.environ.pop("PDE_B200_FWD_NP", None)
        err = max(runners.rel_l2(y.cpu().numpy(), want["y"]), runners.rel_max(y.cpu().numpy(), want["y"]))
        assert err <= TOL, (kind, forced, err)


def test_empty_batch_is_a_no_op():
    """parallel.shard_bounds hands empty shards to ranks beyond the batch; the reference modules accept
    B = 0 (they return an empty tensor and zero gradients)."""
    import torch
    for c in (K.case("empty_fashion", "fashion", B=0), K.case("empty_cifar10", "cifar10", B=0, **K.SCRIPT_INSTANCES["cifar10_pde3"]),
              K.case("empty_svhn", "svhn", B=0, **K.SCRIPT_INSTANCES["svhn"]), K.case("empty_emotion", "emotion", B=0),
              K.case("empty_tiny", "tiny", B=0, **K.SCRIPT_INSTANCES["tiny"])):
        layer = runners.make_cuda_layer(c)
        x = torch.zeros(0, *c.shape, device="cuda", requires_grad=True)
        y = layer(x)
        assert y.shape == x.shape
        y.backward(torch.zeros_like(y))
        assert x.grad.shape == x.shape
        for n, p in layer.named_parameters():
            if p.grad is not None:
                assert torch.count_nonzero(p.grad).item() == 0, (c.name, n)
