"""Parity of the CUDA path (through the C ABI) against the reference fixtures and the oracle.

Bar (BASELINE.json north_star): rel-err <= 1e-5 in fp32 for the output, grad_input and every
parameter gradient -- measured both as rel-L2 and as max-abs / max-ref -- and no further from
the fp64 oracle than the fp32 reference path is, plus 1e-5.
"""
import os

import numpy as np
import pytest

from . import cases as K
from . import golden_io, runners

pytestmark = pytest.mark.gpu
TOL = 1e-5  # north_star tolerance for fp32


def _assert_close(got, want, tol, what):
    errs = runners.compare(got, want)
    assert set(errs) == {k for k, v in want.items() if v is not None}, (sorted(errs), sorted(want))
    bad = {k: e for k, e in errs.items() if not e <= tol}
    assert not bad, f"{what}: {bad}"
    return errs


@pytest.mark.parametrize("c", K.GOLDEN_CASES, ids=lambda c: c.name)
def test_cuda_matches_reference_fixture(c):
    params, io, ref = golden_io.load(c)
    got = runners.run_cuda(c, params=params, io=io)
    _assert_close(got, ref, TOL, c.name + " vs reference fixture")


@pytest.mark.parametrize("c", K.CONFIG_CASES, ids=lambda c: c.name)
def test_cuda_matches_oracle_at_config_shape(c):
    params, io = K.make_params(c), K.make_io(c)
    got = runners.run_cuda(c, params=params, io=io)
    o32 = runners.run_oracle(c, params=params, io=io, dtype=np.float32)
    o64 = runners.run_oracle(c, params=params, io=io, dtype=np.float64)
    _assert_close(got, o32, TOL, c.name + " vs oracle fp32")
    e_cuda = runners.compare(got, o64)
    e_ref = runners.compare(o32, o64)
    bad = {k: (e_cuda[k], e_ref[k]) for k in e_cuda if not e_cuda[k] <= e_ref[k] + TOL}
    assert not bad, f"{c.name}: further from fp64 than the fp32 oracle + 1e-5: {bad}"


@pytest.mark.parametrize("c", [c for c in K.CONFIG_CASES if c.kind in ("mnist", "fashion", "svhn", "cifar10", "cifar2")],
                         ids=lambda c: c.name)
def test_cuda_whole_line_kernels_at_config_shape(monkeypatch, c):
    """Training calls on 28 x 28 / 32 x 32 planes default to the half-line kernels (adi_split.cu) at every
    batch size; the whole-line kernels of adi.cu (other plane sizes, four channels, inference) must stay
    right on the same shapes."""
    monkeypatch.setenv("PDE_B200_ADI_LEGACY", "1")
    params, io = K.make_params(c), K.make_io(c)
    got = runners.run_cuda(c, params=params, io=io)
    o32 = runners.run_oracle(c, params=params, io=io, dtype=np.float32)
    _assert_close(got, o32, TOL, c.name + " (whole-line kernels) vs oracle fp32")


@pytest.mark.parametrize("c", [K.CONFIG_CASES[1], K.CONFIG_CASES[3], K.CONFIG_CASES[7], K.CONFIG_CASES[9]],
                         ids=lambda c: c.name)
def test_cuda_without_grad_input(c):
    """The layer is the first op of every reference model: grad_input is normally not needed."""
    params, io = K.make_params(c), K.make_io(c)
    got = runners.run_cuda(c, params=params, io=io, need_gin=False)
    o32 = runners.run_oracle(c, params=params, io=io, dtype=np.float32, need_gin=False)
    assert got["gin"] is None
    _assert_close(got, o32, TOL, c.name)


def test_cuda_edge_cases():
    # batch of one (mnist_test.py:420 calls the layer on images[i:i+1]) and odd batches
    for B in (1, 3, 5):
        for kind, ctor in (("mnist", {}), ("cifar10", K.SCRIPT_INSTANCES["cifar10_pde3"]),
                           ("svhn", dict(size=16, channels=3, num_steps=3)), ("emotion", {}),
                           ("tiny", dict(size=16, channels=3, num_steps=2))):
            c = K.case(f"edge_{kind}_b{B}", kind, B=B, **ctor)
            params, io = K.make_params(c), K.make_io(c)
            got = runners.run_cuda(c, params=params, io=io)
            want = runners.run_oracle(c, params=params, io=io, dtype=np.float32)
            _assert_close(got, want, TOL, c.name)
    # zero steps: identity
    c0 = K.case("edge_steps0", "cifar10", B=2, size=16, channels=3, num_steps=0)
    got = runners.run_cuda(c0)
    u, g = K.make_io(c0)
    np.testing.assert_array_equal(got["y"], u)
    np.testing.assert_array_equal(got["gin"], g)


@pytest.mark.parametrize("ctor", [dict(Nx=16, Ny=16), dict(Nx=32, Ny=32, T=0.004), dict(Nx=48, Ny=48, T=0.003),
                                  dict(Nx=24, Ny=24), dict(Nx=64, Ny=64, T=0.002)],
                         ids=lambda d: f"N{d['Nx']}")
def test_cuda_emotion_plane_sizes(ctor):
    """Plane edges 16 / 32 / 48 take the register-tiled kernels (a warp per plane), the others the
    generic shared-memory kernels; odd batch, stable coefficients and the unstable default ones."""
    for perturb, B in ((False, 5), (True, 7)):
        c = K.case(f"emo_{ctor['Nx']}_{int(perturb)}", "emotion", B=B, perturb=perturb, **ctor)
        params, io = K.make_params(c), K.make_io(c)
        got = runners.run_cuda(c, params=params, io=io)
        want = runners.run_oracle(c, params=params, io=io, dtype=np.float32)
        _assert_close(got, want, TOL, c.name)
        got = runners.run_cuda(c, params=params, io=io, need_gin=False)
        assert got["gin"] is None
        _assert_close({k: v for k, v in got.items() if k != "gin"},
                      {k: v for k, v in want.items() if k != "gin"}, TOL, c.name + " (no grad_input)")


@pytest.mark.parametrize("size", [2, 5, 7, 30, 36, 48, 64, 96, 128])
def test_cuda_generic_plane_sizes(size):
    """Every plane edge without kernels of its own (odd ones included, up to 128 while a sample's planes fit one
    block's shared memory) is served by adi_generic.cu: all four implicit variants, odd batches, perturbed
    weights with clamped cells, with and without grad_input."""
    todo = [K.case(f"gsize{size}_mnist", "mnist", B=5, size=size, num_steps=3, dt=0.05, dx=0.7, dy=1.3)]
    if size <= 96:
        todo += [K.case(f"gsize{size}_svhn", "svhn", B=3, size=size, channels=3, num_steps=2),
                 K.case(f"gsize{size}_cifar10_c4", "cifar10", B=3, size=size, channels=4, dt=0.01, num_steps=2, dx=1.0, dy=1.5),
                 K.case(f"gsize{size}_cifar2_c2", "cifar2", B=7, size=size, channels=2, dt=0.02, num_steps=3)]
    else:
        todo += [K.case(f"gsize{size}_cifar10_c3", "cifar10", B=2, size=size, channels=3, dt=0.01, num_steps=2)]
    for c in todo:
        params, io = K.make_params(c), K.make_io(c)
        got = runners.run_cuda(c, params=params, io=io)
        want = runners.run_oracle(c, params=params, io=io, dtype=np.float32, nthreads=os.cpu_count() or 1)
        _assert_close(got, want, TOL, c.name)
    c = todo[-1]
    params, io = K.make_params(c), K.make_io(c)
    got = runners.run_cuda(c, params=params, io=io, need_gin=False)
    want = runners.run_oracle(c, params=params, io=io, dtype=np.float32, need_gin=False, nthreads=os.cpu_count() or 1)
    assert got["gin"] is None
    _assert_close({k: v for k, v in got.items() if k != "gin"}, {k: v for k, v in want.items() if k != "gin"}, TOL, c.name)


def test_cuda_generic_plane_size_large_batch_and_inference():
    """More samples than blocks (every block walks several samples and sums their gradients in its own
    accumulators), fp32 and fp64 oracle; and the same layer under no_grad."""
    import torch
    for c in (K.case("gsize36_large", "mnist", B=6001, size=36, num_steps=3, dt=0.05),
              K.case("gsize40_large_svhn", "svhn", B=1203, size=40, channels=3, num_steps=2)):
        params, io = K.make_params(c), K.make_io(c)
        nt = os.cpu_count() or 1
        got = runners.run_cuda(c, params=params, io=io)
        o32 = runners.run_oracle(c, params=params, io=io, dtype=np.float32, nthreads=nt)
        o64 = runners.run_oracle(c, params=params, io=io, dtype=np.float64, nthreads=nt)
        _assert_close(got, o32, TOL, c.name)
        e_cuda, e_ref = runners.compare(got, o64), runners.compare(o32, o64)
        bad = {k: (e_cuda[k], e_ref[k]) for k in e_cuda if not e_cuda[k] <= e_ref[k] + TOL}
        assert not bad, bad
        layer = runners.make_cuda_layer(c, params)
        with torch.no_grad():
            y = layer(torch.from_numpy(io[0]).cuda())
        np.testing.assert_array_equal(y.cpu().numpy(), got["y"])


def _assert_close_or_as_good_as_fp32(c, got, want32, params, io, need_gin, what):
    """<= 1e-5 against the fp32 oracle, or -- for outputs that are ill-conditioned sums (a 1 x 1 channel-matrix
    gradient is one number out of 10^4 cancelling products: the fp32 and fp64 oracles themselves differ by 1e-4
    there, and the skip-weight gradient is the same kind of sum) -- within three times the fp32 oracle's own distance
    from the fp64 oracle, plus 1e-5: the kernels add fp32 summation noise to the trajectory rounding both share."""
    errs = runners.compare(got, want32)
    assert set(errs) == {k for k, v in want32.items() if v is not None}, (sorted(errs), sorted(want32))
    bad = {k: e for k, e in errs.items() if not e <= TOL}
    if not bad:
        return
    want64 = runners.run_oracle(c, params=params, io=io, dtype=np.float64, need_gin=need_gin)
    e_cuda, e_ref = runners.compare(got, want64), runners.compare(want32, want64)
    worse = {k: (e_cuda[k], e_ref[k]) for k in bad if not e_cuda[k] <= 3.0 * e_ref[k] + TOL}
    assert not worse, f"{what}: {worse}"


def test_cuda_generic_random_configurations():
    """Seeded sweep over what adi_generic.cu claims to serve: plane edge 2 ... 64 (any parity), one to four
    channels, all four layer variants, zero to three steps, batches from one sample to more than a block's
    slots, with and without grad_input."""
    rs = np.random.RandomState(20261019)
    specialised = {8, 12, 16, 20, 24, 28, 32}
    done = 0
    while done < 40:
        size = int(rs.randint(2, 65))
        if size in specialised:
            continue
        kind = ["mnist", "svhn", "cifar10", "cifar2"][rs.randint(4)]
        steps, B = int(rs.randint(0, 4)), int(rs.choice([1, 2, 3, 5, 9, 17]))
        if kind == "mnist":
            ctor = dict(size=size, num_steps=steps, dt=float(rs.choice([0.01, 0.05, 0.3])), dx=float(rs.choice([0.7, 1.0])),
                        dy=float(rs.choice([1.0, 1.3])))
        elif kind == "svhn":
            ctor = dict(size=size, channels=int(rs.randint(1, 5)), num_steps=steps, dt=float(rs.choice([0.01, 0.05])))
        else:
            ctor = dict(size=size, channels=int(rs.randint(1, 5)), num_steps=steps, dt=float(rs.choice([0.001, 0.02])),
                        dx=float(rs.choice([1.0, 2.0])), dy=float(rs.choice([1.0, 1.5])))
        need_gin = bool(rs.randint(2))
        c = K.case(f"grand_{done}_{kind}_{size}", kind, B=B, seed=int(rs.randint(1 << 20)), **ctor)
        params, io = K.make_params(c), K.make_io(c)
        got = runners.run_cuda(c, params=params, io=io, need_gin=need_gin)
        want = runners.run_oracle(c, params=params, io=io, dtype=np.float32, need_gin=need_gin)
        if not need_gin:
            assert got["gin"] is None
            got, want = ({k: v for k, v in d.items() if k != "gin"} for d in (got, want))
        _assert_close_or_as_good_as_fp32(c, got, want, params, io, need_gin, f"{c.name} {ctor} B={B} gin={need_gin}")
        done += 1


def test_cuda_compiled_sizes_random_configurations(monkeypatch):
    """The same kind of seeded sweep over the plane edges with kernels of their own (8 ... 32): one to four
    channels, all four variants, zero to four steps, ragged batches, default dispatch or the whole-line kernels
    on request, with and without grad_input."""
    rs = np.random.RandomState(7919)
    for done in range(40):
        size = int(rs.choice([8, 12, 16, 20, 24, 28, 32]))
        kind = ["mnist", "svhn", "cifar10", "cifar2"][rs.randint(4)]
        steps, B = int(rs.randint(0, 5)), int(rs.choice([1, 2, 3, 5, 9, 17, 41, 70]))
        if kind == "mnist":
            ctor = dict(size=size, num_steps=steps, dt=float(rs.choice([0.01, 0.05, 0.3])), dx=float(rs.choice([0.7, 1.0])),
                        dy=float(rs.choice([1.0, 1.3])))
        elif kind == "svhn":
            ctor = dict(size=size, channels=int(rs.randint(1, 5)), num_steps=steps, dt=float(rs.choice([0.01, 0.05])))
        else:
            ctor = dict(size=size, channels=int(rs.randint(1, 5)), num_steps=steps, dt=float(rs.choice([0.001, 0.02])),
                        dx=float(rs.choice([1.0, 2.0])), dy=float(rs.choice([1.0, 1.5])))
        need_gin, whole = bool(rs.randint(2)), bool(rs.randint(3) == 0)
        c = K.case(f"crand_{done}_{kind}_{size}", kind, B=B, seed=int(rs.randint(1 << 20)), **ctor)
        params, io = K.make_params(c), K.make_io(c)
        if whole:
            monkeypatch.setenv("PDE_B200_ADI_LEGACY", "1")
        else:
            monkeypatch.delenv("PDE_B200_ADI_LEGACY", raising=False)
        got = runners.run_cuda(c, params=params, io=io, need_gin=need_gin)
        want = runners.run_oracle(c, params=params, io=io, dtype=np.float32, need_gin=need_gin)
        if not need_gin:
            assert got["gin"] is None
            got, want = ({k: v for k, v in d.items() if k != "gin"} for d in (got, want))
        _assert_close_or_as_good_as_fp32(c, got, want, params, io, need_gin,
                                         f"{c.name} {ctor} B={B} gin={need_gin} whole-line={whole}")


def test_cuda_explicit_layers_random_configurations():
    """Seeded sweep over the two explicit layers: emotion with any plane edge up to 64 (16 / 32 / 48 take the
    register-tiled kernels, the rest the shared-memory ones) and 1 ... 12 steps; tiny with every edge that is a
    multiple of 4 up to 64, one to four channels, one to three steps; ragged batches, grad_input on / off."""
    rs = np.random.RandomState(104729)
    for done in range(30):
        need_gin = bool(rs.randint(2))
        B = int(rs.choice([1, 2, 3, 5, 9, 17, 33]))
        if rs.randint(2):
            n = int(rs.choice([16, 32, 48])) if rs.randint(3) == 0 else int(rs.randint(4, 65))
            ctor = dict(Nx=n, Ny=n, T=float(rs.choice([0.001, 0.003, 0.005, 0.01, 0.012])))
            c = K.case(f"erand_{done}_emotion_{n}", "emotion", B=B, seed=int(rs.randint(1 << 20)), **ctor)
        else:
            n = 4 * int(rs.randint(2, 17))
            ctor = dict(size=n, channels=int(rs.randint(1, 5)), num_steps=int(rs.randint(1, 4)), dt=float(rs.choice([0.01, 0.02])))
            c = K.case(f"erand_{done}_tiny_{n}", "tiny", B=B, seed=int(rs.randint(1 << 20)), **ctor)
        params, io = K.make_params(c), K.make_io(c)
        got = runners.run_cuda(c, params=params, io=io, need_gin=need_gin)
        want = runners.run_oracle(c, params=params, io=io, dtype=np.float32, need_gin=need_gin)
        if not need_gin:
            assert got["gin"] is None
            got, want = ({k: v for k, v in d.items() if k != "gin"} for d in (got, want))
        _assert_close_or_as_good_as_fp32(c, got, want, params, io, need_gin, f"{c.name} {ctor} B={B} gin={need_gin}")


@pytest.mark.parametrize("size", [8, 12, 16, 20, 24])
def test_cuda_other_plane_sizes(size):
    """The reference classes take any `size`; the whole-line kernels are built for every multiple of 4 up to
    32 (28 and 32 are also served by the half-line kernels).  One, three and four channels, odd batches."""
    todo = [K.case(f"size{size}_mnist", "mnist", B=5, size=size, num_steps=3, dt=0.05, dx=0.7, dy=1.3),
            K.case(f"size{size}_svhn", "svhn", B=3, size=size, channels=3, num_steps=2),
            K.case(f"size{size}_cifar10_c4", "cifar10", B=3, size=size, channels=4, dt=0.01, num_steps=2, dx=1.0, dy=1.5),
            K.case(f"size{size}_cifar2_c2", "cifar2", B=7, size=size, channels=2, dt=0.02, num_steps=3)]
    for c in todo:
        params, io = K.make_params(c), K.make_io(c)
        got = runners.run_cuda(c, params=params, io=io)
        want = runners.run_oracle(c, params=params, io=io, dtype=np.float32)
        _assert_close(got, want, TOL, c.name)


def test_cuda_exact_mode_for_large_coefficients():
    """dt large enough that rebuilding sweep inputs would amplify rounding noise: the kernel must
    switch to per-sweep checkpoints on its own (DESIGN.md 'reverse reconstruction')."""
    c = K.case("fashion_dt5", "fashion", B=8, perturb=False, dt=5.0)
    params, io = K.make_params(c), K.make_io(c)
    got = runners.run_cuda(c, params=params, io=io)
    o32 = runners.run_oracle(c, params=params, io=io, dtype=np.float32)
    _assert_close(got, o32, TOL, c.name)


def test_cuda_large_batch_properties():
    """At a batch the oracle would take long for: linearity in u and the adjoint identity
    <J v, w> = <v, J^T w> between our forward and backward kernels."""
    import torch
    c = K.case("prop", "cifar10", B=4096, **K.SCRIPT_INSTANCES["cifar10_pde1"])
    layer = runners.make_cuda_layer(c)
    gen = torch.Generator(device="cuda").manual_seed(7)
    v = torch.randn(c.B, *c.shape, device="cuda", generator=gen)
    w = torch.randn(c.B, *c.shape, device="cuda", generator=gen)
    x = v.clone().requires_grad_(True)
    y = layer(x)
    (gin,) = torch.autograd.grad(y, x, w)
    lhs = (y.double() * w.double()).sum().item()
    rhs = (v.double() * gin.double()).sum().item()
    assert abs(lhs - rhs) <= 1e-5 * max(abs(lhs), abs(rhs), 1.0), (lhs, rhs)
    with torch.no_grad():
        y2 = layer(2.5 * v)
    assert runners.rel_l2(y2.cpu().numpy(), 2.5 * y.detach().cpu().numpy()) <= 1e-6


def test_cuda_layers_are_once_differentiable():
    """The backward passes call raw-pointer kernels: asking autograd for a graph through them
    (create_graph=True: gradient penalties, higher-order gradients) must raise, not return constants."""
    import torch
    for c in (K.case("once_fashion", "fashion", B=2), K.case("once_emotion", "emotion", B=2),
              K.case("once_tiny", "tiny", B=2, **K.SCRIPT_INSTANCES["tiny"])):
        layer = runners.make_cuda_layer(c)
        x = torch.randn(c.B, *c.shape, device="cuda", requires_grad=True)
        y = layer(x)
        with pytest.raises(RuntimeError, match="once differentiable"):
            torch.autograd.grad(y.sum(), x, create_graph=True)
        (g,) = torch.autograd.grad(layer(x).sum(), x)      # the plain first derivative is fine
        assert torch.isfinite(g).all()


def test_cuda_inference_table_cache_is_safe():
    """Under no_grad the coefficient tables are reused while the parameters stand.  They must be rebuilt when
    a parameter is updated in place (optimizer step, load_state_dict), and a new layer whose parameters land
    on the addresses of a dead one must not inherit its tables."""
    import gc
    import torch
    c = K.case("cache_fashion", "fashion", B=33)
    u = torch.from_numpy(K.make_io(c)[0]).cuda()

    def same_as_uncached(layer, y):
        # the grad-mode path never consults the cache (it runs the half-line kernels: equal up to rounding;
        # stale tables would be off by percents)
        ref = layer(u.clone().requires_grad_(True)).detach()
        return runners.rel_l2(y.cpu().numpy(), ref.cpu().numpy()) <= 1e-5

    layer = runners.make_cuda_layer(c, K.make_params(c))
    with torch.no_grad():
        y1, y1b = layer(u), layer(u)                      # second call: cached tables
    assert torch.equal(y1, y1b) and same_as_uncached(layer, y1)
    with torch.no_grad():
        layer.alpha_base.mul_(1.5)                        # in-place update: version bump
        y2 = layer(u)
    assert runners.rel_l2(y2.cpu().numpy(), y1.cpu().numpy()) > 1e-3 and same_as_uncached(layer, y2)
    with torch.no_grad():
        layer.load_state_dict({k: v * 0.5 for k, v in layer.state_dict().items()})
        y3 = layer(u)
    assert same_as_uncached(layer, y3)
    for seed in range(4):                                 # dead layers' addresses get reused by fresh ones
        del layer
        gc.collect()
        c2 = K.case("cache_fashion", "fashion", B=33, seed=100 + seed)
        layer = runners.make_cuda_layer(c2, K.make_params(c2))
        with torch.no_grad():
            y = layer(u)
        assert same_as_uncached(layer, y), seed


def test_cuda_rejects_cpu_tensors_and_bad_shapes():
    import torch
    from cnn_with_pde_b200.mnist_test import DiffusionLayer
    layer = DiffusionLayer().cuda()
    with pytest.raises(RuntimeError):
        layer(torch.zeros(2, 1, 28, 28))
    with pytest.raises(ValueError):
        layer(torch.zeros(2, 3, 28, 28, device="cuda"))


# ------------------------------------------------------------------ half-line (split) kernels
_SPLIT_CASES = [
    K.case("split_fashion", "fashion", B=9),
    K.case("split_mnist", "mnist", B=5),
    K.case("split_cifar10_pde2", "cifar10", B=3, **K.SCRIPT_INSTANCES["cifar10_pde2"]),
    K.case("split_cifar2", "cifar2", B=5, **K.SCRIPT_INSTANCES["cifar2_diffusion1"]),
    K.case("split_svhn", "svhn", B=5, **K.SCRIPT_INSTANCES["svhn"]),
]


@pytest.mark.parametrize("pairs,qf", [("2", "1"), ("2", "2"), ("4", "4"), ("4", "1")], ids=lambda v: str(v))
def test_cuda_split_kernel_variants(monkeypatch, pairs, qf):
    """28x28 and 32x32 planes take the half-line kernels (adi_split.cu); every pairs-per-group /
    groups-per-block instantiation must agree with the oracle (the defaults depend on the batch)."""
    monkeypatch.setenv("PDE_B200_SPLIT_P", pairs)
    monkeypatch.setenv("PDE_B200_SPLIT_QF", qf)
    for c in _SPLIT_CASES:
        params, io = K.make_params(c), K.make_io(c)
        want = runners.run_oracle(c, params=params, io=io, dtype=np.float32)
        got = runners.run_cuda(c, params=params, io=io)
        _assert_close(got, want, TOL, f"{c.name} P={pairs} Qf={qf}")


@pytest.mark.parametrize("env", ["PDE_B200_NO_CKPT", "PDE_B200_ADI_LEGACY"])
def test_cuda_backward_without_saved_checkpoints(monkeypatch, env):
    """pde_adi_backward without checkpoints from the forward call (it makes them in its workspace),
    and the whole-line kernels of adi.cu on the sizes the half-line kernels normally serve."""
    monkeypatch.setenv(env, "1")
    for c in _SPLIT_CASES + [K.case("split_fashion_dt5", "fashion", B=8, perturb=False, dt=5.0)]:
        params, io = K.make_params(c), K.make_io(c)
        want = runners.run_oracle(c, params=params, io=io, dtype=np.float32)
        got = runners.run_cuda(c, params=params, io=io)
        _assert_close(got, want, TOL, f"{c.name} {env}")


def test_cuda_split_large_batch_properties():
    """Batch large enough for four pairs per group and several groups per block: adjoint identity
    <J v, w> = <v, J^T w> and linearity, fashion layer (smoothing, dt = 0.3)."""
    import torch
    c = K.case("prop_fashion", "fashion", B=8200)
    layer = runners.make_cuda_layer(c)
    gen = torch.Generator(device="cuda").manual_seed(11)
    v = torch.randn(c.B, *c.shape, device="cuda", generator=gen)
    w = torch.randn(c.B, *c.shape, device="cuda", generator=gen)
    x = v.clone().requires_grad_(True)
    y = layer(x)
    (gin,) = torch.autograd.grad(y, x, w)
    lhs = (y.double() * w.double()).sum().item()
    rhs = (v.double() * gin.double()).sum().item()
    assert abs(lhs - rhs) <= 1e-5 * max(abs(lhs), abs(rhs), 1.0), (lhs, rhs)
    with torch.no_grad():
        y_eval = layer(v)            # inference takes the whole-line forward kernel
        y2 = layer(2.5 * v)
    assert runners.rel_l2(y_eval.cpu().numpy(), y.detach().cpu().numpy()) <= 2e-6
    assert runners.rel_l2(y2.cpu().numpy(), 2.5 * y.detach().cpu().numpy()) <= 2e-6
    # the first and the last sample of the batch against the oracle (ragged last group: 8200 = 1025 * 8)
    params = K.make_params(c)
    for sl in (slice(0, 2), slice(c.B - 2, c.B)):
        c2 = K.case("prop_fashion_2", "fashion", B=2)
        want = runners.run_oracle(c2, params=params, io=(v[sl].cpu().numpy(), w[sl].cpu().numpy()), dtype=np.float32,
                                  forward_only=True)
        assert runners.rel_l2(y[sl].detach().cpu().numpy(), want["y"]) <= TOL


@pytest.mark.parametrize("c", [c for c in K.GOLDEN_CASES if c.kind in ("mnist", "fashion", "svhn", "cifar10", "cifar2")],
                         ids=lambda c: c.name)
def test_cuda_split_matches_reference_fixture(monkeypatch, c):
    """The half-line kernels against the fixtures generated from the unmodified reference."""
    monkeypatch.setenv("PDE_B200_ADI_SPLIT", "1")
    params, io, ref = golden_io.load(c)
    got = runners.run_cuda(c, params=params, io=io)
    _assert_close(got, ref, TOL, c.name + " (half-line kernels) vs reference fixture")


def test_cuda_split_shapes_and_tails(monkeypatch):
    """Half-line kernels off the beaten path: two channels, 28 x 28 with three channels, batches that
    leave the last group ragged, no grad_input, zero time coefficients crossing a clamp."""
    monkeypatch.setenv("PDE_B200_ADI_SPLIT", "1")
    todo = [
        (K.case("split_c2", "cifar10", B=7, size=32, channels=2, dt=0.002, num_steps=3, dx=1.0, dy=1.0), True),
        (K.case("split_svhn28", "svhn", B=5, size=28, channels=3, num_steps=3), True),
        (K.case("split_cifar2_c1", "cifar2", B=11, size=32, channels=1, dt=0.002, num_steps=4), True),
        (K.case("split_fashion_nogin", "fashion", B=13), False),
        (K.case("split_cifar10_nogin", "cifar10", B=6, **K.SCRIPT_INSTANCES["cifar10_pde3"]), False),
    ]
    for forced in ("", "4"):
        if forced:
            monkeypatch.setenv("PDE_B200_SPLIT_P", forced)
        for c, need_gin in todo:
            params, io = K.make_params(c), K.make_io(c)
            want = runners.run_oracle(c, params=params, io=io, dtype=np.float32, need_gin=need_gin)
            got = runners.run_cuda(c, params=params, io=io, need_gin=need_gin)
            if not need_gin:
                assert got["gin"] is None
            _assert_close(got, want, TOL, f"{c.name} P={forced or 'default'}")


def test_cuda_layers_under_autocast():
    """The CIFAR scripts wrap the model in autocast (cifar10.py:459, cifar_2version.py:521): the
    layers cast their inputs to fp32 (custom_fwd) and give the fp32 result; a half-precision input is
    accepted and its gradient comes back in its own dtype."""
    import torch
    for c in (K.case("amp_cifar10", "cifar10", B=3, **K.SCRIPT_INSTANCES["cifar10_pde3"]), K.case("amp_emotion", "emotion", B=2),
              K.case("amp_tiny", "tiny", B=2, **K.SCRIPT_INSTANCES["tiny"])):
        layer = runners.make_cuda_layer(c)
        u, g = K.make_io(c)
        x = torch.from_numpy(u).cuda().requires_grad_(True)
        y_ref = layer(x)
        y_ref.backward(torch.from_numpy(g).cuda())
        ref_grads = [p.grad.clone() for p in layer.parameters() if p.grad is not None]
        gin_ref = x.grad.clone()
        for p in layer.parameters():
            p.grad = None
        xh = torch.from_numpy(u).cuda().half().requires_grad_(True)
        with torch.autocast("cuda", dtype=torch.float16):
            y = layer(xh)
        assert y.dtype == torch.float32
        y.backward(torch.from_numpy(g).cuda())
        assert xh.grad.dtype == torch.float16
        # the only difference is the rounding of the input to fp16 (2^-11 relative)
        assert runners.rel_l2(y.detach().cpu().numpy(), y_ref.detach().cpu().numpy()) <= 2e-3
        assert runners.rel_l2(xh.grad.float().cpu().numpy(), gin_ref.cpu().numpy()) <= 2e-3
        got = [p.grad for p in layer.parameters() if p.grad is not None]
        assert len(got) == len(ref_grads) and all(torch.isfinite(a).all() for a in got)
        # fp32 input under autocast: bit-identical to the plain call
        for p in layer.parameters():
            p.grad = None
        x2 = torch.from_numpy(u).cuda().requires_grad_(True)
        with torch.autocast("cuda", dtype=torch.bfloat16):
            y2 = layer(x2)
        assert torch.equal(y2, y_ref)


@pytest.mark.parametrize("name,kind,key", [("autocast_cifar10_pde3", "cifar10", "cifar10_pde3"),
                                           ("autocast_cifar2_diffusion2", "cifar2", "cifar2_diffusion2")])
def test_cuda_layer_in_the_scripts_autocast_loop_against_the_reference_run_the_same_way(name, kind, key):
    """The CIFAR scripts call the model inside torch.autocast (cifar10.py:459, cifar_2version.py:521).
    Fixture (tests/golden/make_golden_autocast.py): the unmodified reference layer run inside autocast and
    run plainly.  Its autocast run is 4e-3 ... 9e-3 away from its own fp32 run (the channel mix drops to
    reduced precision, cifar10.py:71).  Our layer, called inside CUDA autocast, must (a) reproduce the
    reference's fp32 numbers at the usual 1e-5 and (b) therefore sit as close to the reference's autocast
    numbers as the reference's own two runs sit to each other."""
    import os
    import torch
    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "autocast_cifar.npz"))
    c = K.case(name, kind, B=4, **K.SCRIPT_INSTANCES[key])
    params, (u, g) = K.make_params(c), K.make_io(c)
    for dtype in (torch.float16, torch.bfloat16):
        layer = runners.make_cuda_layer(c, params)
        x = torch.from_numpy(u).cuda().requires_grad_(True)
        with torch.autocast("cuda", dtype=dtype):
            y = layer(x)
            assert y.dtype == torch.float32
        y.backward(torch.from_numpy(g).cuda())
        got = {"y": y.detach().cpu().numpy(), "gin": x.grad.cpu().numpy()}
        got.update({"g_" + k: p.grad.cpu().numpy() for k, p in layer.named_parameters()})
        for k, v in got.items():
            fp32, amp = z[f"{name}/fp32/{k}"], z[f"{name}/amp/{k}"]
            e_fp32 = max(runners.rel_l2(v, fp32), runners.rel_max(v, fp32))
            assert e_fp32 <= TOL, (k, dtype, e_fp32)
            gap = runners.rel_l2(amp, fp32)                    # the reference against itself
            assert runners.rel_l2(v, amp) <= 1.05 * gap + TOL, (k, dtype, runners.rel_l2(v, amp), gap)


def test_cuda_split_few_steps_and_single_channel_ops(monkeypatch):
    """Half-line kernels at the ends of the schedule: one and two steps (the checkpoint stream runs
    one step ahead, across items), Lie splitting, channel ops with a single channel, batches of
    several items per block."""
    monkeypatch.setenv("PDE_B200_ADI_SPLIT", "1")
    todo = [
        K.case("split_fashion_1step", "fashion", B=37, num_steps=1),
        K.case("split_mnist_2steps", "mnist", B=41, num_steps=2),
        K.case("split_cifar2_1step", "cifar2", B=9, size=32, channels=3, dt=0.002, num_steps=1),
        K.case("split_svhn_c1", "svhn", B=6, size=32, channels=1, num_steps=2),
        K.case("split_cifar10_c1", "cifar10", B=6, size=28, channels=1, dt=0.002, num_steps=2, dx=1.0, dy=1.0),
    ]
    for c in todo:
        params, io = K.make_params(c), K.make_io(c)
        want = runners.run_oracle(c, params=params, io=io, dtype=np.float32)
        got = runners.run_cuda(c, params=params, io=io)
        _assert_close(got, want, TOL, c.name)
    # many items per block: 3000 samples on a persistent grid, first / last samples against the oracle
    import torch
    c = K.case("split_many_items", "cifar10", B=3000, **K.SCRIPT_INSTANCES["cifar10_pde3"])
    params = K.make_params(c)
    layer = runners.make_cuda_layer(c, params)
    gen = torch.Generator(device="cuda").manual_seed(3)
    u = torch.randn(c.B, *c.shape, device="cuda", generator=gen)
    g = torch.randn(c.B, *c.shape, device="cuda", generator=gen)
    x = u.clone().requires_grad_(True)
    y = layer(x)
    y.backward(g)
    for sl in (slice(0, 3), slice(c.B - 3, c.B)):
        c3 = K.case("split_many_items_3", "cifar10", B=3, **K.SCRIPT_INSTANCES["cifar10_pde3"])
        want = runners.run_oracle(c3, params=params, io=(u[sl].cpu().numpy(), g[sl].cpu().numpy()), dtype=np.float32)
        assert runners.rel_l2(y[sl].detach().cpu().numpy(), want["y"]) <= TOL
        assert runners.rel_l2(x.grad[sl].cpu().numpy(), want["gin"]) <= TOL


def test_cuda_split_results_are_bitwise_reproducible(monkeypatch):
    """The half-line kernels synchronise warps of a block through barriers, mbarriers (TMA) and
    cp.async groups, and sum gradients in a fixed order: repeated calls must agree bit for bit
    (a missing barrier shows up as run-to-run differences long before it shows up in a tolerance)."""
    import torch
    monkeypatch.setenv("PDE_B200_ADI_SPLIT", "1")
    for c in (K.case("rep_fashion", "fashion", B=1201), K.case("rep_svhn", "svhn", B=301, **K.SCRIPT_INSTANCES["svhn"]),
              K.case("rep_cifar10", "cifar10", B=610, **K.SCRIPT_INSTANCES["cifar10_pde2"])):
        layer = runners.make_cuda_layer(c)
        gen = torch.Generator(device="cuda").manual_seed(17)
        u = torch.randn(c.B, *c.shape, device="cuda", generator=gen)
        g = torch.randn(c.B, *c.shape, device="cuda", generator=gen)
        first = None
        for _ in range(6):
            for p in layer.parameters():
                p.grad = None
            x = u.clone().requires_grad_(True)
            y = layer(x)
            y.backward(g)
            now = [y.detach().clone(), x.grad.clone()] + [p.grad.clone() for p in layer.parameters() if p.grad is not None]
            if first is None:
                first = now
            else:
                for a, b in zip(first, now):
                    assert torch.equal(a, b), c.name


# ------------------------------------------------------- several layers on the same input, one launch per pass
def _three_cifar_layers(B):
    cs = [K.case(f"multi_{k}", "cifar10", B=B, seed=1234 + i, **K.SCRIPT_INSTANCES[k])
          for i, k in enumerate(("cifar10_pde1", "cifar10_pde2", "cifar10_pde3"))]
    return cs, [K.make_params(c) for c in cs]


@pytest.mark.parametrize("B", [3, 512, 4099])
def test_cuda_multi_branch_launch_matches_single_layer_calls_and_oracle(B):
    """The three PDE layers of cifar10's MultiScaleExtractor (cifar10.py:253-258,272-274) through
    cifar10.apply_to_same_input -- one prepare, one forward, one backward and one finish launch for all of
    them -- against three single-layer calls (outputs and grad_input bit-identical) and against the oracle."""
    import torch
    from cnn_with_pde_b200.cifar10 import apply_to_same_input
    cs, ps = _three_cifar_layers(B)
    layers = [runners.make_cuda_layer(c, p) for c, p in zip(cs, ps)]
    u_np, _ = K.make_io(cs[0])
    gs_np = [K.make_io(c)[1] for c in cs]
    x = torch.from_numpy(u_np).cuda().requires_grad_(True)
    ys = apply_to_same_input(x, layers)
    torch.autograd.backward(list(ys), [torch.from_numpy(g).cuda() for g in gs_np])
    fused = {"gin": x.grad.clone(), "ys": [y.detach().clone() for y in ys],
             "grads": [{k: p.grad.clone() for k, p in l.named_parameters()} for l in layers]}
    x2 = torch.from_numpy(u_np).cuda().requires_grad_(True)
    for l in layers:
        for p in l.parameters():
            p.grad = None
    ys2 = [l(x2) for l in layers]
    torch.autograd.backward(ys2, [torch.from_numpy(g).cuda() for g in gs_np])
    for a, b in zip(fused["ys"], ys2):
        assert torch.equal(a, b)
    assert runners.rel_l2(fused["gin"].cpu().numpy(), x2.grad.cpu().numpy()) <= 1e-6
    for i, (c, p, l) in enumerate(zip(cs, ps, layers)):
        want = runners.run_oracle(c, params=p, io=(u_np, gs_np[i]), dtype=np.float32, nthreads=0)
        got = {"y": fused["ys"][i].cpu().numpy(), "gin": None}
        got.update({"g_" + k: v.cpu().numpy() for k, v in fused["grads"][i].items()})
        want = {k: v for k, v in want.items() if k != "gin"}
        _assert_close({k: v for k, v in got.items() if k != "gin"}, want, TOL, f"{c.name} fused B={B}")
        for k, pp in l.named_parameters():      # and against the single-layer calls, to summation-order noise
            assert runners.rel_l2(fused["grads"][i][k].cpu().numpy(), pp.grad.cpu().numpy()) <= 2e-6, (c.name, k)


def test_cuda_multi_branch_falls_back_when_layers_cannot_share_a_launch():
    """Different plane sizes / inference / a single layer: apply_to_same_input runs the layers one by one."""
    import torch
    from cnn_with_pde_b200.cifar10 import EnhancedDiffusionLayer, apply_to_same_input
    torch.manual_seed(0)
    a = EnhancedDiffusionLayer(32, 3, dt=0.001, num_steps=2).cuda()
    b = EnhancedDiffusionLayer(32, 3, dt=0.002, num_steps=3).cuda()
    x = torch.randn(5, 3, 32, 32, device="cuda")
    with torch.no_grad():
        ya, yb = apply_to_same_input(x, [a, b])
        assert torch.equal(ya, a(x)) and torch.equal(yb, b(x))
    (only,) = apply_to_same_input(x, [a])
    assert torch.equal(only, a(x))
    c16 = EnhancedDiffusionLayer(16, 3, dt=0.001, num_steps=2).cuda()
    d16 = EnhancedDiffusionLayer(16, 3, dt=0.002, num_steps=2).cuda()
    x16 = torch.randn(4, 3, 16, 16, device="cuda", requires_grad=True)
    y1, y2 = apply_to_same_input(x16, [c16, d16])       # 16 x 16 planes: whole-line kernels, one call per layer
    (y1.sum() + y2.sum()).backward()
    assert x16.grad is not None and c16.alpha_base.grad is not None and d16.alpha_base.grad is not None


# ------------------------------------------------------------------------------ bf16 I/O (tiny_imagenet layer)
@pytest.mark.parametrize("ctor,B", [(K.SCRIPT_INSTANCES["tiny"], 5), (dict(size=16, channels=3, num_steps=3, dt=0.02), 7),
                                    (dict(size=32, channels=2, num_steps=2, dt=0.01), 3)], ids=["64x64", "16x16_3steps", "32x32_2steps"])
def test_cuda_tiny_layer_bf16_io(ctor, B):
    """bfloat16 planes in, bfloat16 planes out (pde_tiny_*_bf16): the arithmetic is the fp32 kernels', so on the
    bf16-rounded inputs the parameter gradients (fp32) match the oracle at 1e-5 and the output / grad_input are
    the oracle's fp32 results rounded to bf16 (to within one bf16 unit where the fp32 values straddle a tie)."""
    import torch
    c = K.case("tiny_bf16", "tiny", B=B, **ctor)
    params, (u, g) = K.make_params(c), K.make_io(c)
    ub, gb = torch.from_numpy(u).cuda().bfloat16(), torch.from_numpy(g).cuda().bfloat16()
    layer = runners.make_cuda_layer(c, params)
    x = ub.clone().requires_grad_(True)
    y = layer(x)
    assert y.dtype == torch.bfloat16
    y.backward(gb)
    assert x.grad.dtype == torch.bfloat16
    want = runners.run_oracle(c, params=params, io=(ub.float().cpu().numpy(), gb.float().cpu().numpy()), dtype=np.float32)
    for got, ref in ((y, want["y"]), (x.grad, want["gin"])):
        ref_b = torch.from_numpy(ref).bfloat16().float().numpy()
        got_f = got.detach().float().cpu().numpy()
        assert np.all(np.abs(got_f - ref_b) <= 2.0 ** -7 * np.abs(ref_b) + 1e-30)        # at most one bf16 unit
        assert np.mean(got_f != ref_b) <= 0.01                                            # and almost always none
    for k in ("alpha_base", "channel_scaling"):
        a, b = getattr(layer, k).grad.cpu().numpy(), want["g_" + k]
        assert max(runners.rel_l2(a, b), runners.rel_max(a, b)) <= TOL, k
    # fp32 input on the same layer still takes the fp32 kernels
    y32 = layer(torch.from_numpy(u).cuda())
    assert y32.dtype == torch.float32
