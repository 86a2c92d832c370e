"""The dormant scalar-coefficient methods of tiny_imagenet.ImprovedDiffusionLayer (tiny_imagenet.py:88-233):
oracle against fixtures made by the unmodified reference (CPU), CUDA against both (GPU)."""
import os

import numpy as np
import pytest

from . import cases as K
from . import runners

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "tiny_split.npz")
TOL = 1e-5


def _kw(method, args, dt_layer):
    if method == "implicit_diffusion_step":
        return dict(coeff_x=args[0], coeff_y=args[1], dt=dt_layer)
    if method == "solve_implicit_x":
        return dict(coeff_x=args[0], dt=args[1])
    if method == "solve_implicit_y":
        return dict(coeff_y=args[0], dt=args[1])
    if method == "diffuse_x_explicit":
        return dict(coeff_x=args[0], dt=dt_layer)
    return dict(coeff_y=args[0], dt=dt_layer)


@pytest.mark.parametrize("case", K.TINY_SPLIT_CASES, ids=lambda c: c[0])
def test_oracle_matches_reference_fixture(case):
    import oracle as O
    name, method, shape, dt_layer, args = case
    z = np.load(GOLDEN)
    got = O.tiny_split(method, z[name + "/u"], **_kw(method, args, dt_layer))
    # the restatement follows the reference operation by operation: bit-exact in fp32
    np.testing.assert_array_equal(got, z[name + "/ref_y"])
    # every one of these maps is self-adjoint: the gradient autograd gave the reference for the input is
    # the same map applied to the upstream gradient
    gin = O.tiny_split(method, z[name + "/gout"], **_kw(method, args, dt_layer))
    assert runners.rel_l2(gin, z[name + "/ref_gin"]) <= 2e-6
    # fp64 twin close to the fp32 reference
    got64 = O.tiny_split(method, z[name + "/u"].astype(np.float64), **_kw(method, args, dt_layer))
    assert runners.rel_l2(got64, z[name + "/ref_y"]) <= (2e-6 if "clamped" not in name else 1e-4)


@pytest.mark.gpu
@pytest.mark.parametrize("case", K.TINY_SPLIT_CASES, ids=lambda c: c[0])
def test_cuda_matches_reference_fixture_and_oracle(case):
    import torch
    import oracle as O
    from cnn_with_pde_b200.tiny_imagenet import ImprovedDiffusionLayer
    name, method, shape, dt_layer, args = case
    z = np.load(GOLDEN)
    layer = ImprovedDiffusionLayer(size=shape[1], channels=3, dt=dt_layer, num_steps=1, use_implicit=True).cuda()
    x = torch.from_numpy(z[name + "/u"]).cuda().requires_grad_(True)
    y = getattr(layer, method)(x, *args)
    y.backward(torch.from_numpy(z[name + "/gout"]).cuda())
    got_y, got_g = y.detach().cpu().numpy(), x.grad.cpu().numpy()
    # pivots at the clamp amplify every rounding by 1e6 per row (values reach 4e29): the reciprocal-multiply
    # form of the kernel is then compared at the accuracy the reference's own fp32-vs-fp64 gap allows
    tol = TOL if "clamped" not in name else 1e-4
    for what, got, want in (("y", got_y, z[name + "/ref_y"]), ("gin", got_g, z[name + "/ref_gin"])):
        err = max(runners.rel_l2(got, want), runners.rel_max(got, want))
        assert err <= tol, (name, what, err)
    want = O.tiny_split(method, z[name + "/u"], **_kw(method, args, dt_layer))
    assert max(runners.rel_l2(got_y, want), runners.rel_max(got_y, want)) <= tol
    # forward() still ignores use_implicit, as the reference does (tiny_imagenet.py:21,34-51)
    u4 = torch.randn(2, 3, 64, 64, device="cuda")
    plain = ImprovedDiffusionLayer(size=64, channels=3, dt=dt_layer, num_steps=1, use_implicit=False).cuda()
    assert torch.equal(layer(u4), plain(u4))


@pytest.mark.gpu
def test_cuda_tiny_split_large_batch_adjoint_identity():
    """<J v, w> = <v, J w> (self-adjoint maps) at a batch that fills the GPU several times over, plus a
    ragged last block and an empty batch."""
    import torch
    from cnn_with_pde_b200.tiny_imagenet import ImprovedDiffusionLayer
    layer = ImprovedDiffusionLayer(size=64, channels=3, dt=0.3).cuda()
    gen = torch.Generator(device="cuda").manual_seed(5)
    v = torch.randn(3 * 4097, 64, 64, device="cuda", generator=gen)
    w = torch.randn(3 * 4097, 64, 64, device="cuda", generator=gen)
    jv = layer.implicit_diffusion_step(v, 0.7, 1.9)
    jw = layer.implicit_diffusion_step(w, 0.7, 1.9)
    lhs, rhs = (jv.double() * w.double()).sum().item(), (v.double() * jw.double()).sum().item()
    assert abs(lhs - rhs) <= 1e-5 * max(abs(lhs), abs(rhs), 1.0), (lhs, rhs)
    import oracle as O
    want = O.tiny_split("implicit_diffusion_step", v[-3:].cpu().numpy(), coeff_x=0.7, coeff_y=1.9, dt=0.3)
    assert runners.rel_l2(jv[-3:].cpu().numpy(), want) <= TOL
    assert layer.implicit_diffusion_step(v[:0], 0.7, 1.9).shape == (0, 64, 64)
