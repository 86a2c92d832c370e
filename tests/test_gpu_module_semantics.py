"""What a user of the reference's nn.Modules relies on besides the numbers (`forward` row):
the input is not mutated, views / non-contiguous inputs are accepted, train() and eval() give the same
result, a second backward over a retained graph gives the same gradients (the saved coefficient tables and
step checkpoints are read-only for the backward kernels), gradients accumulate into `.grad`, frozen
parameters stay without a gradient, and several calls in flight (same layer applied twice before any
backward, as `hybrid_pde_regularization`-style code may do) do not share scratch.

Small odd batches of every layer family; comparisons are exact (same kernels, same inputs) unless stated.
"""
import numpy as np
import pytest
import torch

from . import cases as K
from . import runners

pytestmark = pytest.mark.gpu

_CASES = [
    K.case("sem_fashion", "fashion", B=37),
    K.case("sem_mnist", "mnist", B=5),
    K.case("sem_cifar10_pde2", "cifar10", B=9, **K.SCRIPT_INSTANCES["cifar10_pde2"]),
    K.case("sem_cifar2", "cifar2", B=6, **K.SCRIPT_INSTANCES["cifar2_diffusion1"]),
    K.case("sem_svhn", "svhn", B=7, **K.SCRIPT_INSTANCES["svhn"]),
    K.case("sem_generic36_svhn", "svhn", B=5, size=36, channels=3, num_steps=3),
    K.case("sem_emotion", "emotion", B=5),
    K.case("sem_tiny", "tiny", B=4, **K.SCRIPT_INSTANCES["tiny"]),
]


def _layer_and_io(c):
    params, (u, g) = K.make_params(c), K.make_io(c)
    layer = runners.make_cuda_layer(c, params)
    return layer, torch.from_numpy(u).cuda(), torch.from_numpy(g).cuda()


def _grads(layer):
    return {n: p.grad.clone() for n, p in layer.named_parameters() if p.grad is not None}


@pytest.mark.parametrize("c", _CASES, ids=lambda c: c.name)
def test_input_is_not_mutated_and_modes_agree(c):
    layer, u, g = _layer_and_io(c)
    keep = u.clone()
    x = u.requires_grad_(True)
    layer.train()
    y_train = layer(x)
    y_train.backward(g)
    assert torch.equal(x.detach(), keep), "forward / backward wrote into the input"
    layer.eval()
    assert torch.equal(layer(x.detach()), y_train.detach()), "eval() differs from train()"
    with torch.no_grad():
        y_ng = layer(keep)          # inference route (for the 28 / 32 planes: the whole-line forward kernel)
    err = max(runners.rel_l2(y_ng.cpu().numpy(), y_train.detach().cpu().numpy()),
              runners.rel_max(y_ng.cpu().numpy(), y_train.detach().cpu().numpy()))
    assert err <= 1e-5, err


@pytest.mark.parametrize("c", _CASES, ids=lambda c: c.name)
def test_second_backward_over_a_retained_graph_repeats_the_first(c):
    layer, u, g = _layer_and_io(c)
    x = u.requires_grad_(True)
    y = layer(x)
    y.backward(g, retain_graph=True)
    first, gin1 = _grads(layer), x.grad.clone()
    layer.zero_grad(set_to_none=True)
    x.grad = None
    y.backward(g)
    second = _grads(layer)
    assert torch.equal(x.grad, gin1)
    assert first.keys() == second.keys()
    for n in first:
        assert torch.equal(first[n], second[n]), n


@pytest.mark.parametrize("c", _CASES, ids=lambda c: c.name)
def test_gradients_accumulate_and_frozen_parameters_stay_clean(c):
    layer, u, g = _layer_and_io(c)
    y = layer(u)
    y.backward(g)
    once = _grads(layer)
    layer(u).backward(g)             # no zero_grad in between: .grad must now hold the sum
    for n, p in layer.named_parameters():
        if n in once:
            assert torch.equal(p.grad, once[n] + once[n]), n
    layer.zero_grad(set_to_none=True)
    names = [n for n, _ in layer.named_parameters()]
    frozen = names[0]
    getattr(layer, frozen).requires_grad_(False)
    layer(u).backward(g)
    assert getattr(layer, frozen).grad is None
    for n, p in layer.named_parameters():
        if n != frozen and n in once:
            assert torch.equal(p.grad, once[n]), n


@pytest.mark.parametrize("c", _CASES, ids=lambda c: c.name)
def test_views_and_non_contiguous_inputs(c):
    layer, u, g = _layer_and_io(c)
    want = layer(u)
    want.backward(g)
    ref = _grads(layer)
    layer.zero_grad(set_to_none=True)
    # every second sample of a twice-as-long batch, planes stored transposed: a strided view on both counts
    big = torch.empty((2 * u.shape[0],) + tuple(u.shape[1:]), device="cuda").normal_()
    big[::2] = u
    xv = big[::2]
    assert not xv.is_contiguous()
    gt = g.transpose(-1, -2).contiguous().transpose(-1, -2)
    assert not gt.is_contiguous() or g.shape[-1] == 1
    y = layer(xv)
    assert torch.equal(y, want)
    y.backward(gt)
    got = _grads(layer)
    for n in ref:
        assert torch.equal(got[n], ref[n]), n
    if u.shape[1] > 1:
        layer.zero_grad(set_to_none=True)
        xcl = u.contiguous(memory_format=torch.channels_last)
        assert torch.equal(layer(xcl), want)


@pytest.mark.parametrize("c", _CASES, ids=lambda c: c.name)
def test_two_calls_in_flight_do_not_share_scratch(c):
    layer, u, g = _layer_and_io(c)
    u2 = torch.flip(u, dims=(0,)) * 0.5 + 0.1
    # one at a time
    y1 = layer(u); y1.backward(g); g1 = _grads(layer); layer.zero_grad(set_to_none=True)
    y2 = layer(u2); y2.backward(g); g2 = _grads(layer); layer.zero_grad(set_to_none=True)
    # both forwards first, then both backwards in the opposite order (each call owns its tables / checkpoints)
    ya, yb = layer(u), layer(u2)
    assert torch.equal(ya, y1) and torch.equal(yb, y2)
    yb.backward(g)
    gb = _grads(layer); layer.zero_grad(set_to_none=True)
    ya.backward(g)
    ga = _grads(layer)
    for n in g1:
        assert torch.equal(ga[n], g1[n]), n
        assert torch.equal(gb[n], g2[n]), n


def test_parameter_update_between_forward_and_backward_uses_the_forward_values():
    """Autograd semantics: the backward pass differentiates the forward that ran, whatever happened to the
    parameters since (an optimiser step under torch.no_grad() bumps their version but the layer saved its
    own factorised tables)."""
    c = K.case("sem_update", "cifar10", B=8, **K.SCRIPT_INSTANCES["cifar10_pde1"])
    params, (u, g) = K.make_params(c), K.make_io(c)
    want = runners.run_oracle(c, params=params, io=(u, g), dtype=np.float32)
    layer = runners.make_cuda_layer(c, params)
    x = torch.from_numpy(u).cuda().requires_grad_(True)
    y = layer(x)
    with torch.no_grad():
        layer.alpha_base.data.mul_(3.0)      # .data: keeps the version counter, as an optimiser's foreach path may
        layer.beta_time_coeff.data.add_(1.0)
    y.backward(torch.from_numpy(g).cuda())
    got = {"gin": x.grad.cpu().numpy()}
    for n, p in layer.named_parameters():
        if p.grad is not None:
            got["g_" + n] = p.grad.cpu().numpy()
    for k, v in got.items():
        err = max(runners.rel_l2(v, want[k]), runners.rel_max(v, want[k]))
        assert err <= 1e-5, (k, err)
