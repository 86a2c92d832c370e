"""Drop-in for cifar10.EnhancedDiffusionLayer (cifar10.py:24-211)."""
import torch
import torch.nn as nn

from ._base import cached_config, check_input
from .functional import AdiConfig, adi_layer, adi_multi_layer


class EnhancedDiffusionLayer(nn.Module):
    """C-channel Strang ADI, unsmoothed maps clamped to [1e-6, 10], pre-step channel mixing
    u <- M u (cifar10.py:91)."""

    _lie = False

    def __init__(self, size=32, channels=3, dt=0.001, dx=1.0, dy=1.0, num_steps=10):
        super().__init__()
        self.size = size
        self.channels = channels
        self.dt = dt
        self.dx = dx
        self.dy = dy
        self.num_steps = num_steps
        self.alpha_base = nn.Parameter(torch.ones(channels, size, size) * 1.0)
        self.beta_base = nn.Parameter(torch.ones(channels, size, size) * 1.0)
        self.alpha_time_coeff = nn.Parameter(torch.zeros(channels, size, size) * 0.1)
        self.beta_time_coeff = nn.Parameter(torch.zeros(channels, size, size) * 0.1)
        # one randn(C, C) draw, as in the reference constructor (cifar10.py:44)
        self.channel_mixing = nn.Parameter(torch.eye(channels) + torch.randn(channels, channels) * 0.01)
        self.stability_eps = 1e-6

    def _config(self) -> AdiConfig:
        return cached_config(self, (self.size, self.channels, self.num_steps, self.dt, self.dx, self.dy, self.stability_eps), lambda: AdiConfig(
            N=self.size, C=self.channels, steps=self.num_steps, dt=self.dt, hx=self.dx, hy=self.dy, lie=self._lie,
            has_max=True, chan_op=1, cmin=self.stability_eps, cmax=10.0, eps=self.stability_eps))

    def get_alpha_beta_at_time(self, t):
        alpha_t = torch.clamp(self.alpha_base + self.alpha_time_coeff * t, min=self.stability_eps, max=10.0)
        beta_t = torch.clamp(self.beta_base + self.beta_time_coeff * t, min=self.stability_eps, max=10.0)
        return alpha_t, beta_t

    def apply_channel_mixing(self, u):
        """u[b, c, p] <- sum_d M[c, d] u[b, d, p] (PyTorch helper; the kernels fuse it)."""
        B, C, H, W = u.shape
        return torch.matmul(self.channel_mixing, u.reshape(B, C, -1)).view(B, C, H, W)

    def forward(self, u):
        check_input(u, self.channels, self.size, self.size, type(self).__name__)
        return adi_layer(u, self.alpha_base, self.beta_base, self.alpha_time_coeff, self.beta_time_coeff,
                         self.channel_mixing, None, self._config())

    def _branch(self):
        """This layer as one branch of a multi-layer call (apply_to_same_input)."""
        return (self.alpha_base, self.beta_base, self.alpha_time_coeff, self.beta_time_coeff, self.channel_mixing, None,
                self._config())


def apply_to_same_input(u, layers):
    """layers[i](u) for several diffusion layers that read the same input -- the branch loops of
    MultiScaleExtractor.forward (cifar10.py:272-274) and HybridPDEExtractor.forward (cifar_2version.py:287-288) --
    with one launch per pass for all of them (one coefficient-table launch, one forward, one backward, one
    gradient finish) instead of one per layer."""
    for layer in layers:
        check_input(u, layer.channels, layer.size, layer.size, type(layer).__name__)
    return adi_multi_layer(u, [layer._branch() for layer in layers])


from .classifiers import CIFAR10PDENoConv, EnhancedFC, MultiScaleExtractor, SpatialAttention  # noqa: E402,F401  (cifar10.py:215-361)
