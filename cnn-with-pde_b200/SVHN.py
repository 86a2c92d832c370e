"""Drop-in for SVHN.DiffusionLayer (SVHN.py:12-230)."""
import torch
import torch.nn as nn

from ._base import cached_config, check_input, smooth_coefficients as _smooth
from .functional import AdiConfig, adi_layer


class DiffusionLayer(nn.Module):
    """C-channel smoothed Strang ADI with per-channel maps, post-step channel coupling
    u <- K u (SVHN.py:71,78-86) and a sigmoid skip epilogue (SVHN.py:74)."""

    def __init__(self, size=32, channels=3, dt=0.01, dx=1.0, num_steps=10):
        super().__init__()
        self.size = size
        self.channels = channels
        self.dt = dt
        self.dx = dx
        self.num_steps = num_steps
        self.alpha_base = nn.Parameter(torch.ones(channels, size, size) * 0.1)
        self.beta_base = nn.Parameter(torch.ones(channels, size, size) * 0.1)
        # same RNG draws, in the same order, as the reference constructor (SVHN.py:26-27)
        self.alpha_time_coeff = nn.Parameter(torch.randn(channels, size, size) * 0.001)
        self.beta_time_coeff = nn.Parameter(torch.randn(channels, size, size) * 0.001)
        self.channel_coupling = nn.Parameter(torch.eye(channels) * 0.01)
        self.stability_eps = 1e-6
        self.skip_weight = nn.Parameter(torch.tensor(0.9))

    def _config(self) -> AdiConfig:
        return cached_config(self, (self.size, self.channels, self.num_steps, self.dt, self.dx, self.stability_eps), lambda: AdiConfig(
            N=self.size, C=self.channels, steps=self.num_steps, dt=self.dt, hx=self.dx, hy=self.dx, smooth=True,
            chan_op=2, skip=True, cmin=self.stability_eps, eps=self.stability_eps))

    def get_alpha_beta_at_time(self, t):
        alpha_t = torch.clamp(self.alpha_base + self.alpha_time_coeff * t, min=self.stability_eps)
        beta_t = torch.clamp(self.beta_base + self.beta_time_coeff * t, min=self.stability_eps)
        return alpha_t, beta_t

    def forward(self, u):
        check_input(u, self.channels, self.size, self.size, "DiffusionLayer")
        return adi_layer(u, self.alpha_base, self.beta_base, self.alpha_time_coeff, self.beta_time_coeff,
                         self.channel_coupling, self.skip_weight, self._config())

    def apply_channel_coupling(self, u):
        """u[b, c, p] <- sum_d K[c, d] u[b, d, p] (PyTorch helper; the kernels fuse it)."""
        return torch.einsum("cd,bdhw->bchw", self.channel_coupling, u)

    def smooth_coefficients(self, coeffs, dim=1, kernel_size=3):
        return _smooth(coeffs, dim, kernel_size)


from .classifiers import SvhnPDEClassifier as PDEClassifier  # noqa: E402,F401  (SVHN.py:234)
