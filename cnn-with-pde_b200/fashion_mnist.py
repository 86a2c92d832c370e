"""Drop-in for fashion_mnist.DiffusionLayer (fashion_mnist.py:18-196)."""
import torch
import torch.nn as nn

from ._base import cached_config, check_input, smooth_coefficients as _smooth
from .functional import AdiConfig, adi_layer


class DiffusionLayer(nn.Module):
    """Same scheme as the MNIST layer with dt=0.3, 4 steps, init 1.8 and no dy: the y sweeps use
    dx as their spacing (fashion_mnist.py:63)."""

    def __init__(self, size=28, dt=0.3, dx=1.0, num_steps=4):
        super().__init__()
        self.size = size
        self.dt = dt
        self.dx = dx
        self.num_steps = num_steps
        self.alpha_base = nn.Parameter(torch.ones(size, size) * 1.8)
        self.beta_base = nn.Parameter(torch.ones(size, size) * 1.8)
        self.alpha_time_coeff = nn.Parameter(torch.zeros(size, size))
        self.beta_time_coeff = nn.Parameter(torch.zeros(size, size))
        self.stability_eps = 1e-6

    def _config(self) -> AdiConfig:
        return cached_config(self, (self.size, self.num_steps, self.dt, self.dx, self.stability_eps), lambda: AdiConfig(
            N=self.size, C=1, steps=self.num_steps, dt=self.dt, hx=self.dx, hy=self.dx, smooth=True,
            cmin=self.stability_eps, eps=self.stability_eps))

    def get_alpha_beta_at_time(self, t):
        alpha_t = torch.clamp(self.alpha_base + self.alpha_time_coeff * t, min=self.stability_eps)
        beta_t = torch.clamp(self.beta_base + self.beta_time_coeff * t, min=self.stability_eps)
        return alpha_t, beta_t

    def forward(self, u):
        check_input(u, 1, self.size, self.size, "DiffusionLayer")
        return adi_layer(u, self.alpha_base, self.beta_base, self.alpha_time_coeff, self.beta_time_coeff,
                         None, None, self._config())

    def smooth_coefficients(self, coeffs, dim=1, kernel_size=3):
        return _smooth(coeffs, dim, kernel_size)


from .classifiers import FashionPDEClassifier  # noqa: E402,F401  (fashion_mnist.py:200)
