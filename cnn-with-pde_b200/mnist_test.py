"""Drop-in for mnist_test.DiffusionLayer (mnist_test.py:11-219)."""
import torch
import torch.nn as nn

from ._base import cached_config, check_input, smooth_coefficients as _smooth
from .functional import AdiConfig, adi_layer


class DiffusionLayer(nn.Module):
    """Implicit Strang-split diffusion on 1 x size x size images with per-pixel alpha/beta maps,
    a linear-in-time term, 3-tap coefficient smoothing and separate dx / dy."""

    def __init__(self, size=28, dt=0.001, dx=1.0, dy=1.0, num_steps=10):
        super().__init__()
        self.size = size
        self.dt = dt
        self.dx = dx
        self.dy = dy
        self.num_steps = num_steps
        self.alpha_base = nn.Parameter(torch.ones(size, size) * 2.0)
        self.beta_base = nn.Parameter(torch.ones(size, size) * 2.0)
        self.alpha_time_coeff = nn.Parameter(torch.zeros(size, size))
        self.beta_time_coeff = nn.Parameter(torch.zeros(size, size))
        self.stability_eps = 1e-6

    def _config(self) -> AdiConfig:
        return cached_config(self, (self.size, self.num_steps, self.dt, self.dx, self.dy, self.stability_eps), lambda: AdiConfig(
            N=self.size, C=1, steps=self.num_steps, dt=self.dt, hx=self.dx, hy=self.dy, smooth=True,
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

    def get_numerical_stability_info(self):
        with torch.no_grad():
            alpha_max = torch.max(self.alpha_base + torch.abs(self.alpha_time_coeff) * self.dt * self.num_steps)
            beta_max = torch.max(self.beta_base + torch.abs(self.beta_time_coeff) * self.dt * self.num_steps)
            cfl_x = alpha_max * self.dt / (self.dx ** 2)
            cfl_y = beta_max * self.dt / (self.dy ** 2)
            return {"cfl_x": cfl_x.item(), "cfl_y": cfl_y.item(), "dx": self.dx, "dy": self.dy, "dt": self.dt,
                    "stable_x": cfl_x.item() < 0.5, "stable_y": cfl_y.item() < 0.5}


from .classifiers import MnistPDEClassifier as PDEClassifier  # noqa: E402,F401  (mnist_test.py:223)
