"""Drop-in for emotion_recognition.PDELayer (emotion_recognition.py:56-97)."""
import torch
import torch.nn as nn

from ._base import check_input
from .functional import EmoConfig, emotion_layer


class PDELayer(nn.Module):
    """Explicit 5-point update on a reflect-padded plane whose ghost ring is frozen at its initial
    values; row/column coefficient profiles from six learnable scalars."""

    def __init__(self, Nx=48, Ny=48, Lx=1.0, Ly=1.0, T=0.01, dt=0.001):
        super().__init__()
        self.Nx, self.Ny, self.Lx, self.Ly = Nx, Ny, Lx, Ly
        self.T, self.dt = T, dt
        self.dx = Lx / Nx
        self.dy = Ly / Ny
        self.Nt = int(T / dt)
        self.alpha_w1 = nn.Parameter(torch.tensor(0.1))
        self.alpha_w2 = nn.Parameter(torch.tensor(0.1))
        self.alpha_w3 = nn.Parameter(torch.tensor(0.1))
        self.beta_w1 = nn.Parameter(torch.tensor(0.3))
        self.beta_w2 = nn.Parameter(torch.tensor(0.2))
        self.beta_w3 = nn.Parameter(torch.tensor(0.2))
        self.register_buffer("x", torch.linspace(0, Lx, Nx))
        self.register_buffer("y", torch.linspace(0, Ly, Ny))

    def alpha(self, y_val):
        return 0.5 * self.dt * (self.alpha_w1 + self.alpha_w2 * torch.sin(2 * torch.pi * y_val)
                                + self.alpha_w3 * torch.sin(4 * torch.pi * y_val)) / self.dx ** 2

    def beta(self, x_val):
        return self.dt * (self.beta_w1 + self.beta_w2 * torch.cos(2 * torch.pi * x_val)
                          + self.beta_w3 * torch.cos(4 * torch.pi * x_val)) / self.dy ** 2

    def forward(self, u0):
        if self.Nx != self.Ny:
            # the reference broadcasts a (Ny, Nx) coefficient grid onto a (Nx, Ny) plane
            raise ValueError("PDELayer requires Nx == Ny")
        check_input(u0, 1, self.Nx, self.Ny, "PDELayer")
        w6 = torch.stack([self.alpha_w1, self.alpha_w2, self.alpha_w3, self.beta_w1, self.beta_w2, self.beta_w3])
        cfg = EmoConfig(N=self.Nx, Nt=self.Nt, dt=self.dt, dx=self.dx, dy=self.dy)
        return emotion_layer(u0, w6, self.x, self.y, cfg)


from .classifiers import DiffusionClassifier  # noqa: E402,F401  (emotion_recognition.py:170)
