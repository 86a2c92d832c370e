"""Drop-in for tiny_imagenet.ImprovedDiffusionLayer, live path (tiny_imagenet.py:14-72)."""
import torch
import torch.nn as nn

from .functional import TinyConfig, tiny_layer


class ImprovedDiffusionLayer(nn.Module):
    """Explicit per-channel step u <- u + 0.1 ((s u + alpha dt Lap0(s u)) - u), zero ghosts.
    ``beta_base`` and ``use_implicit`` exist but are unused, exactly as in the reference
    (tiny_imagenet.py:21,26): ``beta_base.grad`` stays None."""

    def __init__(self, size=64, channels=3, dt=0.01, num_steps=1, use_implicit=False):
        super().__init__()
        self.size = size
        self.channels = channels
        self.dt = dt
        self.num_steps = num_steps
        self.use_implicit = use_implicit
        self.alpha_base = nn.Parameter(torch.ones(channels) * 0.05)
        self.beta_base = nn.Parameter(torch.ones(channels) * 0.05)
        self.channel_scaling = nn.Parameter(torch.ones(channels))
        self.stability_eps = 1e-6
        self.max_coeff = 0.15

    def forward(self, u):
        if u.dim() != 4 or u.shape[1] != self.channels:
            raise ValueError(f"ImprovedDiffusionLayer: expected (B, {self.channels}, H, W), got {tuple(u.shape)}")
        cfg = TinyConfig(steps=self.num_steps, dt=self.dt, cmin=self.stability_eps, cmax=self.max_coeff)
        return tiny_layer(u, self.alpha_base, self.channel_scaling, cfg)


from .classifiers import BasicBlock, ImprovedTinyImageNetClassifier  # noqa: E402,F401  (tiny_imagenet.py:237-327)
