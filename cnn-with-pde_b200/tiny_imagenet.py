"""Drop-in for tiny_imagenet.ImprovedDiffusionLayer, live path (tiny_imagenet.py:14-72)."""
import torch
import torch.nn as nn

from .functional import TinyConfig, tiny_layer, tiny_split


class ImprovedDiffusionLayer(nn.Module):
    """Explicit per-channel step u <- u + 0.1 ((s u + alpha dt Lap0(s u)) - u), zero ghosts.
    ``beta_base`` and ``use_implicit`` exist but are unused by ``forward``, exactly as in the reference
    (tiny_imagenet.py:21,26): ``beta_base.grad`` stays None.  The methods the reference defines but never
    calls -- the scalar-coefficient ADI step and the explicit x / y splits (tiny_imagenet.py:88-233) --
    are served by their own kernel for whoever wires ``use_implicit`` up."""

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

    # ---- dormant in the reference (never reached from forward): planes u (B, H, W), Python-number coefficients
    def implicit_diffusion_step(self, u, alpha_coeff, beta_coeff):
        """ADI step: implicit x solve then implicit y solve, dt / 2 each (tiny_imagenet.py:88-102)."""
        return tiny_split("implicit_diffusion_step", u, alpha_coeff, beta_coeff, self.dt, self.stability_eps)

    def solve_implicit_x(self, u, coeff, dt):
        return tiny_split("solve_implicit_x", u, coeff, 0.0, dt, self.stability_eps)      # tiny_imagenet.py:104-130

    def solve_implicit_y(self, u, coeff, dt):
        return tiny_split("solve_implicit_y", u, 0.0, coeff, dt, self.stability_eps)      # tiny_imagenet.py:132-157

    def diffuse_x_explicit(self, u, coeff):
        return tiny_split("diffuse_x_explicit", u, coeff, 0.0, self.dt, self.stability_eps)   # tiny_imagenet.py:199-215

    def diffuse_y_explicit(self, u, coeff):
        return tiny_split("diffuse_y_explicit", u, 0.0, coeff, self.dt, self.stability_eps)   # tiny_imagenet.py:217-233

    def thomas_algorithm_batch(self, a, b, c, d):
        """General batched Thomas solve with pivots clamped at stability_eps (tiny_imagenet.py:159-190), for
        arbitrary (batch, n) bands.  A PyTorch helper, off the hot path: the solve_implicit_* methods above do
        not go through it (their bands are constants, factorised once inside the kernel)."""
        n = d.shape[1]
        cp, dp = [c[:, 0] / b[:, 0]], [d[:, 0] / b[:, 0]]
        for i in range(1, n):
            den = torch.clamp(b[:, i] - a[:, i] * cp[-1], min=self.stability_eps)
            cp.append(c[:, i] / den if i < n - 1 else torch.zeros_like(den))
            dp.append((d[:, i] - a[:, i] * dp[-1]) / den)
        xs = [dp[-1]]
        for i in range(n - 2, -1, -1):
            xs.append(dp[i] - cp[i] * xs[-1])
        return torch.stack(xs[::-1], dim=1)


from .classifiers import BasicBlock, ImprovedTinyImageNetClassifier  # noqa: E402,F401  (tiny_imagenet.py:237-327)
