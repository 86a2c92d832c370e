"""torch.autograd.Function wrappers over the C ABI (the only callers of _cabi).

Each Function saves its inputs and, for the implicit family, the factorised coefficient tables of
the call plus -- when the half-line kernels serve it -- the state at the end of every step
(`pde_adi_forward_train`); everything else of the forward trajectory is rebuilt on-chip by the
backward kernels.
"""
from __future__ import annotations

import os

from ctypes import byref
from dataclasses import dataclass

import torch

from . import _cabi
from .schedule import adi_schedule


def _require_cuda(x: torch.Tensor, what: str):
    if not isinstance(x, torch.Tensor) or not x.is_cuda:
        raise RuntimeError(
            f"{what}: expected a CUDA tensor -- the B200 PDE layers have no CPU fallback "
            f"(got {'a ' + str(x.device) + ' tensor' if isinstance(x, torch.Tensor) else type(x)})")
    if x.dtype != torch.float32:
        raise RuntimeError(f"{what}: expected float32, got {x.dtype}")


def _ptr(t):
    return None if t is None else t.data_ptr()


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _bytes(n: int, device) -> torch.Tensor:
    # caching-allocator blocks are 512-byte aligned, which covers the 256-byte requirement
    return torch.empty(max(int(n), 1), dtype=torch.uint8, device=device)


# ------------------------------------------------------------------------------ implicit ADI
@dataclass(frozen=True)
class AdiConfig:
    N: int
    C: int
    steps: int
    dt: float
    hx: float
    hy: float
    lie: bool = False
    smooth: bool = False
    has_max: bool = False
    chan_op: int = 0
    skip: bool = False
    cmin: float = 1e-6
    cmax: float = 10.0
    eps: float = 1e-6

    def desc(self, B: int) -> "_cabi.AdiDesc":
        return _cabi.AdiDesc(B, self.C, self.N, self.steps, int(self.lie), int(self.smooth), int(self.has_max),
                             self.chan_op, int(self.skip), self.cmin, self.cmax, self.eps)


class _AdiFunction(torch.autograd.Function):
    @staticmethod
    @torch.amp.custom_fwd(device_type="cuda", cast_inputs=torch.float32)
    def forward(ctx, u, alpha_base, beta_base, alpha_tc, beta_tc, chan, skipw, cfg: AdiConfig, grad_mode: bool = True):
        _require_cuda(u, "PDE layer input")
        L = _cabi.lib()
        u = u.contiguous()
        B = u.shape[0]
        maps = [p.detach().contiguous() for p in (alpha_base, beta_base, alpha_tc, beta_tc)]
        for p in maps:
            _require_cuda(p, "PDE layer parameter")
        chan_c = None if chan is None else chan.detach().contiguous()
        skip_c = None if skipw is None else skipw.detach().contiguous()
        d = cfg.desc(B)
        with torch.cuda.device(u.device):
            tables = _bytes(L.pde_adi_tables_bytes(byref(d)), u.device)
            if tables.numel() <= 1:
                raise _cabi.PdeB200Error(
                    f"unsupported implicit-layer configuration (size={cfg.N}, channels={cfg.C}, steps={cfg.steps}); "
                    "supported: size in {8,12,16,28,32}, channels <= 4")
            sched = adi_schedule(cfg.steps, cfg.dt, cfg.hx, cfg.hy, cfg.lie)
            _cabi.check(L.pde_adi_prepare(byref(d), byref(sched), *[_ptr(p) for p in maps], _ptr(tables), _stream()),
                        "pde_adi_prepare")
            out = torch.empty_like(u)
            # under autograd the forward kernel also writes the state at the end of every step for
            # the backward kernel (0 bytes when the configuration is served by the kernels that
            # rebuild the trajectory on-chip)
            # PDE_B200_NO_CKPT=1 trades them for a recomputation inside pde_adi_backward (saves
            # num_steps x the input in memory between forward and backward)
            # (needs_input_grad ignores torch.no_grad() and grad mode is always off inside forward():
            # the caller passes the mode it saw)
            want_ck = grad_mode and any(ctx.needs_input_grad) and not os.environ.get("PDE_B200_NO_CKPT")
            ck_bytes = L.pde_adi_checkpoint_bytes(byref(d)) if want_ck else 0
            ckpt = _bytes(ck_bytes, u.device) if ck_bytes else None
            _cabi.check(L.pde_adi_forward_train(byref(d), _ptr(tables), _ptr(u), _ptr(chan_c), _ptr(skip_c), _ptr(out),
                                                _ptr(ckpt), _stream()), "pde_adi_forward_train")
        ctx.cfg = cfg
        ctx.has_chan = chan is not None
        ctx.has_skip = skipw is not None
        ctx.has_ckpt = ckpt is not None
        ctx.save_for_backward(u, tables, *(t for t in (chan_c, skip_c, ckpt) if t is not None))
        ctx.param_shapes = [p.shape for p in (alpha_base, beta_base, alpha_tc, beta_tc)]
        return out

    @staticmethod
    @torch.amp.custom_bwd(device_type="cuda")
    def backward(ctx, gout):
        L = _cabi.lib()
        cfg = ctx.cfg
        saved = list(ctx.saved_tensors)
        u, tables = saved[0], saved[1]
        rest = saved[2:]
        chan = rest.pop(0) if ctx.has_chan else None
        skipw = rest.pop(0) if ctx.has_skip else None
        ckpt = rest.pop(0) if ctx.has_ckpt else None
        gout = gout.contiguous().float()
        B = u.shape[0]
        d = cfg.desc(B)
        with torch.cuda.device(u.device):
            ws_bytes = (L.pde_adi_backward_saved_workspace_bytes if ckpt is not None
                        else L.pde_adi_backward_workspace_bytes)(byref(d))
            ws = _bytes(ws_bytes, u.device)
            gin = torch.empty_like(u) if ctx.needs_input_grad[0] else None
            gmaps = [torch.empty((cfg.C, cfg.N, cfg.N), dtype=torch.float32, device=u.device) for _ in range(4)]
            gchan = torch.empty((cfg.C, cfg.C), dtype=torch.float32, device=u.device) if chan is not None else None
            gskip = torch.empty((), dtype=torch.float32, device=u.device) if skipw is not None else None
            _cabi.check(L.pde_adi_backward_saved(byref(d), _ptr(tables), _ptr(u), _ptr(gout), _ptr(chan), _ptr(skipw),
                                                 _ptr(ckpt), _ptr(gin), _ptr(gmaps[0]), _ptr(gmaps[1]), _ptr(gmaps[2]),
                                                 _ptr(gmaps[3]), _ptr(gchan), _ptr(gskip), _ptr(ws), ws_bytes,
                                                 _stream()), "pde_adi_backward_saved")
        gmaps = [g.reshape(s) for g, s in zip(gmaps, ctx.param_shapes)]
        return (gin, gmaps[0], gmaps[1], gmaps[2], gmaps[3], gchan, gskip, None, None)


def adi_layer(u, alpha_base, beta_base, alpha_tc, beta_tc, chan, skipw, cfg: AdiConfig):
    return _AdiFunction.apply(u, alpha_base, beta_base, alpha_tc, beta_tc, chan, skipw, cfg, torch.is_grad_enabled())


# ------------------------------------------------------------------- explicit, frozen ghost ring
@dataclass(frozen=True)
class EmoConfig:
    N: int
    Nt: int
    dt: float
    dx: float
    dy: float

    def desc(self, B: int) -> "_cabi.EmoDesc":
        import numpy as np
        f = lambda v: float(np.float32(v))
        return _cabi.EmoDesc(B, self.N, self.Nt, f(0.5 * self.dt), f(self.dt), f(self.dx ** 2), f(self.dy ** 2))


class _EmotionFunction(torch.autograd.Function):
    @staticmethod
    @torch.amp.custom_fwd(device_type="cuda", cast_inputs=torch.float32)
    def forward(ctx, u0, w6, xs, ys, cfg: EmoConfig):
        _require_cuda(u0, "PDELayer input")
        L = _cabi.lib()
        u0 = u0.contiguous()
        w6c, xs, ys = w6.detach().contiguous(), xs.contiguous(), ys.contiguous()
        d = cfg.desc(u0.shape[0])
        with torch.cuda.device(u0.device):
            out = torch.empty_like(u0)
            _cabi.check(L.pde_emotion_forward(byref(d), _ptr(u0), _ptr(w6c), _ptr(xs), _ptr(ys), _ptr(out), _stream()),
                        "pde_emotion_forward")
        ctx.cfg = cfg
        ctx.save_for_backward(u0, w6c, xs, ys)
        return out

    @staticmethod
    @torch.amp.custom_bwd(device_type="cuda")
    def backward(ctx, gout):
        L = _cabi.lib()
        u0, w6, xs, ys = ctx.saved_tensors
        gout = gout.contiguous().float()
        d = ctx.cfg.desc(u0.shape[0])
        with torch.cuda.device(u0.device):
            ws_bytes = L.pde_emotion_backward_workspace_bytes(byref(d))
            ws = _bytes(ws_bytes, u0.device)
            gin = torch.empty_like(u0) if ctx.needs_input_grad[0] else None
            gw = torch.empty(6, dtype=torch.float32, device=u0.device)
            _cabi.check(L.pde_emotion_backward(byref(d), _ptr(u0), _ptr(gout), _ptr(w6), _ptr(xs), _ptr(ys), _ptr(gin),
                                               _ptr(gw), _ptr(ws), ws_bytes, _stream()), "pde_emotion_backward")
        return gin, gw, None, None, None


def emotion_layer(u0, w6, xs, ys, cfg: EmoConfig):
    return _EmotionFunction.apply(u0, w6, xs, ys, cfg)


# ----------------------------------------------------------------------- explicit, zero ghosts
@dataclass(frozen=True)
class TinyConfig:
    steps: int
    dt: float
    cmin: float
    cmax: float
    blend: float = 0.1


class _TinyFunction(torch.autograd.Function):
    @staticmethod
    @torch.amp.custom_fwd(device_type="cuda", cast_inputs=torch.float32)
    def forward(ctx, u, alpha_base, channel_scaling, cfg: TinyConfig):
        _require_cuda(u, "ImprovedDiffusionLayer input")
        L = _cabi.lib()
        u = u.contiguous()
        B, C, H, W = u.shape
        al, sc = alpha_base.detach().contiguous(), channel_scaling.detach().contiguous()
        d = _cabi.TinyDesc(B, C, H, W, cfg.steps, cfg.dt, cfg.cmin, cfg.cmax, cfg.blend)
        with torch.cuda.device(u.device):
            out = torch.empty_like(u)
            _cabi.check(L.pde_tiny_forward(byref(d), _ptr(u), _ptr(al), _ptr(sc), _ptr(out), _stream()),
                        "pde_tiny_forward")
        ctx.cfg = cfg
        ctx.save_for_backward(u, al, sc)
        return out

    @staticmethod
    @torch.amp.custom_bwd(device_type="cuda")
    def backward(ctx, gout):
        L = _cabi.lib()
        u, al, sc = ctx.saved_tensors
        cfg = ctx.cfg
        gout = gout.contiguous().float()
        B, C, H, W = u.shape
        d = _cabi.TinyDesc(B, C, H, W, cfg.steps, cfg.dt, cfg.cmin, cfg.cmax, cfg.blend)
        with torch.cuda.device(u.device):
            ws_bytes = L.pde_tiny_backward_workspace_bytes(byref(d))
            ws = _bytes(ws_bytes, u.device)
            gin = torch.empty_like(u) if ctx.needs_input_grad[0] else None
            ga = torch.empty(C, dtype=torch.float32, device=u.device)
            gs = torch.empty(C, dtype=torch.float32, device=u.device)
            _cabi.check(L.pde_tiny_backward(byref(d), _ptr(u), _ptr(gout), _ptr(al), _ptr(sc), _ptr(gin), _ptr(ga),
                                            _ptr(gs), _ptr(ws), ws_bytes, _stream()), "pde_tiny_backward")
        return gin, ga, gs, None


def tiny_layer(u, alpha_base, channel_scaling, cfg: TinyConfig):
    return _TinyFunction.apply(u, alpha_base, channel_scaling, cfg)
