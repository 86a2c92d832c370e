"""Out-of-bounds writes, checked by hand (compute-sanitizer is closed on the GPU pool: profiles/r02_sanitizer.txt).

Every buffer a kernel writes -- coefficient tables, output, step checkpoints, grad_input, gradient maps,
workspace -- is carved out of one poisoned allocation with a guard band on either side, sized EXACTLY as the
C ABI's size queries say, and the entry points are called through ctypes with those raw pointers.  After the
calls every guard byte must still hold the poison, and the results must match the oracle (so the kernels
did write where they should).  Ragged batches and every kernel variant.
"""
from ctypes import byref

import numpy as np
import pytest

from . import cases as K
from . import runners

pytestmark = pytest.mark.gpu
GUARD = 4096
POISON = 0xA5


class Arena:
    def __init__(self, total):
        import torch
        self.buf = torch.full((total,), POISON, dtype=torch.uint8, device="cuda")
        self.off = 0
        self.spans = []

    def take(self, nbytes, dtype=None, shape=None):
        import torch
        self.off = (self.off + GUARD + 511) // 512 * 512
        lo = self.off
        self.off += max(int(nbytes), 1)
        self.spans.append((lo, self.off))
        raw = self.buf[lo:lo + max(int(nbytes), 1)]
        assert raw.data_ptr() % 256 == 0
        if dtype is None:
            return raw
        return raw.view(dtype).view(shape)

    def check(self, what):
        import torch
        mask = torch.ones_like(self.buf, dtype=torch.bool)
        for lo, hi in self.spans:
            mask[lo:hi] = False
        touched = int((self.buf[mask] != POISON).sum().item())
        assert touched == 0, f"{what}: {touched} guard bytes overwritten"


def _adi_case(c, tuning, need_gin=True):
    import torch
    import cnn_with_pde_b200 as P
    from cnn_with_pde_b200.schedule import adi_schedule
    L = P._cabi.lib()
    params, (u_np, g_np) = K.make_params(c), K.make_io(c)
    layer = runners.make_cuda_layer(c, params)
    cfg = layer._config()
    d = cfg.desc(c.B, tuning)
    sched = adi_schedule(cfg.steps, cfg.dt, cfg.hx, cfg.hy, cfg.lie)
    C, N = cfg.C, cfg.N
    nt, nck = L.pde_adi_tables_bytes(byref(d)), L.pde_adi_checkpoint_bytes(byref(d))
    nws = L.pde_adi_backward_saved_workspace_bytes(byref(d)) if nck else L.pde_adi_backward_workspace_bytes(byref(d))
    cells = c.B * C * N * N
    A = Arena(nt + nck + nws + 3 * 4 * cells + 4 * 4 * C * N * N + 64 * 1024 + 16 * GUARD)
    f32 = torch.float32
    tables, ckpt, ws = A.take(nt), (A.take(nck) if nck else None), A.take(nws)
    out, gin = A.take(4 * cells, f32, (c.B, C, N, N)), (A.take(4 * cells, f32, (c.B, C, N, N)) if need_gin else None)
    gm = [A.take(4 * C * N * N, f32, (C, N, N)) for _ in range(4)]
    chan = getattr(layer, "channel_mixing", getattr(layer, "channel_coupling", None))
    skw = getattr(layer, "skip_weight", None)
    gchan = A.take(4 * C * C, f32, (C, C)) if chan is not None else None
    gskip = A.take(4, f32, ()) if skw is not None else None
    u, g = torch.from_numpy(u_np).cuda(), torch.from_numpy(g_np).cuda()
    st = torch.cuda.current_stream().cuda_stream
    p = lambda t: None if t is None else t.data_ptr()   # noqa: E731
    maps = [layer.alpha_base, layer.beta_base, layer.alpha_time_coeff, layer.beta_time_coeff]
    P._cabi.check(L.pde_adi_prepare(byref(d), byref(sched), *[p(m.detach().contiguous()) for m in maps], p(tables), st), "prepare")
    P._cabi.check(L.pde_adi_forward_train(byref(d), p(tables), p(u), p(chan), p(skw), p(out), p(ckpt), st), "forward")
    P._cabi.check(L.pde_adi_backward_saved(byref(d), p(tables), p(u), p(g), p(chan), p(skw), p(ckpt), p(gin), p(gm[0]), p(gm[1]),
                                           p(gm[2]), p(gm[3]), p(gchan), p(gskip), p(ws), nws, st), "backward")
    torch.cuda.synchronize()
    A.check(f"{c.name} tuning={tuning}")
    want = runners.run_oracle(c, params=params, io=(u_np, g_np), dtype=np.float32, need_gin=need_gin)
    got = {"y": out.cpu().numpy(), "gin": gin.cpu().numpy() if need_gin else None}
    for k, t in zip(("alpha_base", "beta_base", "alpha_time_coeff", "beta_time_coeff"), gm):
        got["g_" + k] = t.cpu().numpy().reshape(np.asarray(params[k]).shape)
    errs = runners.compare(got, {k: v for k, v in want.items() if k in got})
    assert max(errs.values()) <= 1e-5, (c.name, errs)


def test_implicit_kernels_stay_inside_their_buffers():
    import cnn_with_pde_b200 as P
    T = P._cabi.adi_tuning
    todo = [
        (K.case("guard_fashion", "fashion", B=19), 0, True),
        (K.case("guard_fashion_p4q4", "fashion", B=70), T(pairs=4, qf=4), True),
        (K.case("guard_mnist_p4q2", "mnist", B=21, num_steps=3), T(pairs=4, qf=2), False),
        (K.case("guard_cifar10", "cifar10", B=7, **K.SCRIPT_INSTANCES["cifar10_pde3"]), 0, True),
        (K.case("guard_cifar2_q2", "cifar2", B=9, **K.SCRIPT_INSTANCES["cifar2_diffusion2"]), T(qf=2), False),
        (K.case("guard_svhn", "svhn", B=5, size=32, channels=3, num_steps=3), 0, True),
        (K.case("guard_svhn28", "svhn", B=3, size=28, channels=3, num_steps=2), 0, True),
        (K.case("guard_exact", "fashion", B=8, perturb=False, dt=5.0), 0, True),
        (K.case("guard_whole_fashion", "fashion", B=9), T(impl=P._cabi.TUNE_IMPL_WHOLE_LINE), True),
        (K.case("guard_whole_cifar10", "cifar10", B=5, **K.SCRIPT_INSTANCES["cifar10_pde3"]), T(impl=P._cabi.TUNE_IMPL_WHOLE_LINE), True),
        (K.case("guard_svhn16", "svhn", B=3, size=16, channels=3, num_steps=2), 0, True),
        (K.case("guard_mnist12", "mnist", B=5, size=12, num_steps=2), 0, False),
        # adi_generic.cu (plane edge at run time): odd edge, unaligned planes; three channels with both channel ops
        (K.case("guard_generic7", "mnist", B=3, size=7, num_steps=2), 0, False),
        (K.case("guard_generic36_svhn", "svhn", B=5, size=36, channels=3, num_steps=2), 0, True),
        (K.case("guard_generic40_cifar10", "cifar10", B=4, size=40, channels=3, dt=0.01, num_steps=2), 0, True),
    ]
    for c, tuning, need_gin in todo:
        _adi_case(c, tuning, need_gin)


def test_explicit_kernels_stay_inside_their_buffers():
    import torch
    import cnn_with_pde_b200 as P
    from cnn_with_pde_b200 import functional as F
    L = P._cabi.lib()
    f32 = torch.float32
    st = torch.cuda.current_stream().cuda_stream
    p = lambda t: None if t is None else t.data_ptr()   # noqa: E731
    for c, generic in ((K.case("guard_emotion", "emotion", B=9), False), (K.case("guard_emotion24", "emotion", B=5, Nx=24, Ny=24), False),
                       (K.case("guard_emotion_generic", "emotion", B=7, Nx=32, Ny=32, T=0.004), True)):
        params, (u_np, g_np) = K.make_params(c), K.make_io(c)
        layer = runners.make_cuda_layer(c, params)
        d, dref, nws = F._emo_plan(F.EmoConfig(N=layer.Nx, Nt=layer.Nt, dt=layer.dt, dx=layer.dx, dy=layer.dy), c.B, generic, 0)
        n = c.B * layer.Nx * layer.Ny
        A = Arena(nws + 8 * n + 64 * 1024 + 8 * GUARD)
        ws, out, gin, gw = A.take(nws), A.take(4 * n, f32, (c.B, 1, layer.Nx, layer.Ny)), A.take(4 * n, f32, (c.B, 1, layer.Nx, layer.Ny)), A.take(24, f32, (6,))
        w6 = torch.stack([layer.alpha_w1, layer.alpha_w2, layer.alpha_w3, layer.beta_w1, layer.beta_w2, layer.beta_w3]).detach()
        u, g = torch.from_numpy(u_np).cuda(), torch.from_numpy(g_np).cuda()
        P._cabi.check(L.pde_emotion_forward(dref, p(u), p(w6), p(layer.x), p(layer.y), p(out), st), "emotion forward")
        P._cabi.check(L.pde_emotion_backward(dref, p(u), p(g), p(w6), p(layer.x), p(layer.y), p(gin), p(gw), p(ws), nws, st), "emotion backward")
        torch.cuda.synchronize()
        A.check(c.name)
        want = runners.run_oracle(c, params=params, io=(u_np, g_np), dtype=np.float32)
        assert runners.rel_l2(out.cpu().numpy(), want["y"]) <= 1e-5 and runners.rel_l2(gin.cpu().numpy(), want["gin"]) <= 1e-5
    for c in (K.case("guard_tiny", "tiny", B=5, **K.SCRIPT_INSTANCES["tiny"]), K.case("guard_tiny16", "tiny", B=3, size=16, channels=3, num_steps=3, dt=0.02)):
        params, (u_np, g_np) = K.make_params(c), K.make_io(c)
        layer = runners.make_cuda_layer(c, params)
        cfg = F.TinyConfig(steps=layer.num_steps, dt=layer.dt, cmin=layer.stability_eps, cmax=layer.max_coeff)
        d, dref, nws = F._tiny_plan(cfg, tuple(u_np.shape), 0)
        n = u_np.size
        A = Arena(nws + 8 * n + 64 * 1024 + 8 * GUARD)
        ws, out, gin = A.take(nws), A.take(4 * n, f32, u_np.shape), A.take(4 * n, f32, u_np.shape)
        ga, gs = A.take(4 * c.shape[0], f32, (c.shape[0],)), A.take(4 * c.shape[0], f32, (c.shape[0],))
        u, g = torch.from_numpy(u_np).cuda(), torch.from_numpy(g_np).cuda()
        al, sc = layer.alpha_base.detach(), layer.channel_scaling.detach()
        P._cabi.check(L.pde_tiny_forward(dref, p(u), p(al), p(sc), p(out), st), "tiny forward")
        P._cabi.check(L.pde_tiny_backward(dref, p(u), p(g), p(al), p(sc), p(gin), p(ga), p(gs), p(ws), nws, st), "tiny backward")
        torch.cuda.synchronize()
        A.check(c.name)
        want = runners.run_oracle(c, params=params, io=(u_np, g_np), dtype=np.float32)
        assert runners.rel_l2(out.cpu().numpy(), want["y"]) <= 1e-5 and runners.rel_l2(gin.cpu().numpy(), want["gin"]) <= 1e-5
    # the dormant tiny methods: odd plane sizes, ragged last block
    for shape, mode in (((5, 64, 64), 0), ((3, 33, 47), 0), ((7, 30, 64), 4), ((2, 6, 6), 1)):
        n = int(np.prod(shape))
        A = Arena(4 * n + 4 * GUARD)
        out = A.take(4 * n, f32, shape)
        u = torch.randn(*shape, device="cuda")
        d = F._tiny_split_desc(mode, *shape, 0.7, 1.9, 0.3, 1e-6)
        P._cabi.check(L.pde_tiny_split(byref(d), p(u), p(out), st), "tiny split")
        torch.cuda.synchronize()
        A.check(f"tiny_split {shape} mode {mode}")
        assert torch.isfinite(out).all()
