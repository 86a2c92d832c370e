import os, time, torch
print("ALLOC_CONF", os.environ.get("PYTORCH_CUDA_ALLOC_CONF"), os.environ.get("PYTORCH_ALLOC_CONF"))
u = torch.randn(256,1,28,28, device="cuda")
def t(fn, n=2000):
    for _ in range(20): fn()
    torch.cuda.synchronize()
    t0=time.perf_counter()
    for _ in range(n): fn()
    dt=(time.perf_counter()-t0)/n*1e6
    torch.cuda.synchronize()
    return dt
print("empty_like (freed each iter)", t(lambda: torch.empty_like(u)))
print("empty same size", t(lambda: torch.empty(u.shape, device="cuda")))
print("empty uint8 1MB", t(lambda: torch.empty(1<<20, dtype=torch.uint8, device="cuda")))
print("empty uint8 20MB", t(lambda: torch.empty(20<<20, dtype=torch.uint8, device="cuda")))
keep=[]
def hold():
    keep.append(torch.empty_like(u))
    if len(keep)>3: keep.pop(0)
print("empty_like (3 alive)", t(hold))
x = u.clone().requires_grad_(True)
w = torch.randn(1, device="cuda", requires_grad=True)
def fb():
    y = (x*w)
    y.backward(u)
print("mul fwd+bwd", t(fb, 500))
import torch.autograd as A
class F(A.Function):
    @staticmethod
    def forward(ctx, a, b):
        ctx.save_for_backward(a,b); return torch.empty_like(a)
    @staticmethod
    def backward(ctx, g):
        a,b=ctx.saved_tensors
        return torch.empty_like(a), torch.empty_like(b)
def fb2():
    y = F.apply(x,w); y.backward(u)
print("custom fn fwd+bwd", t(fb2, 500))
def fb3():
    x.grad=None; w.grad=None
    y = F.apply(x,w); y.backward(u)
print("custom fn fwd+bwd, grads reset", t(fb3, 500))
