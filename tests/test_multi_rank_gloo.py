"""world_size-2 gloo test of the N>1 path's host logic (CPU; the kernels need a GPU).

Each rank takes its shard of the batch, computes the coefficient gradients of its shard (the C
oracle stands in for the CUDA backward here -- checker only), and the flat all-reduce of
cnn_with_pde_b200.parallel must reproduce the single-process gradients of the full batch.
"""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from . import cases as K
from . import runners


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, tmpdir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from cnn_with_pde_b200.parallel import allreduce_coefficient_grads, shard_bounds
        c = K.case("ddp", "cifar10", B=5, size=16, channels=3, num_steps=2, dt=0.01)   # ragged: 3 + 2
        params, (u, g) = K.make_params(c), K.make_io(c)
        lo, hi = shard_bounds(c.B, rank, world)
        cs = K.case("ddp_shard", "cifar10", B=hi - lo, size=16, channels=3, num_steps=2, dt=0.01)
        r = runners.run_oracle(cs, params=params, io=(u[lo:hi], g[lo:hi]), dtype=np.float64)
        names = K.grad_names(c)
        plist = []
        for n in names:
            p = torch.nn.Parameter(torch.from_numpy(np.asarray(params[n], np.float64)))
            p.grad = torch.from_numpy(np.asarray(r["g_" + n], np.float64).copy())
            plist.append(p)
        unused = torch.nn.Parameter(torch.zeros(3))      # no .grad: must be skipped, not crash
        allreduce_coefficient_grads(plist + [unused])
        if rank == 0:
            np.savez(os.path.join(tmpdir, "reduced.npz"), **{n: p.grad.numpy() for n, p in zip(names, plist)})
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_sharded_backward_plus_allreduce_equals_full_batch(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    got = np.load(tmp_path / "reduced.npz")
    c = K.case("ddp", "cifar10", B=5, size=16, channels=3, num_steps=2, dt=0.01)
    full = runners.run_oracle(c, dtype=np.float64)
    for n in K.grad_names(c):
        np.testing.assert_allclose(got[n], full["g_" + n], rtol=1e-11, atol=1e-13, err_msg=n)
