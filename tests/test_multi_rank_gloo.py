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
        unused = torch.nn.Parameter(torch.zeros(3, dtype=torch.float64))   # no .grad on any rank: stays None
        # a gradient on rank 0 only (an empty shard / a skipped branch on the other rank): the flat
        # buffers must still line up and both ranks must end with the sum
        lopsided = torch.nn.Parameter(torch.zeros(4, dtype=torch.float64))
        if rank == 0:
            lopsided.grad = torch.arange(4, dtype=torch.float64)
        allreduce_coefficient_grads([unused] + plist[:2] + [lopsided] + plist[2:])
        assert unused.grad is None
        assert lopsided.grad is not None and torch.equal(lopsided.grad, torch.arange(4, dtype=torch.float64))
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


def _flat_worker(rank, world, port, tmpdir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from cnn_with_pde_b200.train import FlatGradSync
        torch.manual_seed(0)
        net = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.ReLU(), torch.nn.Linear(5, 3))
        sync = FlatGradSync(net, world)
        x = torch.arange(24, dtype=torch.float32).reshape(4, 6) / 10.0
        for _ in range(2):                      # second pass: the views must survive zero() + backward
            sync.zero()
            net(x[2 * rank:2 * rank + 2]).square().mean().backward()
            assert all(p.grad.untyped_storage().data_ptr() == sync.flat.untyped_storage().data_ptr()
                       for p in net.parameters())
            sync.all_reduce()
        if rank == 0:
            torch.save([p.grad.clone() for p in net.parameters()], os.path.join(tmpdir, "flat.pt"))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_flat_gradient_sync_matches_full_batch_mean(tmp_path):
    """The launcher's gradient sync (every .grad a view of one flat buffer, one all-reduce, mean
    over ranks as DDP) on 2 CPU ranks equals the single-process gradient of the mean loss."""
    world = 2
    mp.spawn(_flat_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    got = torch.load(tmp_path / "flat.pt")
    torch.manual_seed(0)
    net = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.ReLU(), torch.nn.Linear(5, 3))
    x = torch.arange(24, dtype=torch.float32).reshape(4, 6) / 10.0
    (0.5 * (net(x[:2]).square().mean() + net(x[2:]).square().mean())).backward()
    for g, p in zip(got, net.parameters()):
        torch.testing.assert_close(g, p.grad, rtol=1e-6, atol=1e-7)
