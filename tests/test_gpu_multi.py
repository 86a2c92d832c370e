"""N-GPU vs 1-GPU on hardware ("Distributed"): needs >= 2 CUDA devices, skipped
otherwise (run it with `gpurun --gpus 2 -- python -m pytest tests/test_gpu_multi.py -m gpu`).

  * the coefficient gradients of a batch sharded over two GPUs (ragged split) and summed with the flat
    NCCL all-reduce of cnn_with_pde_b200.parallel equal the single-GPU gradients of the whole batch to
    fp32 reduction-order noise;
  * the data-parallel launcher with the all-reduce captured INSIDE the CUDA graph takes the same
    optimiser steps as with the all-reduce eager between two graphs.
"""
import os
import socket

import numpy as np
import pytest
import torch

from . import cases as K
from . import runners

pytestmark = pytest.mark.gpu
TOL = 1e-5


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _need_two_gpus():
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs two CUDA devices")


def _grads_worker(rank, world, port, tmpdir):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        from cnn_with_pde_b200.parallel import allreduce_coefficient_grads, shard_bounds
        res = {}
        for c in (K.case("dp_cifar10", "cifar10", B=1237, **K.SCRIPT_INSTANCES["cifar10_pde2"]),
                  K.case("dp_svhn", "svhn", B=301, **K.SCRIPT_INSTANCES["svhn"]),
                  K.case("dp_fashion", "fashion", B=5001), K.case("dp_emotion", "emotion", B=77)):
            params, (u, g) = K.make_params(c), K.make_io(c)
            layer = runners.make_cuda_layer(c, params, device=torch.device("cuda", rank))
            lo, hi = shard_bounds(c.B, rank, world)
            x = torch.from_numpy(u[lo:hi]).cuda(rank)
            layer(x).backward(torch.from_numpy(g[lo:hi]).cuda(rank))
            allreduce_coefficient_grads(layer.parameters())
            if rank == 0:
                sharded = {k: p.grad.clone() for k, p in layer.named_parameters() if p.grad is not None}
                for p in layer.parameters():
                    p.grad = None
                layer(torch.from_numpy(u).cuda(0)).backward(torch.from_numpy(g).cuda(0))
                for k, p in layer.named_parameters():
                    if p.grad is not None:
                        a, b = sharded[k].cpu().numpy(), p.grad.cpu().numpy()
                        res[f"{c.name}/{k}"] = max(runners.rel_l2(a, b), runners.rel_max(a, b))
            dist.barrier()
        if rank == 0:
            np.savez(os.path.join(tmpdir, "dp.npz"), **res)
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(600)
def test_sharded_gradients_plus_nccl_allreduce_equal_single_gpu(tmp_path):
    _need_two_gpus()
    import torch.multiprocessing as mp
    mp.spawn(_grads_worker, args=(2, _free_port(), str(tmp_path)), nprocs=2, join=True)
    z = np.load(tmp_path / "dp.npz")
    assert len(z.files) >= 4 * 4
    bad = {k: float(z[k]) for k in z.files if not float(z[k]) <= TOL}
    assert not bad, bad


def _train_worker(rank, world, port, tmpdir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    import torch.distributed as dist
    from cnn_with_pde_b200 import train
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        outs = {}
        for mode, kw in (("in_graph", dict(graph=True, nccl_in_graph=True)), ("between_graphs", dict(graph=True, nccl_in_graph=False)),
                         ("eager", dict(graph=False))):
            warm = 2 + (0 if kw["graph"] else train.GRAPH_PRIMING_STEPS)
            o = train.run("cifar10", 64, 4, warm, quiet=True, no_dropout=True, keep_model=True, **kw)
            outs[mode] = (o["loss"], o["grad_sync"], o["optimizer_steps"],
                          {k: v.detach().cpu() for k, v in o["_model"].state_dict().items()})
        if rank == 0:
            torch.save(outs, os.path.join(tmpdir, "train.pt"))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(900)
def test_launcher_allreduce_inside_the_graph_equals_eager_allreduce(tmp_path):
    _need_two_gpus()
    import torch.multiprocessing as mp
    mp.spawn(_train_worker, args=(2, _free_port(), str(tmp_path)), nprocs=2, join=True)
    outs = torch.load(tmp_path / "train.pt")
    assert outs["in_graph"][1] == "flat all-reduce inside the CUDA graph"
    assert outs["in_graph"][2] == outs["between_graphs"][2] == outs["eager"][2]
    ref_loss, ref_sd = outs["eager"][0], outs["eager"][3]
    for mode in ("in_graph", "between_graphs"):
        loss, sd = outs[mode][0], outs[mode][3]
        assert abs(loss - ref_loss) <= 1e-5 * abs(ref_loss), (mode, loss, ref_loss)
        for k in ref_sd:
            a, b = ref_sd[k].double(), sd[k].double()
            if a.numel() and a.dtype.is_floating_point:
                assert float((a - b).norm() / a.norm().clamp_min(1e-30)) <= 1e-5, (mode, k)
