"""world_size-2 gloo test of the relation-sharding host logic (planner, edge partition, autograd collectives).

CPU only: the per-rank compute is stood in for by the oracle (tests may use it); on GPUs the same host code
drives the CUDA engine (tests/test_gpu_parity.py::test_relation_sharded_* under torchrun-less single GPU,
bench.py --gpus N for real).
"""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import rgcn_oracle as orc


def _free_port():
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        return s.getsockname()[1]


class _OracleShard(torch.autograd.Function):
    """Partial propagate over this rank's edges, computed by the oracle (stand-in for the CUDA kernels)."""

    @staticmethod
    def forward(ctx, x, W, tp, val):
        ctx.save_for_backward(x, W, tp, val)
        return torch.from_numpy(orc.propagate(tp.numpy(), val.numpy(), W.numpy(), x.numpy(), None, x.shape[0]))

    @staticmethod
    def backward(ctx, g):
        x, W, tp, val = ctx.saved_tensors
        gx, gw = orc.propagate_backward(tp.numpy(), val.numpy(), W.numpy(), g.numpy(), x.numpy())
        return torch.from_numpy(gx), torch.from_numpy(gw), None, None


def _worker(rank, world, port, out_dir):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from torch_rgcn_b200.parallel import plan_relation_shards, partition_edges, _CopyToShards, _ReduceFromShards
    N, R = 60, 5
    rng = np.random.RandomState(0)
    t = np.stack([rng.randint(0, N, 400), rng.randint(0, R, 400), rng.randint(0, N, 400)], 1)
    tp = orc.add_inverse_and_self(t, N, R)
    Rp = 2 * R + 1
    val = orc.nc_edge_values(tp, N, Rp, False)           # horizontal: weights come from the FULL graph
    W = torch.tensor(rng.randn(Rp, 6, 4))
    x = torch.tensor(rng.randn(N, 6), requires_grad=True)
    G = torch.tensor(rng.randn(N, 4))
    owner = plan_relation_shards(np.bincount(tp[:, 1], minlength=Rp), world)
    mask = partition_edges(torch.from_numpy(tp), owner, rank).numpy()
    Wl = W.clone().requires_grad_(True)
    xs = _CopyToShards.apply(x, None)
    part = _OracleShard.apply(xs, Wl, torch.from_numpy(tp[mask]), torch.from_numpy(val[mask]))
    out = _ReduceFromShards.apply(part, None, None)
    out.backward(G)
    dist.all_reduce(Wl.grad)                              # disjoint supports -> sum is the full gradient
    full = orc.propagate(tp, val, W.numpy(), x.detach().numpy(), None, N)
    gx, gw = orc.propagate_backward(tp, val, W.numpy(), G.numpy(), x.detach().numpy())
    np.testing.assert_allclose(out.detach().numpy(), full, atol=1e-10)
    np.testing.assert_allclose(x.grad.numpy(), gx, atol=1e-10)
    np.testing.assert_allclose(Wl.grad.numpy(), gw, atol=1e-10)
    assert mask.sum() > 0 and mask.sum() < len(tp)
    gathered = [torch.zeros(len(tp), dtype=torch.bool) for _ in range(world)]
    dist.all_gather(gathered, torch.from_numpy(mask))
    assert torch.stack(gathered).sum(0).eq(1).all()       # every edge on exactly one rank
    open(os.path.join(out_dir, f'ok{rank}'), 'w').write('ok')
    dist.destroy_process_group()


def test_relation_sharding_world2_gloo(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    assert all(os.path.exists(tmp_path / f'ok{r}') for r in range(world))
