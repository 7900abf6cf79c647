"""world_size-2 gloo test of the relation-sharding host logic (planner, edge partition, autograd collectives).

CPU only: the per-rank compute is stood in for by the oracle (tests may use it); on GPUs the same host code
drives the CUDA engine (tests/test_gpu_parity.py::test_relation_sharded_* under torchrun-less single GPU,
bench.py --gpus N for real).
"""
import os
import socket

import numpy as np
import pytest
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


# ---- row sharding (torch_rgcn_b200.parallel.RowShardedNC's host logic) ---------------------------------------
class _OracleRowShard:
    """What _RowShardedApply needs from a shard, with the oracle standing in for the CUDA engine: forward over the
    edges INTO this rank's rows, backward over the edges OUT OF this rank's rows."""

    def __init__(self, tp, val, N, rank, world, group=None):
        from torch_rgcn_b200.parallel import plan_row_shards, partition_edges_by_rows
        self.num_nodes, self.group = N, group
        self.rows_per, ranges = plan_row_shards(N, world)
        self.lo, self.hi = ranges[rank]
        t = torch.from_numpy(tp)
        self.fm = partition_edges_by_rows(t, 0, self.lo, self.hi).numpy()
        self.bm = partition_edges_by_rows(t, 2, self.lo, self.hi).numpy()
        self.tp, self.val = tp, val
        self.out_comm_dtype = self.grad_comm_dtype = None
        self.param_names = ['weights', 'bias']

    def forward_local(self, features, params):
        W, bias = params
        return torch.from_numpy(orc.propagate(self.tp[self.fm], self.val[self.fm], W.numpy(), features.numpy(),
                                              bias.numpy(), self.num_nodes))

    def backward_local(self, features, params, grad_out, need_features, need_params):
        W, bias = params
        gx, gw = orc.propagate_backward(self.tp[self.bm], self.val[self.bm], W.numpy(), grad_out.numpy(), features.numpy())
        return torch.from_numpy(gx), [torch.from_numpy(gw), grad_out.sum(0)]


def _row_worker(rank, world, port, out_dir):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from torch_rgcn_b200.parallel import _RowShardedApply
    N, R = 61, 5                                         # 61 rows over 2 or 3 ranks: a ragged last block
    rng = np.random.RandomState(1)
    t = np.stack([rng.randint(0, N, 500), rng.randint(0, R, 500), rng.randint(0, N, 500)], 1)
    tp = orc.add_inverse_and_self(t, N, R)
    Rp = 2 * R + 1
    val = orc.nc_edge_values(tp, N, Rp, False)
    W = torch.tensor(rng.randn(Rp, 6, 4), requires_grad=True)
    bias = torch.tensor(rng.randn(4), requires_grad=True)
    x = torch.tensor(rng.randn(N, 6), requires_grad=True)
    G = torch.tensor(rng.randn(N, 4))
    shard = _OracleRowShard(tp, val, N, rank, world)
    out = _RowShardedApply.apply(shard, x, W, bias)
    out.backward(G)
    full = orc.propagate(tp, val, W.detach().numpy(), x.detach().numpy(), bias.detach().numpy(), N)
    gx, gw = orc.propagate_backward(tp, val, W.detach().numpy(), G.numpy(), x.detach().numpy())
    np.testing.assert_allclose(out.detach().numpy(), full, atol=1e-10)          # every row from exactly one rank
    np.testing.assert_allclose(x.grad.numpy(), gx, atol=1e-10)
    np.testing.assert_allclose(W.grad.numpy(), gw, atol=1e-10)                  # all-reduced inside backward
    np.testing.assert_allclose(bias.grad.numpy(), G.sum(0).numpy(), atol=1e-10)  # NOT multiplied by the world size
    masks = [torch.zeros(len(tp), dtype=torch.bool) for _ in range(world)]
    dist.all_gather(masks, torch.from_numpy(shard.fm))
    assert torch.stack(masks).sum(0).eq(1).all()                                 # every edge in one forward shard
    dist.all_gather(masks, torch.from_numpy(shard.bm))
    assert torch.stack(masks).sum(0).eq(1).all()                                 # ... and in one backward shard
    open(os.path.join(out_dir, f'row_ok{rank}'), 'w').write('ok')
    dist.destroy_process_group()


@pytest.mark.parametrize('world', [2, 3])
def test_row_sharding_gloo(tmp_path, world):
    mp.spawn(_row_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    assert all(os.path.exists(tmp_path / f'row_ok{r}') for r in range(world))
