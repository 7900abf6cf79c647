"""Relation-sharded multi-GPU execution (one process per GPU, torch.distributed / NCCL).

The reference is single-device.  The sum over relations in  out[s] = bias + sum_p sum_{e in p} val_e X[o] W_p
is linear, so relations partition across ranks: rank k holds the edges (and reads only the weights) of the
relations it owns, computes a partial (N, O) output, and ONE all-reduce(sum) of that output finishes the
forward; the backward needs one all-reduce of the feature gradient.  Weight gradients of owned relations are
complete locally; non-owned relations get exact zeros.

Host logic here (planner, edge partition, the two autograd collectives) is device-agnostic and is covered on
CPU with a world_size-2 gloo test; the per-rank compute is the CUDA engine.
"""
import ctypes
import os

import numpy as np
import torch
import torch.distributed as dist

from . import _lib
from .functional import _Propagate
from .graph import GraphPlan


def plan_relation_shards(rel_counts, world):
    """relation -> rank by longest-processing-time bin packing on edge counts (rgcn_shard_plan, host code)."""
    counts = np.ascontiguousarray(np.asarray(rel_counts, dtype=np.int64))
    owner = np.empty(counts.shape[0], dtype=np.int32)
    _lib.check(_lib.lib.rgcn_shard_plan(counts.ctypes.data_as(ctypes.c_void_p), counts.shape[0], int(world),
                                        owner.ctypes.data_as(ctypes.c_void_p)))
    return torch.from_numpy(owner)


def partition_edges(triples_plus, rel_to_rank, rank):
    """Boolean mask of the rows of `triples_plus` whose relation is owned by `rank`."""
    owner = rel_to_rank.to(triples_plus.device)
    return owner[triples_plus[:, 1]] == rank


class _CopyToShards(torch.autograd.Function):
    """Identity forward; all-reduce(sum) of the gradient (every rank holds a partial feature gradient)."""

    @staticmethod
    def forward(ctx, x, group):
        ctx.group = group
        return x.view_as(x)

    @staticmethod
    def backward(ctx, g):
        g = g.contiguous()
        dist.all_reduce(g, op=dist.ReduceOp.SUM, group=ctx.group)
        return g, None


class _ReduceFromShards(torch.autograd.Function):
    """All-reduce(sum) forward; identity backward (the upstream gradient is replicated).

    comm_dtype=torch.bfloat16 sends the partial sums in bf16 (half the NVLink bytes); used for bf16-feature layers,
    whose per-edge messages are bf16 already.  The result is returned in the input dtype.
    """

    @staticmethod
    def forward(ctx, x, group, comm_dtype=None):
        if comm_dtype is not None and comm_dtype != x.dtype:
            y = x.to(comm_dtype)
            dist.all_reduce(y, op=dist.ReduceOp.SUM, group=group)
            return y.to(x.dtype)
        x = x.contiguous()
        dist.all_reduce(x, op=dist.ReduceOp.SUM, group=group)
        return x

    @staticmethod
    def backward(ctx, g):
        return g, None, None


class RelationShardedNC(torch.nn.Module):
    """Runs a RelationalGraphConvolutionNC over the relations this rank owns and all-reduces the result.

    Every rank constructs the same layer (same seed) and passes the same features; the output is identical on
    all ranks.  After backward, `sync_parameter_grads()` sums the per-rank weight gradients (disjoint supports;
    only `bases` genuinely overlaps) so replicas can step identical optimisers.
    """

    def __init__(self, layer, group=None):
        super().__init__()
        self.layer = layer
        self.group = group
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)
        self._local = None

    def _local_plan(self, device, features=None):
        tile_edges = self.layer._tile_edges(features)
        fuse = dict(fuse_rows=self.layer._fuse_rows(features),
                    fuse_item_tiles=int(os.environ.get('RGCN_FUSE_ITEM_TILES', '512')),
                             fuse_order=int(os.environ.get('RGCN_FUSE_ORDER', '1')))
        if (self._local is None or self._local.device != device or self._local.tile_edges != tile_edges or
                self._local.fuse_rows != fuse['fuse_rows']):
            L = self.layer
            tp = L.triples.to(device)
            counts = torch.bincount(tp[:, 1], minlength=L.num_relations).cpu()
            self.rel_to_rank = plan_relation_shards(counts.numpy(), self.world)
            mask = partition_edges(tp, self.rel_to_rank, self.rank)
            if L.vertical_stacking:
                # (p, s) segment counts are relation-local: normalise the shard directly
                self._local = GraphPlan(tp[mask], L.num_nodes, L.num_relations, _lib.NORM_ROW,
                                        validate=L.validate_triples, tile_edges=tile_edges,
                                        ring_depth=int(os.environ.get('RGCN_RING_DEPTH', '8')), **fuse)
            else:
                # the horizontal permutation pairs each edge with its inverse in another relation: take the
                # per-edge weights from the full graph, then keep this rank's rows
                full = L._plan(device)
                self._local = GraphPlan(tp[mask], L.num_nodes, L.num_relations, _lib.NORM_EXPLICIT,
                                        val=full.val[:full.nnz][mask], validate=False, tile_edges=tile_edges,
                                        ring_depth=int(os.environ.get('RGCN_RING_DEPTH', '8')), **fuse)
                L._plan_cache = None
        return self._local

    def forward(self, features=None):
        L = self.layer
        assert (features is None) == (L.in_features is None), "in_features not provided!"
        lead = L._decomposed()[0]
        plan = self._local_plan(lead.device, features)
        in_dim = L.in_features if L.in_features is not None else L.num_nodes
        if features is not None:
            features = _CopyToShards.apply(features, self.group)
        kw = dict(weights=None, bases=None, comps=None, blocks=None)
        if L.diag_weight_matrix:
            form, kw['weights'] = 'diag', L.weights
        elif L.weight_decomp is None:
            form, kw['weights'] = 'dense', L.weights
        elif L.weight_decomp == 'basis':
            form, kw['bases'], kw['comps'] = 'basis', L.bases, L.comps
        else:
            form, kw['blocks'] = 'block', L.blocks
        # bias enters the sum exactly once: rank 0's kernel adds it, every rank gets its (replicated) gradient
        out = _Propagate.apply(plan, form, in_dim, L.out_features, features, kw['weights'], kw['bases'], kw['comps'],
                               kw['blocks'], None, L.bias, None, self.rank == 0)
        comm = torch.bfloat16 if (features is not None and features.dtype == torch.bfloat16) else None
        return _ReduceFromShards.apply(out, self.group, comm)

    def sync_parameter_grads(self):
        for name, p in self.layer.named_parameters():
            if p.grad is not None and name != 'bias':      # the bias gradient is already complete on every rank
                dist.all_reduce(p.grad, op=dist.ReduceOp.SUM, group=self.group)
