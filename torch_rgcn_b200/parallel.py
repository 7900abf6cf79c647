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
from .layers import _fuse_dirs


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
                    fuse_item_tiles=int(os.environ.get('RGCN_FUSE_ITEM_TILES', '4096')), fuse_dirs=_fuse_dirs(self.layer.out_features or 64))
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


# ------------------------------------------------------------------------------------------------------
# Row sharding (experimental; host logic covered by a gloo test, GPU parity through tests/sharded_check.py with
# SHARD=rows — not yet run on B200s).
#
# Relation sharding all-reduces a full (N, O) output per direction whatever the world size.  Sharding ROWS instead
# needs no reduction of node data: the forward partitions edges by destination row (rank k produces complete output
# rows [lo_k, hi_k) from ALL relations), the backward partitions them by source row (rank k produces complete
# feature-gradient rows), and the row blocks are all-gathered — half the bytes of an all-reduce, and exact (every row
# comes from one rank).  Only the parameter gradients (small) are all-reduced.
# ------------------------------------------------------------------------------------------------------
def plan_row_shards(num_nodes, world):
    """Equal row blocks (all_gather needs equal chunks): rows_per = ceil(N / world); rank k owns
    [k * rows_per, min(N, (k + 1) * rows_per))."""
    rows_per = (int(num_nodes) + int(world) - 1) // int(world)
    return rows_per, [(min(k * rows_per, num_nodes), min((k + 1) * rows_per, num_nodes)) for k in range(world)]


def partition_edges_by_rows(triples_plus, column, lo, hi):
    """Mask of the rows of `triples_plus` whose `column` entry (0 = destination s, 2 = source o) lies in [lo, hi)."""
    c = triples_plus[:, column]
    return (c >= lo) & (c < hi)


def gather_row_blocks(x, num_nodes, rows_per, lo, hi, group=None, comm_dtype=None):
    """x (N, d) with rows [lo, hi) valid on this rank -> (N, d) with every rank's rows, by ONE all-gather."""
    world = dist.get_world_size(group)
    d = x.size(1)
    dt = comm_dtype if comm_dtype is not None else x.dtype
    mine = torch.zeros(rows_per, d, dtype=dt, device=x.device)
    mine[: hi - lo] = x[lo:hi].to(dt)
    full = torch.empty(world * rows_per, d, dtype=dt, device=x.device)
    dist.all_gather_into_tensor(full, mine, group=group)
    return full[:num_nodes].to(x.dtype)


class _RowShardedApply(torch.autograd.Function):
    """forward: local rows by `shard.forward_local`, all-gather; backward: local feature-gradient rows by
    `shard.backward_local`, all-gather, all-reduce of the parameter gradients (bias excluded: every rank computes it
    from the replicated upstream gradient)."""

    @staticmethod
    def forward(ctx, shard, features, *params):
        out = shard.forward_local(features, params)
        ctx.shard = shard
        ctx.n_params = len(params)
        ctx.save_for_backward(features, *params)
        make = getattr(shard, '_generic_exchange', None)      # stand-in shards (CPU tests) exchange through the group
        ex = make(out.device, out.size(1), features) if make is not None else None
        if ex is not None:
            return ex.gather_rows(0, out, shard.lo, shard.hi, shard.num_nodes, out.dtype)
        return gather_row_blocks(out, shard.num_nodes, shard.rows_per, shard.lo, shard.hi, shard.group, shard.out_comm_dtype)

    @staticmethod
    def backward(ctx, grad_out):
        shard = ctx.shard
        features, *params = ctx.saved_tensors
        need = ctx.needs_input_grad                       # (shard, features, *params)
        g_feat, g_params = shard.backward_local(features, params, grad_out.contiguous(), need[1], need[2:])
        ex = getattr(shard, '_gx', None)
        if g_feat is not None and ex is not None:
            g_feat = ex.gather_rows(1, g_feat, shard.lo, shard.hi, shard.num_nodes, g_feat.dtype)
        elif g_feat is not None:
            g_feat = gather_row_blocks(g_feat, shard.num_nodes, shard.rows_per, shard.lo, shard.hi, shard.group,
                                       shard.grad_comm_dtype)
        for name, g in zip(shard.param_names, g_params):
            if g is not None and name != 'bias':
                dist.all_reduce(g, op=dist.ReduceOp.SUM, group=shard.group)
        return (None, g_feat, *g_params)


class _Exchange:
    """Symmetric (rows, 64) bf16 buffers of a row-sharded layer: allocated with torch's symmetric-memory allocator and
    mapped on every rank of the NVLink domain, so a rank can store (kernel) or copy (copy engine) its rows straight
    into every peer's buffer.  Two buffers per direction, used alternately: a rank that is one step ahead writes the
    other buffer, and the barrier that ends every exchange keeps it from getting two steps ahead, so no barrier is
    needed before the writes.  slot(direction) -> (local buffer, handle, peer views, peer pointers)."""

    def __init__(self, rows, device, group, widths=(64, 64), dtype=torch.bfloat16):
        import torch.distributed._symmetric_memory as symm
        pg = group if group is not None else dist.group.WORLD
        self.rows, self.widths, self.dtype = rows, tuple(widths), dtype
        self.slots = [[], []]                    # [direction][parity]
        self.step = [0, 0]
        for direction in range(2):
            w = self.widths[direction]
            for _ in range(2):
                t = symm.empty(rows, w, dtype=dtype, device=device)
                h = symm.rendezvous(t, pg)
                peers = [h.get_buffer(q, (rows, w), dtype) for q in range(h.world_size)]
                self.slots[direction].append((t, h, peers, [int(x) for x in h.buffer_ptrs]))

    def slot(self, direction):
        k = self.step[direction]
        self.step[direction] = k + 1
        return self.slots[direction][k & 1]

    def gather_rows(self, direction, x, lo, hi, num_nodes, out_dtype):
        """x (N, d) with rows [lo, hi) valid on this rank -> (N, d) with every rank's rows: the copy engines push the
        rank's rows (in the exchange dtype) into every rank's buffer over NVLink, one barrier, one widening copy."""
        buf, h, peers, _ptrs = self.slot(direction)
        W, r = h.world_size, h.rank
        with torch.cuda.device(x.device):
            if lo < hi:
                mine = x[lo:hi] if x.dtype == self.dtype else x[lo:hi].to(self.dtype)
                for q in range(W):
                    peers[(r + q) % W][lo:hi].copy_(mine, non_blocking=True)
            h.barrier(channel=0)
        return buf[:num_nodes].to(out_dtype, copy=True)        # always a copy: the buffer is reused two steps later

    @staticmethod
    def create(rows, device, group, widths=(64, 64), dtype=torch.bfloat16):
        """None when symmetric memory cannot be set up (no peer access, a CPU group): callers use NCCL instead."""
        if os.environ.get('RGCN_SHARD_COMM', 'symm') != 'symm' or dist.get_world_size(group) > 8 or \
                dist.get_backend(group) != 'nccl':
            return None
        try:
            return _Exchange(rows, device, group, widths, dtype)
        except Exception as exc:  # noqa: BLE001
            if dist.get_rank(group) == 0:
                print(f'torch_rgcn_b200: symmetric memory unavailable ({type(exc).__name__}: {str(exc)[:200]}); '
                      f'row-sharded layers fall back to NCCL all-gathers', flush=True)
            return None


class _RowShardedFused(torch.autograd.Function):
    """Row-sharded layer on the fused row-block path (bf16 features, 64 -> 64, four blocks).

    Rows are owned in `chunks` interleaved pieces per rank (piece k of rank r = row blocks [(k W + r) cb, (k W + r + 1) cb)),
    so that piece k of all ranks is one contiguous slab of the bf16 exchange buffer.  forward: for every piece the fused
    kernel writes this rank's rows straight into the slab (bf16 output, row range) and an asynchronous in-place
    all-gather of the slab starts at once: the NVLink transfer of piece k overlaps the gather kernel of piece k + 1.
    backward: the two-phase kernels over the edges whose SOURCE row this rank owns give complete feature-gradient rows
    (bf16) and partial parameter gradients; the rows are all-gathered slab by slab, the parameter gradients all-reduced.
    """

    @staticmethod
    def forward(ctx, shard, features, blocks, bias):
        import ctypes as C
        from .functional import _params_struct, _f32c
        N, W, r, cb, H, nch = shard.num_nodes, shard.world, shard.rank, shard.chunk_blocks, shard.block_rows, shard.chunks
        dev = features.device
        slab = W * cb * H
        buf = torch.empty(nch * slab, 64, dtype=torch.bfloat16, device=dev)
        feats = features.contiguous()
        wb, bs = _f32c(blocks), _f32c(bias)
        plan = shard._plan_f
        p = _params_struct('block', False, 64, 64, None, None, None, wb, None, bs, None)
        p.out_dtype = _lib.BF16
        ws_bytes = _lib.lib.rgcn_forward_workspace_bytes(C.byref(plan.c), C.byref(p), _lib.BF16)
        ws = torch.empty(max(ws_bytes, 1), dtype=torch.uint8, device=dev)
        ex = shard._exchange
        if ex is not None:
            # ONE kernel: the flush stores every output row of this rank to all ranks' exchange buffers over NVLink
            # (peer-to-peer stores, rgcn_params.peer_out); the barriers order it against the peers' reads and writes
            buf, h, _peers, ptrs = ex.slot(0)
            lo = r * cb * H
            out = torch.empty(N, 64, dtype=torch.float32, device=dev)
            with torch.cuda.device(dev):
                if lo < N:
                    p.row_lo, p.row_hi = lo, min(lo + cb * H, N)
                    p.num_peer_out = W
                    for q in range(W):
                        p.peer_out[q] = ptrs[q]
                    _lib.check(_lib.lib.rgcn_forward(C.byref(plan.c), C.byref(p), _lib.ptr(feats), _lib.BF16,
                                                     _lib.ptr(buf), _lib.ptr(ws), ws_bytes, _lib.stream_ptr()))
                h.barrier(channel=0)                     # every rank's rows have arrived
                _lib.check(_lib.lib.rgcn_widen_rows(_lib.ptr(buf), N * 64, _lib.ptr(out), _lib.stream_ptr()))
            ctx.shard = shard
            ctx.save_for_backward(feats, blocks, bias)
            return out
        works = []
        with torch.cuda.device(dev):
            for k in range(nch):
                lo = (k * W + r) * cb * H
                if lo < N:
                    p.row_lo, p.row_hi = lo, min(lo + cb * H, N)
                    _lib.check(_lib.lib.rgcn_forward(C.byref(plan.c), C.byref(p), _lib.ptr(feats), _lib.BF16, _lib.ptr(buf),
                                                     _lib.ptr(ws), ws_bytes, _lib.stream_ptr()))
                works.append(dist.all_gather_into_tensor(buf[k * slab:(k + 1) * slab], buf[lo:lo + cb * H],
                                                         group=shard.group, async_op=True))
        for w in works:
            w.wait()
        ctx.shard = shard
        ctx.save_for_backward(feats, blocks, bias)
        return buf[:N].float()

    @staticmethod
    def backward(ctx, grad_out):
        shard = ctx.shard
        feats, blocks, bias = ctx.saved_tensors
        N, W, r, cb, H, nch = shard.num_nodes, shard.world, shard.rank, shard.chunk_blocks, shard.block_rows, shard.chunks
        need = ctx.needs_input_grad                      # (shard, features, blocks, bias)
        shard.param_names = ['blocks', 'bias']
        own = (r * cb * H, min((r + 1) * cb * H, N)) if (shard._exchange is not None and r * cb * H < N) else None
        g_feat, (g_blocks, g_bias) = shard.backward_local(feats, [blocks, bias], grad_out.contiguous(), need[1],
                                                          [need[2], need[3]], rows=own)
        works = []
        if g_blocks is not None:
            works.append(dist.all_reduce(g_blocks, op=dist.ReduceOp.SUM, group=shard.group, async_op=True))
        gx = None
        ex = shard._exchange
        if g_feat is not None and ex is not None:
            # copy engines push this rank's feature-gradient rows into every rank's exchange buffer (no SMs involved)
            buf, h, peers, _ptrs = ex.slot(1)
            lo, hi = r * cb * H, min((r + 1) * cb * H, N)
            with torch.cuda.device(g_feat.device):
                if lo < N:
                    for q in range(W):
                        peers[(r + q) % W][lo:hi].copy_(g_feat[lo:hi], non_blocking=True)
                h.barrier(channel=0)
            gx = buf[:N].clone()
        elif g_feat is not None:
            slab = W * cb * H
            buf = torch.empty(nch * slab, 64, dtype=g_feat.dtype, device=g_feat.device)
            for k in range(nch):
                lo = (k * W + r) * cb * H
                hi = min(lo + cb * H, N)
                if lo < N:
                    buf[lo:hi].copy_(g_feat[lo:hi])
                works.append(dist.all_gather_into_tensor(buf[k * slab:(k + 1) * slab], buf[lo:lo + cb * H],
                                                         group=shard.group, async_op=True))
            gx = buf[:N]
        for w in works:
            w.wait()
        return None, gx, g_blocks, g_bias


class RowShardedNC(torch.nn.Module):
    """Runs a RelationalGraphConvolutionNC with the OUTPUT ROWS sharded over the ranks (see the block comment above).

    Same contract as RelationShardedNC: every rank constructs the same layer and passes the same features; the output
    and, after backward, all gradients are identical on all ranks (no sync_parameter_grads needed).  Layers the
    fused row-block kernel serves (bf16 features, 64 -> 64, four blocks) take the overlapped path of
    _RowShardedFused; every other layer the generic path (_RowShardedApply)."""

    def __init__(self, layer, group=None, chunks=None):
        super().__init__()
        self.layer = layer
        self.group = group
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)
        self.num_nodes = layer.num_nodes
        self.rows_per, ranges = plan_row_shards(layer.num_nodes, self.world)
        self.lo, self.hi = ranges[self.rank]
        self.out_comm_dtype = self.grad_comm_dtype = None
        self.chunks = int(chunks if chunks is not None else os.environ.get('RGCN_SHARD_CHUNKS', '2'))
        self._plans = None
        self._exchange = None
        self._gx = None                                   # exchange buffers of the generic path
        self._gx_key = None

    def sync_parameter_grads(self):
        """Nothing to do (kept for interface parity with RelationShardedNC)."""

    def _generic_exchange(self, device, out_width, features):
        """Symmetric-memory exchange buffers of the generic path (bf16 layers with features: rows travel as bf16), or None:
        then the rows go through NCCL all-gathers."""
        if features is None or features.dtype != torch.bfloat16:
            return None
        key = (str(device), out_width, features.size(1))
        if self._gx_key != key:
            self._gx_key = key
            self._gx = _Exchange.create(self.num_nodes, device, self.group, (out_width, features.size(1)), torch.bfloat16)
        return self._gx

    def _fused_rows(self, features):
        """Block height of the fused row-block kernel if this layer takes the overlapped 64-wide path, else 0."""
        return self.layer._fuse_rows(features) if self.layer.out_features == 64 else 0

    def row_owner_mask(self, rows, block_rows, chunk_blocks):
        """Rows owned by this rank under the interleaved-piece ownership of _RowShardedFused."""
        return ((rows // block_rows) // chunk_blocks) % self.world == self.rank

    def _local_plans(self, device, features):
        L = self.layer
        tile_edges = L._tile_edges(features)
        H = self._fused_rows(features)
        kw = dict(tile_edges=tile_edges, ring_depth=int(os.environ.get('RGCN_RING_DEPTH', '8')))
        key = (str(device), tile_edges, H, self.chunks)
        if self._plans is None or self._plans[0] != key:
            tp = L.triples.to(device)
            full = L._plan(device)                               # per-edge weights of the FULL graph, caller order
            val = full.val[:full.nnz]
            if H > 0:
                nblocks = (L.num_nodes + H - 1) // H
                self.block_rows = H
                # symmetric memory: one contiguous piece per rank, exchanged by the kernel itself / the copy engines;
                # otherwise `chunks` interleaved pieces whose NCCL all-gathers overlap the next piece's kernel
                cb1 = (nblocks + self.world - 1) // self.world
                self._exchange = _Exchange.create(self.world * cb1 * H, device, self.group)
                if self._exchange is not None:
                    self.chunks = 1
                self.chunk_blocks = (nblocks + self.world * self.chunks - 1) // (self.world * self.chunks)
                fwd_mask = self.row_owner_mask(tp[:, 0], H, self.chunk_blocks)
                bwd_mask = self.row_owner_mask(tp[:, 2], H, self.chunk_blocks)
                fkw = dict(fuse_rows=H, fuse_item_tiles=1 << 20, fuse_dirs=1)      # unsplit blocks: row ranges = item ranges
                bkw = {}
            else:
                fwd_mask = partition_edges_by_rows(tp, 0, self.lo, self.hi)
                bwd_mask = partition_edges_by_rows(tp, 2, self.lo, self.hi)
                fkw = bkw = {}
                Hw = L._fuse_rows(features)                      # wider layers: fused kernels inside the generic path
                if Hw > 0:
                    item_tiles = int(os.environ.get('RGCN_FUSE_ITEM_TILES', '4096'))
                    fkw = dict(fuse_rows=Hw, fuse_item_tiles=item_tiles, fuse_dirs=1)
                    if _fuse_dirs(L.out_features) == 3:
                        bkw = dict(fuse_rows=Hw, fuse_item_tiles=item_tiles, fuse_dirs=2)
            plan_f = GraphPlan(tp[fwd_mask], L.num_nodes, L.num_relations, _lib.NORM_EXPLICIT, val=val[fwd_mask],
                               validate=False, **kw, **fkw)
            plan_b = GraphPlan(tp[bwd_mask], L.num_nodes, L.num_relations, _lib.NORM_EXPLICIT, val=val[bwd_mask],
                               validate=False, **kw, **bkw)
            L._plan_cache = None
            self._plans = (key, plan_f, plan_b)
        return self._plans[1], self._plans[2]

    # -- what _RowShardedApply calls ------------------------------------------------------------------
    def _named(self, params):
        d = dict(weights=None, bases=None, comps=None, blocks=None, bias=None)
        d.update(dict(zip(self.param_names, params)))
        return d

    def forward_local(self, features, params):
        p = self._named(params)
        with torch.no_grad():
            return _Propagate.apply(self._plan_f, self._form, self._in_dim, self.layer.out_features, features,
                                    p['weights'], p['bases'], p['comps'], p['blocks'], None, p['bias'], None, True)

    def backward_local(self, features, params, grad_out, need_features, need_params, rows=None):
        """The engine's backward over the source-row shard, through _Propagate.backward with a stand-in context (the
        engine keeps no forward state: it needs the plan, the inputs and the upstream gradient only)."""
        from types import SimpleNamespace
        from .functional import _f32c
        p = self._named(params)
        needs = dict(zip(self.param_names, need_params))
        if features is not None:                                 # what _Propagate.forward does before saving its inputs
            if features.dtype not in (torch.float32, torch.bfloat16):
                features = features.float()
            features = features.contiguous()
        saved = tuple(_f32c(p[k]) for k in ('weights', 'bases', 'comps', 'blocks')) + (None, _f32c(p['bias']), None)
        ctx = SimpleNamespace(
            saved_tensors=(features,) + saved, plan=self._plan_b, form=self._form,
            dims=(self._in_dim, self.layer.out_features),
            in_dtypes=[None if t is None else t.dtype
                       for t in (features, p['weights'], p['bases'], p['comps'], p['blocks'], None, p['bias'])],
            needs_input_grad=(False, False, False, False, bool(need_features and features is not None),
                              needs.get('weights', False), needs.get('bases', False), needs.get('comps', False),
                              needs.get('blocks', False), False, needs.get('bias', False), False, False))
        ctx.rows = rows                                          # (lo, hi): only these feature-gradient rows are needed
        grads = _Propagate.backward(ctx, grad_out)
        by_name = dict(weights=grads[5], bases=grads[6], comps=grads[7], blocks=grads[8], bias=grads[10])
        return grads[4], [by_name[n] for n in self.param_names]

    def forward(self, features=None):
        L = self.layer
        assert (features is None) == (L.in_features is None), "in_features not provided!"
        lead = L._decomposed()[0]
        self._plan_f, self._plan_b = self._local_plans(lead.device, features)
        self._in_dim = L.in_features if L.in_features is not None else L.num_nodes
        if L.diag_weight_matrix:
            self._form, names = 'diag', ['weights']
        elif L.weight_decomp is None:
            self._form, names = 'dense', ['weights']
        elif L.weight_decomp == 'basis':
            self._form, names = 'basis', ['bases', 'comps']
        else:
            self._form, names = 'block', ['blocks']
        if L.bias is not None:
            names = names + ['bias']
        self.param_names = names
        bf16 = features is not None and features.dtype == torch.bfloat16
        if self._fused_rows(features) > 0 and self._plan_f.fused_ok[0]:
            return _RowShardedFused.apply(self, features, L.blocks, L.bias)
        self.out_comm_dtype = torch.bfloat16 if bf16 else None       # like RelationShardedNC: bf16 layers send bf16
        self.grad_comm_dtype = None                                  # the feature gradient already has the feature dtype
        return _RowShardedApply.apply(self, features, *[getattr(L, n) for n in names])
