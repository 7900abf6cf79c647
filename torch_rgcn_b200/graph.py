"""GraphPlan: the sorted edge lists + per-edge weights the CUDA kernels walk.

Replaces what the reference rebuilds on every forward — stack_matrices, sum_sparse, the
'transpose trick' permutation and the sparse COO constructor (reference torch_rgcn/layers.py:255-279,
:490-516) — with one device-side build.  torch is used only to own the device memory.
"""
import ctypes as C

import torch

from . import _lib


class GraphPlan:
    """Device-resident plan for one edge list `triples_plus` (nnz, 3) int64.

    norm: _lib.NORM_ROW (vertical stacking), _lib.NORM_COL_SWAPPED (horizontal stacking; needs the
    reference's (n, i) = (n_general, n_self)) or _lib.NORM_EXPLICIT (caller-provided `val`).
    """

    def __init__(self, triples_plus, num_nodes, num_rels, norm, n_general=0, n_self=0, val=None, validate=True,
                 tile_edges=0, ring_depth=8, fuse_rows=0, fuse_item_tiles=4096, fuse_dirs=1):
        _lib.require_cuda(triples_plus)
        assert triples_plus.dtype == torch.long, 'triples must be torch.long'   # reference utils.py:148
        t = triples_plus.contiguous()
        dev = t.device
        nnz = int(t.size(0))
        self.num_nodes, self.num_rels, self.nnz, self.device = int(num_nodes), int(num_rels), nnz, dev
        i32 = dict(dtype=torch.int32, device=dev)
        f32 = dict(dtype=torch.float32, device=dev)
        n1 = max(nnz, 1)
        self.d_rowptr = torch.empty(num_nodes + 1, **i32)
        self.s_rowptr = torch.empty(num_nodes + 1, **i32)
        self.r_relptr = torch.empty(num_rels + 1, **i32)
        # one allocation each for the int and float edge arrays
        self._ints = torch.empty(8, n1, **i32)
        self._floats = torch.empty(4, n1, **f32)
        (self.d_src, self.d_rel, self.s_dst, self.s_rel, self.r_dst, self.r_src, self.r_dslot,
         self.r_sslot) = self._ints.unbind(0)
        self.r_chunkptr = torch.empty(num_rels + 1, **i32)
        self.d_val, self.s_val, self.r_val, self.val = self._floats.unbind(0)
        self.status = torch.zeros(8, **i32)
        self.d_long = torch.empty(nnz // 512 + 1, **i32)
        self.s_long = torch.empty(nnz // 512 + 1, **i32)
        g = _lib.Graph()
        g.num_nodes, g.num_rels, g.nnz = num_nodes, num_rels, nnz
        for name in ('d_rowptr', 'd_src', 'd_rel', 'd_val', 's_rowptr', 's_dst', 's_rel', 's_val',
                     'r_relptr', 'r_dst', 'r_src', 'r_val', 'r_dslot', 'r_sslot', 'r_chunkptr', 'val', 'status',
                     'd_long', 's_long'):
            setattr(g, name, getattr(self, name).data_ptr())
        # optional super-tiling for the L2-resident message ring (see include/rgcn_b200.h: rgcn_tiling)
        self.tile_edges = int(tile_edges) if nnz > 0 else 0
        g.tile_edges = self.tile_edges
        g.ring_depth = self.ring_depth = max(2, min(int(ring_depth), 64))
        self._tiling = []
        if self.tile_edges > 0:
            T = (nnz - 1) // self.tile_edges + 1
            n_items = _lib.lib.rgcn_tile_items_bound(nnz, num_nodes, num_rels, self.tile_edges)
            for tl in (g.ft, g.bt):
                arrs = dict(tilerow=torch.empty(T + 1, **i32), row=torch.empty(n1, **i32), col=torch.empty(n1, **i32),
                            rel=torch.empty(n1, **i32), slot=torch.empty(n1, **i32), val=torch.empty(n1, **f32),
                            stepptr=torch.empty(_lib.lib.rgcn_tile_steps_len(nnz, self.tile_edges, self.ring_depth), **i32),
                            slotneed=torch.empty(T, **i32), items=torch.empty(n_items, 8, **i32))
                self._tiling.append(arrs)
                for k, v in arrs.items():
                    setattr(tl, k, v.data_ptr())
        # optional fused row-block lists (see include/rgcn_b200.h: rgcn_fused): the edges of every block of
        # `fuse_rows` rows sorted by (relation, row parity, row), runs dealt over 16-entry tiles; capacity 2 * nnz entries
        self.fuse_rows = int(fuse_rows) if nnz > 0 else 0
        g.fuse_rows = self.fuse_rows
        self._fused = []
        if self.fuse_rows > 0:
            assert self.fuse_rows % 16 == 0 and 16 <= self.fuse_rows <= 4096
            cap = (2 * nnz + 15) // 16 * 16 + 16 * 1024
            g.fuse_cap, g.fuse_item_tiles = cap, max(1, min(int(fuse_item_tiles), 1 << 20))
            NB = (num_nodes + self.fuse_rows - 1) // self.fuse_rows
            n_items = _lib.lib.rgcn_fused_items_bound(num_nodes, self.fuse_rows, cap, g.fuse_item_tiles)
            g.fuse_dirs = self.fuse_dirs = int(fuse_dirs)        # bit 0: forward lists, bit 1: feature-gradient lists
            assert self.fuse_dirs in (1, 2, 3)
            for d, fl in enumerate((g.ff, g.fb)):
                if not (self.fuse_dirs >> d) & 1:
                    self._fused.append(None)
                    continue
                arrs = dict(col=torch.empty(cap, **i32), rec=torch.empty(cap // 16, _lib.FUSE_REC_WORDS, **i32),
                            blk_tile=torch.empty(NB + 1, **i32), items=torch.empty(n_items, 4, **i32),
                            meta=torch.zeros(8, **i32))
                self._fused.append(arrs)
                for k, v in arrs.items():
                    setattr(fl, k, v.data_ptr())
        g.num_long_dst = g.num_long_src = -1
        g.max_rel_edges = 0
        self.c = g
        if val is not None:
            val = val.to(device=dev, dtype=torch.float32).contiguous()
            assert val.numel() == nnz
        ws_bytes = _lib.lib.rgcn_graph_workspace_bytes(nnz, num_nodes, num_rels, self.tile_edges, self.fuse_rows)
        ws = torch.empty(max(ws_bytes, 1), dtype=torch.uint8, device=dev)
        with torch.cuda.device(dev):
            _lib.check(_lib.lib.rgcn_graph_build(_lib.ptr(t), nnz, num_nodes, num_rels, norm, int(n_general),
                                                 int(n_self), _lib.ptr(val), C.byref(g), _lib.ptr(ws), ws_bytes,
                                                 _lib.stream_ptr()))
        self.fused_ok = [False, False]
        self.fused_flagged = [0, 0]
        if self.fuse_rows > 0:                       # host copies of the list sizes; an overflowing list is not used
            for d, arrs in enumerate(self._fused):
                if arrs is None:
                    continue
                items, tiles, overflow, split, flagged = arrs['meta'].tolist()[:5]
                self.fused_ok[d] = overflow == 0 and items > 0
                g.fuse_items[d] = items if self.fused_ok[d] else 0
                g.fuse_split[d] = split
                g.fuse_tiles[d] = tiles
                self.fused_flagged[d] = flagged
        if self.tile_edges > 0 or validate:
            st = self.status.tolist()               # one host sync per plan build (the reference's asserts sync too)
            g.tile_capacity = self.tile_capacity = max(st[1], st[2])
            g.num_long_dst, g.num_long_src = st[4], st[5]
            g.max_rel_edges = st[6]
            bad = st[0]
            assert bad == 0 or not validate, f'{bad} triples have a node or relation id out of range ' \
                                             f'(num_nodes={num_nodes}, num_relations={num_rels})'

    def relation_counts(self):
        """Edges per relation (host tensor), e.g. for the relation shard planner."""
        return (self.r_relptr[1:] - self.r_relptr[:-1]).to(torch.int64).cpu()
