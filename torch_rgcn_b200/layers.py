"""Drop-in RGCN layers backed by the B200 CUDA engine.

`RelationalGraphConvolutionNC` / `RelationalGraphConvolutionLP` keep the constructor arguments, parameter
names and shapes, initialisation order, `forward` signatures and error behaviour of the reference classes
(torch_rgcn/layers.py:101-308 and :311-565) so that torch_rgcn/models.py and the experiment scripts can use
them unchanged.  What changes is everything inside `forward`: the sparse-adjacency construction, degree
normalisation, weight decomposition and message passing run as CUDA kernels behind `rgcn_propagate`.
"""
import math
import os

import torch
from torch import nn
from torch.nn import Module, Parameter

from . import _lib
from .functional import rgcn_propagate
from .graph import GraphPlan
from .utils import select_b_init, select_w_init, schlichtkrull_normal_
from .decoder import DistMult                      # noqa: F401  (reference layers.py:9 defines it in this module)

# RGCN_FUSED: '1' (default) routes the forward of bf16 block layers with 16x16 blocks (width 64 .. 512) to the fused
# row-block kernel (propagate_fused.cuh); the backward of a 64-wide layer keeps the two-phase tensor-core kernels
# (propagate_mma.cuh), wider layers also compute the feature gradient with the fused kernel.  '2': fused feature
# gradient for every width.  '0': two-phase / tiled kernels throughout.
_FUSED_DEFAULT = '1'


def _fuse_dirs(width=64):
    """Lists the plan builds for the fused kernel: 1 = forward only, 3 = forward and feature gradient.  By default the
    feature gradient of a 64-wide layer stays with the fused two-phase backward (one gather serves both gradients,
    1.54 vs 1.65 ms at am64); wider layers, whose messages would not fit and go through the tiled ring kernels instead,
    take the fused kernel for the feature gradient too (synthetic 512-wide layer: 141 vs 230 ms)."""
    mode = os.environ.get('RGCN_FUSED', _FUSED_DEFAULT)
    return 3 if (mode == '2' or (mode == '1' and width > 64)) else 1


def _unpack_decomposition(decomposition):
    d = decomposition if decomposition is not None else {}
    return d.get('type'), d.get('num_bases'), d.get('num_blocks')


def _check_blocks(num_blocks, in_dim, out_dim):
    assert num_blocks > 0, \
        'Number of blocks should be set to a value higher than zero for block diagonal decomposition!'
    assert in_dim % num_blocks == 0 and out_dim % num_blocks == 0, \
        f'For block diagonal decomposition, input dimensions ({in_dim}, {out_dim}) must be divisible ' \
        f'by number of blocks ({num_blocks})'


class _PlanCacheMixin:
    """The cached GraphPlan holds ctypes structs with device pointers: it is dropped when a layer is copied or pickled
    (copy.deepcopy(model), torch.save(model)) and rebuilt lazily at the next forward."""

    def __getstate__(self):
        state = dict(self.__dict__)
        if '_plan_cache' in state:
            state['_plan_cache'] = None
        return state

    def __deepcopy__(self, memo):
        import copy
        cls = self.__class__
        new = cls.__new__(cls)
        memo[id(self)] = new
        for k, v in self.__dict__.items():
            new.__dict__[k] = None if k == '_plan_cache' else copy.deepcopy(v, memo)
        return new


class RelationalGraphConvolutionNC(_PlanCacheMixin, Module):
    """Relational graph convolution for node classification (graph fixed at construction).

    Mirrors reference torch_rgcn/layers.py:101-308.  The graph plan (sorted CSR + per-edge weights) is
    built once per device on first use and cached, instead of per forward.
    """

    def __init__(self, triples=None, num_nodes=None, num_relations=None, in_features=None, out_features=None,
                 edge_dropout=None, edge_dropout_self_loop=None, bias=True, decomposition=None,
                 vertical_stacking=False, diag_weight_matrix=False, reset_mode='glorot_uniform'):
        super().__init__()
        assert (triples is not None or num_nodes is not None or num_relations is not None or
                out_features is not None), \
            "The following must be specified: triples, number of nodes, number of relations and output dimension!"
        in_dim = in_features if in_features is not None else num_nodes     # featureless: one-hot input
        weight_decomp, num_bases, num_blocks = _unpack_decomposition(decomposition)

        self.triples = triples
        self.num_nodes = num_nodes
        self.num_relations = num_relations
        self.in_features = in_features
        self.out_features = out_features
        self.weight_decomp = weight_decomp
        self.num_bases = num_bases
        self.num_blocks = num_blocks
        self.vertical_stacking = vertical_stacking
        self.diag_weight_matrix = diag_weight_matrix
        self.edge_dropout = edge_dropout                       # stored, never applied (as in the reference)
        self.edge_dropout_self_loop = edge_dropout_self_loop

        if diag_weight_matrix:
            self.weights = Parameter(torch.empty((num_relations, in_features)), requires_grad=True)
            self.out_features = in_features
            self.weight_decomp = None
            bias = False
        elif weight_decomp is None:
            self.weights = Parameter(torch.FloatTensor(num_relations, in_dim, out_features))
        elif weight_decomp == 'basis':
            assert num_bases > 0, 'Number of bases should be set to higher than zero for basis decomposition!'
            self.bases = Parameter(torch.FloatTensor(num_bases, in_dim, out_features))
            self.comps = Parameter(torch.FloatTensor(num_relations, num_bases))
        elif weight_decomp == 'block':
            _check_blocks(num_blocks, in_dim, out_features)
            self.blocks = Parameter(torch.FloatTensor(num_relations, num_blocks, in_dim // num_blocks,
                                                      out_features // num_blocks))
        else:
            raise NotImplementedError(f'{weight_decomp} decomposition has not been implemented')

        if bias:
            self.bias = Parameter(torch.FloatTensor(out_features))
        else:
            self.register_parameter('bias', None)

        self._plan_cache = None
        self.validate_triples = True
        self.reset_parameters(reset_mode)

    def _decomposed(self):
        if self.weight_decomp == 'block':
            return [self.blocks]
        if self.weight_decomp == 'basis':
            return [self.bases, self.comps]
        return [self.weights]

    def reset_parameters(self, reset_mode='glorot_uniform'):
        """Same draws, in the same order, as reference layers.py:182-220."""
        if reset_mode in ('glorot_uniform', 'schlichtkrull'):
            gain = nn.init.calculate_gain('relu')
            for p in self._decomposed():
                nn.init.xavier_uniform_(p, gain=gain)
            if self.bias is not None:
                nn.init.zeros_(self.bias)
        elif reset_mode == 'uniform':
            stdv = 1.0 / math.sqrt(self.weights.size(1))       # AttributeError for decomposed layers, as upstream
            for p in self._decomposed():
                p.data.uniform_(-stdv, stdv)
            if self.bias is not None:
                self.bias.data.uniform_(-stdv, stdv)
        else:
            raise NotImplementedError(f'{reset_mode} parameter initialisation method has not been implemented')

    # -- graph plan -------------------------------------------------------------------------------------
    def _tile_edges(self, features):
        """Edges per row super-tile for the L2-resident message ring (bf16 features, 16x16 blocks only)."""
        if (features is None or features.dtype != torch.bfloat16 or self.weight_decomp != 'block' or
                self.in_features != self.out_features or self.in_features // self.num_blocks != 16 or
                self.num_blocks % 4 != 0 or self.num_blocks // 4 not in (1, 2, 4, 8)):
            return 0
        # RGCN_TILE_MB: message bytes per row super-tile.  Default: tiling only when the untiled message buffer
        # (nnz rows of bf16 messages) would exceed 8 GB, e.g. the 200 M-edge / 512-wide synthetic config.
        env = os.environ.get('RGCN_TILE_MB')
        row_bytes = self.out_features * 2
        if env is None:
            nnz = self.triples.size(0)
            if nnz * row_bytes <= (8 << 30):
                return 0
            tile_bytes = 256 << 20
        else:
            tile_bytes = int(float(env) * (1 << 20))
        return max(tile_bytes // row_bytes, 4096) if tile_bytes > 0 else 0

    def _fuse_rows(self, features):
        """Rows per block of the fused row-block kernel, 0 = off.  It serves bf16 features with 16x16 weight blocks in
        groups of four (width 64, 128, ... 512: the layer is width / 64 independent 64-column layers over the same graph).

        RGCN_FUSED=0 disables it, RGCN_FUSE_ROWS overrides the block height (a multiple of 16; 640 rows x 64 fp32
        columns = 160 KB of the CTA's shared memory)."""
        if (features is None or features.dtype != torch.bfloat16 or self.weight_decomp != 'block' or
                self.in_features != self.out_features or self.num_blocks % 4 != 0 or
                self.in_features != 16 * self.num_blocks):
            return 0
        if os.environ.get('RGCN_FUSED', _FUSED_DEFAULT) == '0':
            return 0
        return int(os.environ.get('RGCN_FUSE_ROWS', '640'))

    def _plan(self, device, features=None):
        t = self.triples
        tile_edges = self._tile_edges(features)
        fuse_rows = self._fuse_rows(features)
        key = (str(device), t.data_ptr(), t._version, tuple(t.shape), self.vertical_stacking, tile_edges, fuse_rows)
        if self._plan_cache is None or self._plan_cache[0] != key:
            nnz = t.size(0)
            n_general = int((nnz - self.num_nodes) / 2)          # reference layers.py:235
            norm = _lib.NORM_ROW if self.vertical_stacking else _lib.NORM_COL_SWAPPED
            plan = GraphPlan(t.to(device), self.num_nodes, self.num_relations, norm, n_general, self.num_nodes,
                             validate=self.validate_triples, tile_edges=tile_edges,
                             ring_depth=int(os.environ.get('RGCN_RING_DEPTH', '8')), fuse_rows=fuse_rows,
                             fuse_item_tiles=int(os.environ.get('RGCN_FUSE_ITEM_TILES', '4096')),
                             fuse_dirs=_fuse_dirs(self.out_features or 64))
            self._plan_cache = (key, plan)
        return self._plan_cache[1]

    def set_plan(self, plan):
        """Install an externally built plan (e.g. a relation shard, see parallel.py)."""
        t = self.triples
        key = (str(plan.device), t.data_ptr(), t._version, tuple(t.shape), self.vertical_stacking, plan.tile_edges,
               plan.fuse_rows)
        self._plan_cache = (key, plan)

    def forward(self, features=None):
        """One pass of message propagation: (num_nodes, out_features) fp32."""
        assert (features is None) == (self.in_features is None), "in_features not provided!"
        lead = self._decomposed()[0]
        _lib.require_cuda(lead, features)
        if self.in_features is None and self.vertical_stacking:
            # the reference reaches torch.mm with mismatched shapes here (layers.py:286-288)
            raise RuntimeError('featureless message passing requires horizontal stacking (vertical_stacking=False)')
        if self.diag_weight_matrix and self.vertical_stacking:
            raise RuntimeError('diagonal weight matrices require horizontal stacking (vertical_stacking=False)')
        plan = self._plan(lead.device, features)
        in_dim = self.in_features if self.in_features is not None else self.num_nodes
        if self.diag_weight_matrix:
            assert self.weights.size() == (self.num_relations, in_dim)
            return rgcn_propagate(plan, 'diag', in_dim, self.out_features, features, weights=self.weights)
        if self.weight_decomp is None:
            assert self.weights.size() == (self.num_relations, in_dim, self.out_features)
            return rgcn_propagate(plan, 'dense', in_dim, self.out_features, features, weights=self.weights,
                                  bias=self.bias)
        if self.weight_decomp == 'basis':
            return rgcn_propagate(plan, 'basis', in_dim, self.out_features, features, bases=self.bases,
                                  comps=self.comps, bias=self.bias)
        if self.weight_decomp == 'block':
            return rgcn_propagate(plan, 'block', in_dim, self.out_features, features, blocks=self.blocks,
                                  bias=self.bias)
        raise NotImplementedError(f'{self.weight_decomp} decomposition has not been implemented')


class RelationalGraphConvolutionLP(Module):
    """Relational graph convolution for link prediction (graph passed per forward).

    Mirrors reference torch_rgcn/layers.py:311-565, including its edge list [T; inverse(T); T; self-loops]
    (the original triples appear twice, utils.py:124), the bernoulli self-loop dropout and the
    'schlichtkrull-dropout' mask on the dense self-relation of the block decomposition.
    """

    def __init__(self, num_nodes=None, num_relations=None, in_features=None, out_features=None, edge_dropout=None,
                 edge_dropout_self_loop=None, decomposition=None, vertical_stacking=False, w_init='glorot-normal',
                 w_gain=False, b_init=None):
        super().__init__()
        assert (num_nodes is not None or num_relations is not None or out_features is not None), \
            "The following must be specified: number of nodes, number of relations and output dimension!"
        device = 'cuda' if torch.cuda.is_available() else 'cpu'      # reference layers.py:334
        in_dim = in_features if in_features is not None else num_nodes
        weight_decomp, num_bases, num_blocks = _unpack_decomposition(decomposition)

        self.num_nodes = num_nodes
        self.num_relations = num_relations
        self.in_features = in_dim
        self.out_features = out_features
        self.weight_decomp = weight_decomp
        self.num_bases = num_bases
        self.num_blocks = num_blocks
        self.vertical_stacking = vertical_stacking
        self.edge_dropout = edge_dropout
        self.edge_dropout_self_loop = edge_dropout_self_loop
        self.w_init = w_init
        self.w_gain = w_gain
        self.b_init = b_init

        if weight_decomp is None:
            self.weights = Parameter(torch.FloatTensor(num_relations, in_dim, out_features).to(device))
        elif weight_decomp == 'basis':
            assert num_bases > 0, 'Number of bases should be set to higher than zero for basis decomposition!'
            self.bases = Parameter(torch.FloatTensor(num_bases, in_dim, out_features).to(device))
            self.comps = Parameter(torch.FloatTensor(num_relations, num_bases).to(device))
        elif weight_decomp == 'block':
            _check_blocks(num_blocks, in_dim, out_features)
            self.blocks = Parameter(torch.FloatTensor(num_relations - 1, num_blocks, in_dim // num_blocks,
                                                      out_features // num_blocks).to(device))
            self.blocks_self = Parameter(torch.FloatTensor(in_dim, out_features).to(device))
        else:
            raise NotImplementedError(f'{weight_decomp} decomposition has not been implemented')

        if b_init:
            self.bias = Parameter(torch.FloatTensor(out_features))
        else:
            self.register_parameter('bias', None)

        self.validate_triples = True
        self._test_keep = None        # tests inject RNG outcomes here (CPU and CUDA generators differ)
        self._test_self_mask = None
        self.initialise_weights()
        if self.bias is not None:
            self.initialise_biases()

    def initialise_biases(self):
        select_b_init(self.b_init)(self.bias)

    def initialise_weights(self):
        """Same draws as reference layers.py:405-447."""
        gain = nn.init.calculate_gain('relu') if self.w_gain else 1.0
        init = select_w_init(self.w_init)
        if self.weight_decomp == 'block':
            shape = [(self.num_relations - 1) // 2, self.in_features // self.num_blocks]
            schlichtkrull_normal_(self.blocks, shape=shape, gain=gain)
            schlichtkrull_normal_(self.blocks_self, shape=shape, gain=gain)
        elif self.weight_decomp == 'basis':
            init(self.bases, gain=gain)
            init(self.comps, gain=gain)
        else:
            init(self.weights, gain=gain)

    def forward(self, triples, features=None):
        """One pass of message propagation over `triples` (E, 3): (num_nodes, out_features) fp32."""
        assert (features is None) == (self.in_features is None), "in_features not given"
        lead = self.blocks if self.weight_decomp == 'block' else (
            self.bases if self.weight_decomp == 'basis' else self.weights)
        _lib.require_cuda(lead)
        device = lead.device
        triples = triples.to(device)
        features = features.to(device)
        N, Rp = self.num_nodes, self.num_relations
        R = int((Rp - 1) / 2)                                          # reference layers.py:460
        if self.weight_decomp == 'block' and self.vertical_stacking:
            # reference layers.py:527-528 concatenates a 3-D and a 2-D tensor here
            raise RuntimeError('block decomposition in the LP layer requires horizontal stacking')

        schlichtkrull = self.training and self.edge_dropout["self_loop_type"] == 'schlichtkrull-dropout'
        keep_prob = 1 - self.edge_dropout["self_loop"] if (self.training and not schlichtkrull) else 1
        with torch.no_grad():
            # same RNG draw as generate_self_loops (reference utils.py:120-121), even when keep_prob == 1
            keep = torch.bernoulli(torch.empty(size=(N,), dtype=torch.float, device=device).fill_(keep_prob))
            if self._test_keep is not None:
                keep = self._test_keep.to(device)
            nodes = torch.arange(N, device=device)
            if keep_prob != 1 or self._test_keep is not None:
                nodes = nodes[keep.to(torch.bool)]
            t = triples.to(torch.long).contiguous()
            E, n_self = t.size(0), nodes.numel()
            triples_plus = torch.empty(3 * E + n_self, 3, dtype=torch.long, device=device)
            with torch.cuda.device(device):
                _lib.check(_lib.lib.rgcn_lp_triples_plus(_lib.ptr(t), E, R, _lib.ptr(nodes), n_self,
                                                         _lib.ptr(triples_plus), _lib.stream_ptr()))
            norm = _lib.NORM_ROW if self.vertical_stacking else _lib.NORM_COL_SWAPPED
            plan = GraphPlan(triples_plus, N, Rp, norm, E, E + n_self, validate=self.validate_triples)

        if self.weight_decomp is None:
            return rgcn_propagate(plan, 'dense', self.in_features, self.out_features, features,
                                  weights=self.weights, bias=self.bias)
        if self.weight_decomp == 'basis':
            return rgcn_propagate(plan, 'basis', self.in_features, self.out_features, features, bases=self.bases,
                                  comps=self.comps, bias=self.bias)
        self_mask = None
        if schlichtkrull:
            # same draw as F.dropout on the (1, N, O) self-loop messages (reference layers.py:544-546)
            self_mask = nn.functional.dropout(torch.ones(1, N, self.out_features, device=device),
                                              p=self.edge_dropout["self_loop"], training=True)[0]
            if self._test_self_mask is not None:
                self_mask = self._test_self_mask.to(device)
        return rgcn_propagate(plan, 'block', self.in_features, self.out_features, features, blocks=self.blocks,
                              blocks_self=self.blocks_self, bias=self.bias, self_mask=self_mask)


def graph_lp_layer(layer, triples, features):
    """CUDA-graph the link-prediction layer for a fixed problem shape (reference layers.py:450-565 rebuilds its graph
    every forward; here that is ~30 small kernels — augmentation, three radix sorts, decode, normalisation, the
    propagation kernels — whose launch gaps, not their work, dominate a WN18-sized step).

    Returns a callable `f(triples, features) -> out` that replays ONE captured graph for the forward and one for the
    backward (`torch.cuda.make_graphed_callables`): same numerics and autograd behaviour as `layer(triples, features)`,
    for inputs of the SAME shapes and dtypes as the samples (a training loop with a fixed sampled-graph size, or the
    evaluation graph).  The layer must not need host decisions inside the forward: eval mode or no self-loop
    dropout by node removal, and `validate_triples = False` (index validation is a host read-back)."""
    assert isinstance(layer, RelationalGraphConvolutionLP)
    assert not (layer.training and layer.edge_dropout is not None and layer.edge_dropout.get("self_loop", 0) and
                layer.edge_dropout.get("self_loop_type") != 'schlichtkrull-dropout'), \
        'self-loop dropout by node removal changes the edge count per step: it cannot be captured in a CUDA graph'
    layer.validate_triples = False
    return torch.cuda.make_graphed_callables(layer, (triples, features))
