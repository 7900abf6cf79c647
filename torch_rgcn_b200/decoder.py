"""DistMult decoder and negative sampling of the link-prediction models on the CUDA library.

Drop-ins for reference torch_rgcn/layers.py:9-98 (`DistMult`: same constructor, parameter names / shapes /
initialisation order, `forward(triples, nodes)` and `s_penalty(triples, nodes)`) and utils/misc.py:174-189
(`negative_sampling`: same signature and the same random draws in the same order).  The gathers, products,
reductions and gradient scatters run in torch_rgcn_b200/csrc/distmult.cu behind the C ABI; torch owns the memory,
the RNG and the autograd bookkeeping.  No CPU fallback.
"""
import ctypes as C

import torch
from torch import nn
from torch.nn import Module, Parameter

from . import _lib
from .utils import select_b_init, select_w_init


def _flat_triples(triples):
    assert triples.dtype == torch.long, 'triples must be torch.long'
    assert triples.size(-1) == 3 and triples.dim() in (2, 3), 'triples must be (B, 3) or (B, K, 3)'
    return triples.reshape(-1, 3).contiguous(), triples.shape[:-1]


def _f32c(t):
    return None if t is None else t.to(torch.float32).contiguous()


def _check_status(status, what, num_nodes, num_rels):
    bad = int(status.item())
    if bad:
        raise IndexError(f'{bad} {what} triples index a node >= {num_nodes} or a relation >= {num_rels}')


class _Score(torch.autograd.Function):
    @staticmethod
    def forward(ctx, triples, nodes, relations, sbias, pbias, obias, validate):
        t, lead = _flat_triples(triples)
        ctx.in_dtypes = [None if x is None else x.dtype for x in (nodes, relations, sbias, pbias, obias)]
        nodes, relations, sbias, pbias, obias = (_f32c(x) for x in (nodes, relations, sbias, pbias, obias))
        _lib.require_cuda(t, nodes, relations, sbias, pbias, obias)
        assert nodes.dim() == 2 and relations.dim() == 2 and nodes.size(1) == relations.size(1), \
            'nodes (N, d) and relations (R, d) must share the embedding size'
        dev = nodes.device
        B = t.size(0)
        scores = torch.empty(B, dtype=torch.float32, device=dev)
        status = torch.zeros(1, dtype=torch.int32, device=dev) if validate else None
        with torch.cuda.device(dev):
            _lib.check(_lib.lib.rgcn_distmult_forward(_lib.ptr(t), B, _lib.ptr(nodes), nodes.size(0), _lib.ptr(relations),
                                                      relations.size(0), nodes.size(1), _lib.ptr(sbias), _lib.ptr(pbias),
                                                      _lib.ptr(obias), _lib.ptr(scores), _lib.ptr(status),
                                                      _lib.stream_ptr()))
        if validate:
            _check_status(status, 'scored', nodes.size(0), relations.size(0))
        ctx.save_for_backward(t, nodes, relations, sbias)
        ctx.lead = lead
        return scores.reshape(lead)

    @staticmethod
    def backward(ctx, grad):
        t, nodes, relations, sbias = ctx.saved_tensors
        dev = nodes.device
        need = ctx.needs_input_grad               # triples nodes relations sbias pbias obias validate
        grad = _f32c(grad).reshape(-1)
        N, R, d = nodes.size(0), relations.size(0), nodes.size(1)
        g_nodes = torch.empty_like(nodes) if need[1] else None
        g_rel = torch.empty_like(relations) if need[2] else None
        want_bias = sbias is not None and (need[3] or need[4] or need[5])
        g_sb = torch.empty(N, dtype=torch.float32, device=dev) if want_bias else None
        g_ob = torch.empty(N, dtype=torch.float32, device=dev) if want_bias else None
        g_pb = torch.empty(R, dtype=torch.float32, device=dev) if want_bias else None
        with torch.cuda.device(dev):
            _lib.check(_lib.lib.rgcn_distmult_backward(_lib.ptr(t), t.size(0), _lib.ptr(nodes), N, _lib.ptr(relations), R, d,
                                                       _lib.ptr(grad), _lib.ptr(g_nodes), _lib.ptr(g_rel), _lib.ptr(g_sb),
                                                       _lib.ptr(g_pb), _lib.ptr(g_ob), _lib.stream_ptr()))
        grads = [g_nodes, g_rel, g_sb if need[3] else None, g_pb if need[4] else None, g_ob if need[5] else None]
        grads = [g if (g is None or dt is None or g.dtype == dt) else g.to(dt) for g, dt in zip(grads, ctx.in_dtypes)]
        return (None, *grads, None)


class _Penalty(torch.autograd.Function):
    @staticmethod
    def forward(ctx, triples, nodes, relations, validate):
        t, _ = _flat_triples(triples)
        ctx.in_dtypes = [nodes.dtype, relations.dtype]
        nodes, relations = _f32c(nodes), _f32c(relations)
        _lib.require_cuda(t, nodes, relations)
        dev = nodes.device
        out = torch.empty(1, dtype=torch.float32, device=dev)
        status = torch.zeros(1, dtype=torch.int32, device=dev) if validate else None
        ws_bytes = _lib.lib.rgcn_distmult_penalty_workspace_bytes(nodes.size(0), relations.size(0))
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
        with torch.cuda.device(dev):
            _lib.check(_lib.lib.rgcn_distmult_penalty(_lib.ptr(t), t.size(0), _lib.ptr(nodes), nodes.size(0),
                                                      _lib.ptr(relations), relations.size(0), nodes.size(1), _lib.ptr(out),
                                                      _lib.ptr(status), _lib.ptr(ws), ws_bytes, _lib.stream_ptr()))
        if validate:
            _check_status(status, 'penalised', nodes.size(0), relations.size(0))
        ctx.save_for_backward(ws, nodes, relations)      # ws holds the occurrence counts of the batch
        ctx.num_triples = t.size(0)
        return out.reshape(())

    @staticmethod
    def backward(ctx, grad):
        ws, nodes, relations = ctx.saved_tensors
        need = ctx.needs_input_grad
        grad = _f32c(grad).reshape(1)
        g_nodes = torch.empty_like(nodes) if need[1] else None
        g_rel = torch.empty_like(relations) if need[2] else None
        with torch.cuda.device(nodes.device):
            _lib.check(_lib.lib.rgcn_distmult_penalty_backward(_lib.ptr(ws), ctx.num_triples, _lib.ptr(nodes), nodes.size(0),
                                                               _lib.ptr(relations), relations.size(0), nodes.size(1),
                                                               _lib.ptr(grad), _lib.ptr(g_nodes), _lib.ptr(g_rel),
                                                               _lib.stream_ptr()))
        grads = [g if (g is None or g.dtype == dt) else g.to(dt) for g, dt in zip((g_nodes, g_rel), ctx.in_dtypes)]
        return (None, *grads, None)


def distmult_score(triples, nodes, relations, sbias=None, pbias=None, obias=None, validate=True):
    """scores with the leading shape of `triples` — reference layers.py:86-98."""
    return _Score.apply(triples, nodes, relations, sbias, pbias, obias, validate)


def distmult_penalty(triples, nodes, relations, validate=True):
    """Schlichtkrull L2 penalty of the decoder (0-d tensor) — reference layers.py:77-84."""
    return _Penalty.apply(triples, nodes, relations, validate)


class DistMult(Module):
    """DistMult scoring function — mirrors reference torch_rgcn/layers.py:9-98."""

    def __init__(self, indim, outdim, num_nodes, num_rel, w_init='standard-normal', w_gain=False, b_init=None):
        super().__init__()
        self.w_init = w_init
        self.w_gain = w_gain
        self.b_init = b_init
        self.relations = nn.Parameter(torch.FloatTensor(indim, outdim))
        if b_init:
            self.sbias = Parameter(torch.FloatTensor(num_nodes))
            self.obias = Parameter(torch.FloatTensor(num_nodes))
            self.pbias = Parameter(torch.FloatTensor(num_rel))
        else:
            self.register_parameter('sbias', None)
            self.register_parameter('obias', None)
            self.register_parameter('pbias', None)
        self.validate_triples = True          # one host sync per call, like the reference's IndexError on the CPU
        self.initialise_parameters()

    def initialise_parameters(self):
        """Same draws, in the same order, as reference layers.py:38-75."""
        init = select_w_init(self.w_init)
        if self.w_gain:
            init(self.relations, gain=nn.init.calculate_gain('relu'))
        else:
            init(self.relations)
        if self.b_init:
            init = select_b_init(self.b_init)
            init(self.sbias)
            init(self.pbias)
            init(self.obias)

    def s_penalty(self, triples, nodes):
        return distmult_penalty(triples, nodes, self.relations, self.validate_triples)

    def forward(self, triples, nodes):
        return distmult_score(triples, nodes, self.relations, self.sbias, self.pbias, self.obias, self.validate_triples)


def negative_sampling(batch, num_nodes, head_corrupt_prob, device=None):
    """Corrupt the head or the tail of every triple of `batch` (bs, ns, 3) in place; returns (bs * ns, 3).

    Same signature and the same two random draws in the same order as reference utils/misc.py:174-189
    (`randint` for the new entities, then `bernoulli` for head-vs-tail); the masked assignment
    `batch[mask] = corruptions` runs in the CUDA library."""
    if device is None or (torch.device(device).type == 'cpu' and batch.is_cuda):
        device = batch.device                   # the reference's default 'cpu' cannot feed the CUDA kernel
    bs, ns, _ = batch.size()
    corruptions = torch.randint(size=(bs * ns,), low=0, high=num_nodes, dtype=torch.long, device=device)
    mask = torch.bernoulli(torch.empty(size=(bs, ns, 1), dtype=torch.float, device=device).fill_(head_corrupt_prob)).to(torch.bool)
    _lib.require_cuda(batch, corruptions, mask)
    assert batch.dtype == torch.long and batch.is_contiguous(), 'batch must be a contiguous torch.long tensor'
    head = mask.reshape(-1).to(torch.uint8)
    with torch.cuda.device(batch.device):
        _lib.check(_lib.lib.rgcn_corrupt_triples(_lib.ptr(batch), _lib.ptr(head), _lib.ptr(corruptions), bs * ns,
                                                 _lib.stream_ptr()))
    return batch.view(bs * ns, -1)
