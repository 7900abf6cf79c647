"""Device implementations of the graph helpers of the reference's torch_rgcn/utils.py.

Same names, argument meaning and results as the reference functions cited on each helper, but the
arithmetic runs in the CUDA library (include/rgcn_b200.h).  Inputs may live on the CPU (the reference's
models build `triples_plus` before `.cuda()`): they are moved to the current CUDA device, computed
there and returned on the device named by `device` — there is no CPU implementation here.

The parameter-initialisation helpers at the bottom are not on the hot path; they exist so the drop-in
layers draw the same random numbers, in the same order, as the reference constructors.
"""
from math import floor, sqrt
import random

import torch

from . import _lib


def _dev():
    _lib.require_cuda()
    return torch.device('cuda', torch.cuda.current_device())


def _to_cuda_long(t):
    t = torch.as_tensor(t)
    dev = t.device if t.is_cuda else _dev()
    return t.to(device=dev, dtype=torch.long).contiguous()


def add_inverse_and_self(triples, num_nodes, num_rels, device='cpu'):
    """[triples; inverse (o, p+R, s); self-loops (v, 2R, v)]  — reference utils.py:127-141."""
    t = _to_cuda_long(triples)
    out = torch.empty(2 * t.size(0) + num_nodes, 3, dtype=torch.long, device=t.device)
    with torch.cuda.device(t.device):
        _lib.check(_lib.lib.rgcn_add_inverse_and_self(_lib.ptr(t), t.size(0), num_nodes, num_rels, _lib.ptr(out),
                                                      _lib.stream_ptr()))
    return out.to(device)


def generate_inverses(triples, num_rels):
    """(o, p+R, s) for every triple — reference utils.py:100-107."""
    src = torch.as_tensor(triples)
    t = _to_cuda_long(src)
    out = torch.empty_like(t)
    with torch.cuda.device(t.device):
        _lib.check(_lib.lib.rgcn_generate_inverses(_lib.ptr(t), t.size(0), num_rels, _lib.ptr(out), _lib.stream_ptr()))
    return out.to(src.device)


def generate_self_loops(triples, num_nodes, num_rels, self_loop_keep_prob, device='cpu'):
    """cat([triples, kept self-loops]) — reference utils.py:110-124 (yes, it returns the triples too)."""
    t = _to_cuda_long(triples)
    mask = torch.bernoulli(torch.empty(size=(num_nodes,), dtype=torch.float, device=t.device)
                           .fill_(self_loop_keep_prob)).to(torch.bool)
    nodes = torch.arange(num_nodes, device=t.device)[mask]
    loops = torch.stack([nodes, torch.full_like(nodes, 2 * num_rels), nodes], dim=1)
    return torch.cat([t, loops], dim=0).to(device)


def stack_matrices(triples, num_nodes, num_rels, vertical_stacking=True, device='cpu'):
    """COO coordinates + size of the stacked adjacency — reference utils.py:143-166."""
    assert triples.dtype == torch.long
    r, n = num_rels, num_nodes
    size = (r * n, n) if vertical_stacking else (n, r * n)
    t = _to_cuda_long(triples)
    indices = torch.empty(t.size(0), 2, dtype=torch.long, device=t.device)
    bounds = torch.empty(2, dtype=torch.long, device=t.device)
    with torch.cuda.device(t.device):
        _lib.check(_lib.lib.rgcn_stack_matrices(_lib.ptr(t), t.size(0), n, r, 1 if vertical_stacking else 0,
                                                _lib.ptr(indices), _lib.ptr(bounds), _lib.stream_ptr()))
    if t.size(0):
        hi = bounds.tolist()
        assert hi[0] < size[0], f'{hi[0]}, {size}, {r}'
        assert hi[1] < size[1], f'{hi[1]}, {size}, {r}'
    return indices.to(device), size


def sum_sparse(indices, values, size, row_normalisation=True, device='cpu'):
    """Row/column sums of a sparse matrix redistributed to its entries — reference utils.py:71-97."""
    assert len(indices.size()) == len(values.size()) + 1
    idx = _to_cuda_long(indices)
    vals = torch.as_tensor(values).to(device=idx.device, dtype=torch.float32).contiguous()
    k = idx.size(0)
    length = size[0] if row_normalisation else size[1]
    table = torch.empty(max(int(length), 1), dtype=torch.float32, device=idx.device)
    out = torch.empty(k, dtype=torch.float32, device=idx.device)
    with torch.cuda.device(idx.device):
        _lib.check(_lib.lib.rgcn_sum_sparse(_lib.ptr(idx), _lib.ptr(vals), k, int(size[0]), int(size[1]),
                                            1 if row_normalisation else 0, _lib.ptr(table), _lib.ptr(out),
                                            _lib.stream_ptr()))
    return out.to(device).view(k)


class _BlockDiag(torch.autograd.Function):
    """rgcn_block_diag with the gradient the reference's differentiable block_diag has: the diagonal blocks of grad."""

    @staticmethod
    def forward(ctx, mm):
        n, nb, bi, bo = mm.shape
        ctx.dims = (nb, bi, bo)
        out = torch.empty(n, nb * bi, nb * bo, dtype=torch.float32, device=mm.device)
        with torch.cuda.device(mm.device):
            _lib.check(_lib.lib.rgcn_block_diag(_lib.ptr(mm), n, nb, bi, bo, _lib.ptr(out), _lib.stream_ptr()))
        return out

    @staticmethod
    def backward(ctx, grad):
        nb, bi, bo = ctx.dims
        g = grad.reshape(-1, nb, bi, nb, bo)
        idx = torch.arange(nb, device=grad.device)
        return g[:, idx, :, idx, :].permute(1, 0, 2, 3).contiguous()     # (n, nb, bi, bo)


def block_diag(m):
    """(..., nb, bi, bo) -> (..., nb*bi, nb*bo) block-diagonal — reference utils.py:168-196 (differentiable, like
    the reference's: it sits on the autograd path of layers.py:244 and :521)."""
    if type(m) is list:
        m = torch.cat([m1.unsqueeze(-3) for m1 in m], -3)
    lead = m.shape[:-3]
    nb, bi, bo = m.shape[-3:]
    src = m
    mm = m.to(device=m.device if m.is_cuda else _dev(), dtype=torch.float32).reshape(-1, nb, bi, bo).contiguous()
    out = _BlockDiag.apply(mm)
    return out.reshape(lead + (nb * bi, nb * bo)).to(src.device)


def attach_dim(v, n_dim_to_prepend=0, n_dim_to_append=0):
    """reference utils.py:198-199"""
    return v.reshape(torch.Size([1] * n_dim_to_prepend) + v.shape + torch.Size([1] * n_dim_to_append))


def split_spo(triples):
    """reference utils.py:201-206"""
    if triples.dim() == 2:
        return triples[:, 0], triples[:, 1], triples[:, 2]
    return triples[:, :, 0], triples[:, :, 1], triples[:, :, 2]


def drop_edges(triples, num_nodes, general_edo, self_loop_edo):
    """Edge dropout by row selection — reference utils.py:57-69 (host-side index sampling, unused by the layers)."""
    nt = triples.size(0) - num_nodes
    keep = random.sample(range(nt), k=int(floor((1.0 - general_edo) * nt)))
    keep += random.sample(range(nt, nt + num_nodes), k=int(floor((1.0 - self_loop_edo) * num_nodes)))
    return triples[keep, :]


# ---- initialisers (reference utils.py:6-55): kept call-compatible so RNG streams match -------------------
def schlichtkrull_std(shape, gain):
    fan_in, fan_out = shape[0], shape[1]
    return gain * 3.0 / sqrt(float(fan_in + fan_out))


def schlichtkrull_normal_(tensor, shape, gain=1.):
    with torch.no_grad():
        return tensor.normal_(0.0, schlichtkrull_std(shape, gain))


def schlichtkrull_uniform_(tensor, gain=1.):
    # the reference passes the tensor itself as `shape` (utils.py:21); same arithmetic, same failure modes
    std = schlichtkrull_std(tensor, gain)
    with torch.no_grad():
        return tensor.uniform_(-std, std)


def select_b_init(init):
    init = init.lower()
    table = {'zeros': torch.nn.init.zeros_, 'zero': torch.nn.init.zeros_, 'ones': torch.nn.init.ones_,
             'one': torch.nn.init.ones_, 'uniform': torch.nn.init.uniform_, 'normal': torch.nn.init.normal_}
    if init not in table:
        raise NotImplementedError(f'{init} initialisation has not been implemented!')
    return table[init]


def select_w_init(init):
    init = init.lower()
    table = {'glorot-uniform': torch.nn.init.xavier_uniform_, 'xavier-uniform': torch.nn.init.xavier_uniform_,
             'glorot-normal': torch.nn.init.xavier_normal_, 'xavier-normal': torch.nn.init.xavier_normal_,
             'schlichtkrull-uniform': schlichtkrull_uniform_, 'schlichtkrull-normal': schlichtkrull_normal_,
             'normal': torch.nn.init.normal_, 'standard-normal': torch.nn.init.normal_,
             'uniform': torch.nn.init.uniform_}
    if init not in table:
        raise NotImplementedError(f'{init} initialisation has not been implemented!')
    return table[init]
