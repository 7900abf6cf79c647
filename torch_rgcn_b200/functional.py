"""Autograd shim over rgcn_forward / rgcn_backward (include/rgcn_b200.h).

`rgcn_propagate` computes  out[s] = bias + sum_e val_e * T_{p_e}(X[o_e])  for one GraphPlan and one
weight form, and its gradients with respect to the features and every parameter.  It replaces the
reference's sparse-adjacency construction + `torch.mm` / `torch.spmm` / `einsum` branches
(torch_rgcn/layers.py:276-306 and :513-556) and the autograd graph behind them.
"""
import ctypes as C

import torch

from . import _lib

_FORMS = {'dense': _lib.W_DENSE, 'basis': _lib.W_BASIS, 'block': _lib.W_BLOCK, 'diag': _lib.W_DIAG}


def _f32c(t):
    if t is None:
        return None
    if t.dtype != torch.float32:
        t = t.float()
    return t.contiguous()


def _params_struct(form, featureless, in_dim, out_dim, weights, bases, comps, blocks, blocks_self, bias, self_mask):
    p = _lib.Params()
    p.form = _FORMS[form]
    p.featureless = 1 if featureless else 0
    p.in_dim, p.out_dim = in_dim, out_dim
    p.num_bases = bases.size(0) if bases is not None else 0
    p.num_blocks = blocks.size(1) if blocks is not None else 0
    p.num_block_rels = blocks.size(0) if blocks is not None else 0
    for name, t in (('weights', weights), ('bases', bases), ('comps', comps), ('blocks', blocks),
                    ('blocks_self', blocks_self), ('bias', bias), ('self_mask', self_mask)):
        setattr(p, name, t.data_ptr() if t is not None else None)
    return p


class _Propagate(torch.autograd.Function):
    @staticmethod
    def forward(ctx, plan, form, in_dim, out_dim, features, weights, bases, comps, blocks, blocks_self, bias,
                self_mask, add_bias=True):
        featureless = features is None
        # gradients are handed back in the dtype each input arrived in (the kernels compute in fp32 / bf16)
        ctx.in_dtypes = [None if t is None else t.dtype
                         for t in (features, weights, bases, comps, blocks, blocks_self, bias)]
        tensors = [_f32c(t) for t in (weights, bases, comps, blocks, blocks_self, bias, self_mask)]
        weights, bases, comps, blocks, blocks_self, bias, self_mask = tensors
        if features is not None:
            if features.dtype not in (torch.float32, torch.bfloat16):
                features = features.float()
            features = features.contiguous()
            assert features.shape == (plan.num_nodes, in_dim), \
                f'features must be ({plan.num_nodes}, {in_dim}), got {tuple(features.shape)}'
        _lib.require_cuda(features, *tensors)
        dev = plan.device
        # add_bias=False: the bias takes part in autograd but another relation shard adds it (parallel.py)
        p = _params_struct(form, featureless, in_dim, out_dim, weights, bases, comps, blocks, blocks_self,
                           bias if add_bias else None, self_mask)
        out = torch.empty(plan.num_nodes, out_dim, dtype=torch.float32, device=dev)
        dt = _lib.BF16 if (features is not None and features.dtype == torch.bfloat16) else _lib.F32
        ws_bytes = _lib.lib.rgcn_forward_workspace_bytes(C.byref(plan.c), C.byref(p), dt)
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev) if ws_bytes else None
        with torch.cuda.device(dev):
            _lib.check(_lib.lib.rgcn_forward(C.byref(plan.c), C.byref(p), _lib.ptr(features), dt, _lib.ptr(out),
                                             _lib.ptr(ws), ws_bytes, _lib.stream_ptr()))
        ctx.plan, ctx.form, ctx.dims = plan, form, (in_dim, out_dim)
        ctx.save_for_backward(features, weights, bases, comps, blocks, blocks_self, bias, self_mask)
        return out

    @staticmethod
    def backward(ctx, grad_out):
        features, weights, bases, comps, blocks, blocks_self, bias, self_mask = ctx.saved_tensors
        plan, form, (in_dim, out_dim) = ctx.plan, ctx.form, ctx.dims
        dev = plan.device
        grad_out = _f32c(grad_out)
        # positions in forward(): plan form in_dim out_dim features weights bases comps blocks blocks_self bias mask
        need = ctx.needs_input_grad
        p = _params_struct(form, features is None, in_dim, out_dim, weights, bases, comps, blocks, blocks_self, bias,
                           self_mask)
        rows = getattr(ctx, 'rows', None)        # a row-sharded caller needs (and its plan covers) only these rows
        if rows is not None:
            p.row_lo, p.row_hi = rows

        def alloc(flag, like, dtype=torch.float32):
            return torch.empty(like.shape, dtype=dtype, device=dev) if (flag and like is not None) else None

        # bf16 features: the engine writes the feature gradient in bf16 directly (no fp32 round trip)
        g_feat = alloc(need[4], features, torch.bfloat16 if (features is not None and features.dtype == torch.bfloat16)
                       else torch.float32)
        g_w = alloc(need[5], weights)
        want_basis = (need[6] or need[7]) and bases is not None
        g_bases = alloc(want_basis, bases)
        g_comps = alloc(want_basis, comps)
        g_blocks = alloc(need[8], blocks)
        g_self = alloc(need[9], blocks_self)
        g_bias = alloc(need[10], bias)
        gr = _lib.Grads()
        for name, t in (('features', g_feat), ('weights', g_w), ('bases', g_bases), ('comps', g_comps),
                        ('blocks', g_blocks), ('blocks_self', g_self), ('bias', g_bias)):
            setattr(gr, name, t.data_ptr() if t is not None else None)
        gr.features_dtype = _lib.BF16 if (g_feat is not None and g_feat.dtype == torch.bfloat16) else _lib.F32
        dt = _lib.BF16 if (features is not None and features.dtype == torch.bfloat16) else _lib.F32
        ws_bytes = _lib.lib.rgcn_backward_workspace_bytes(C.byref(plan.c), C.byref(p), dt)
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev) if ws_bytes else None
        with torch.cuda.device(dev):
            _lib.check(_lib.lib.rgcn_backward(C.byref(plan.c), C.byref(p), _lib.ptr(features), dt,
                                              _lib.ptr(grad_out), C.byref(gr), _lib.ptr(ws), ws_bytes,
                                              _lib.stream_ptr()))
        grads = [g_feat, g_w, g_bases if need[6] else None, g_comps if need[7] else None, g_blocks, g_self, g_bias]
        grads = [g if (g is None or dt is None or g.dtype == dt) else g.to(dt) for g, dt in zip(grads, ctx.in_dtypes)]
        return (None, None, None, None, *grads, None, None)


def rgcn_propagate(plan, form, in_dim, out_dim, features=None, weights=None, bases=None, comps=None, blocks=None,
                   blocks_self=None, bias=None, self_mask=None):
    """out (N, out_dim) fp32.  `features=None` means the featureless (one-hot input) layer."""
    return _Propagate.apply(plan, form, in_dim, out_dim, features, weights, bases, comps, blocks, blocks_self, bias,
                            self_mask, True)
