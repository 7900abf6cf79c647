"""Parity of the headline kernels AT HEADLINE SIZE ("What's weak" 1): the full AM-shaped bf16 64 -> 64 layer
(13.6 M edges, 267 relations; uniform and power-law graphs) and a 512-wide / 32-block layer at 10 M edges, forward and
backward, against the fp64 oracle restricted to sampled rows:

  * 10,000 sampled destination rows of `out`          (every edge into those rows, oracle closed form)
  * 10,000 sampled source rows of the feature gradient (every edge out of those rows)
  * the FULL `blocks.grad`                             (all edges; per-relation fp64 GEMMs; 24 sampled relations of
                                                        the 512-wide layer)

Per-edge weights come from the oracle's literal normalisation rule on the full graph (oracle.rgcn_oracle.edge_values), so
int32 slot maps, chunk pointers, long-row lists, split row blocks and the 267-relation weight tables are all exercised
at the size the bench runs.  Tolerance: the stated bf16 bar, 1e-2 of the tensor's scale (features, upstream gradient and
MMA operands are bf16; products and sums fp32)."""
import numpy as np
import pytest
import torch

from oracle import rgcn_oracle as orc

pytestmark = pytest.mark.gpu


def _oracle_rows(tp, val, blocks, X, rows, gather_col, scatter_col, transpose):
    """fp64 sum_e val_e * X[e.gather] @ blockdiag(blocks[p_e]) (or its transpose) over the edges whose `scatter_col`
    entry is in `rows`; returns (len(rows), width)."""
    Rp, nb, bi, bo = blocks.shape
    lut = np.full(int(tp[:, scatter_col].max()) + 1, -1, np.int64)
    lut[rows] = np.arange(len(rows))
    sel = np.flatnonzero(lut[tp[:, scatter_col]] >= 0)
    out = np.zeros((len(rows), nb * (bi if transpose else bo)))
    W = blocks.astype(np.float64)
    for c0 in range(0, len(sel), 50000):
        e = sel[c0:c0 + 50000]
        x = X[tp[e, gather_col]].astype(np.float64).reshape(len(e), nb, -1)
        w = W[tp[e, 1]]
        msg = np.einsum('ebo,ebio->ebi', x, w) if transpose else np.einsum('ebi,ebio->ebo', x, w)
        np.add.at(out, lut[tp[e, scatter_col]], val[e, None].astype(np.float64) * msg.reshape(len(e), -1))
    return out


def _oracle_block_grad(tp, val, X, G, Rp, nb, bi, bo, rels=None):
    """fp64 gblocks[p, b] = sum_{e in p} val_e X[o_e, b]^T G[s_e, b]: one GEMM per (relation, block); rows of the
    relations not in `rels` (default: all) stay zero."""
    order = np.argsort(tp[:, 1], kind='stable')
    bounds = np.searchsorted(tp[order, 1], np.arange(Rp + 1))
    g = np.zeros((Rp, nb, bi, bo))
    for p in (range(Rp) if rels is None else rels):
        for c0 in range(bounds[p], bounds[p + 1], 400000):
            e = order[c0:min(c0 + 400000, bounds[p + 1])]
            xg = (val[e, None].astype(np.float64) * X[tp[e, 2]].astype(np.float64)).reshape(len(e), nb, bi)
            gg = G[tp[e, 0]].astype(np.float64).reshape(len(e), nb, bo)
            g[p] += np.einsum('ebi,ebo->bio', xg, gg, optimize=True)
    return g


def _close(got, ref, name, tol=1e-2):
    scale = np.abs(ref).max()
    err = np.abs(got - ref)
    bad = int((err > tol * scale + tol * np.abs(ref)).sum())
    assert bad == 0, f'{name}: max err {err.max():.4g} (scale {scale:.4g}), {bad} bad of {err.size}'


def _check_layer(dev, tp_dev, N, Rp, width, nb, vertical, seed, monkeypatch, env=None, grad_rels=None):
    from torch_rgcn_b200.layers import RelationalGraphConvolutionNC
    for k, v in (env or {}).items():
        monkeypatch.setenv(k, v)
    torch.manual_seed(seed)
    layer = RelationalGraphConvolutionNC(triples=tp_dev, num_nodes=N, num_relations=Rp, in_features=width,
                                         out_features=width, decomposition={'type': 'block', 'num_blocks': nb},
                                         vertical_stacking=vertical).to(dev)
    with torch.no_grad():
        layer.bias.copy_(torch.randn(width, device=dev))
    gen = torch.Generator(device=dev).manual_seed(seed + 1)
    x = torch.randn(N, width, device=dev, generator=gen).to(torch.bfloat16).requires_grad_(True)
    G = torch.randn(N, width, device=dev, generator=gen)
    out = layer(x)
    out.backward(G)
    torch.cuda.synchronize()
    plan = layer._plan_cache[1]
    tp = tp_dev.cpu().numpy()
    nnz = tp.shape[0]
    n_gen = int((nnz - N) / 2)
    val = orc.edge_values(tp, N, Rp, vertical, n_gen, N)
    np.testing.assert_array_equal(plan.val[:nnz].cpu().numpy(), val)            # the literal rule, bit-exact, at size
    X = x.detach().float().cpu().numpy()
    Gn = G.cpu().numpy()
    blocks = layer.blocks.detach().cpu().numpy()
    rng = np.random.RandomState(seed)
    rows = np.sort(rng.choice(N, 10000, replace=False))
    # include the heaviest rows (hubs) of both sides
    heavy_d = np.argsort(np.bincount(tp[:, 0], minlength=N))[-50:]
    heavy_s = np.argsort(np.bincount(tp[:, 2], minlength=N))[-50:]
    rows_d, rows_s = np.union1d(rows, heavy_d), np.union1d(rows, heavy_s)
    ref_out = _oracle_rows(tp, val, blocks, X, rows_d, 2, 0, False) + layer.bias.detach().cpu().numpy().astype(np.float64)
    _close(out.detach()[torch.as_tensor(rows_d, device=dev)].cpu().numpy(), ref_out, 'out (sampled destination rows)')
    ref_gx = _oracle_rows(tp, val, blocks, Gn, rows_s, 0, 2, True)
    _close(x.grad[torch.as_tensor(rows_s, device=dev)].float().cpu().numpy(), ref_gx, 'feature gradient (sampled source rows)')
    rels = None if grad_rels is None else np.sort(rng.choice(Rp, grad_rels, replace=False))
    ref_gw = _oracle_block_grad(tp, val, X, Gn, Rp, nb, width // nb, width // nb, rels)
    got_gw = layer.blocks.grad.cpu().numpy()
    if rels is not None:
        got_gw, ref_gw = got_gw[rels], ref_gw[rels]
    _close(got_gw, ref_gw, 'blocks.grad (all edges of %s relations)' % ('all' if rels is None else len(rels)))
    np.testing.assert_allclose(layer.bias.grad.cpu().numpy(), Gn.astype(np.float64).sum(0), rtol=1e-4, atol=1e-2)
    return plan


@pytest.mark.parametrize('skew,fused', [(False, '1'), (True, '1'), (False, '2'), (False, '0')])
def test_am64_full_size(cuda_device, monkeypatch, skew, fused):
    """BASELINE configs[2] at full size: fused row-block forward (default; uniform and power-law graph), fused forward +
    feature gradient, two-phase kernels."""
    from torch_rgcn_b200.synthetic import SHAPES, random_triples
    from torch_rgcn_b200.utils import add_inverse_and_self
    N, R, E = SHAPES['am']
    t = random_triples(N, R, E, seed=0, device=cuda_device, rel_dist='zipf' if skew else 'uniform', node_skew=skew)
    tp = add_inverse_and_self(t, N, R, device=cuda_device)
    plan = _check_layer(cuda_device, tp, N, 2 * R + 1, 64, 4, False, 3, monkeypatch, {'RGCN_FUSED': fused})
    assert (plan.fuse_rows > 0) == (fused != '0')
    if fused != '0':
        assert plan.fused_ok[0]
        if skew:
            assert plan.c.fuse_split[0] >= 0


def test_syn_width_ten_million_edges(cuda_device, monkeypatch):
    """The 512-wide / 32-block shape of BASELINE configs[4] at 10 M arbitrary triples over 256 relations (row
    normalisation): the tiled span kernels and the relation-batched weight-gradient pass."""
    from torch_rgcn_b200.synthetic import random_triples
    N, Rp, E = 250000, 256, 10_000_000
    tp = random_triples(N, Rp, E, seed=5, device=cuda_device)
    _check_layer(cuda_device, tp, N, Rp, 512, 32, True, 7, monkeypatch, grad_rels=24)
