"""GPU tests of the fused row-block path (propagate_fused.cuh + the rgcn_fused lists of the graph plan).

1. the lists themselves against a numpy construction of the same definition (exact, integer work);
2. the fused kernel (forward and feature gradient) against the fp64 oracle at the bf16 tolerance, over block
   heights, item sizes that force split (atomically flushed) blocks, hub graphs and multi-edge segments;
3. bit-reproducibility of the forward (no atomics on unsplit blocks) and agreement with the two-phase kernels.
"""
import numpy as np
import pytest
import torch

from oracle import rgcn_oracle as orc
from fused_ref import expected_lists as _expected_lists, records as _records

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize('FR,item_tiles,skew', [(64, 4096, False), (512, 4096, False), (32, 3, True)])
def test_fused_lists_match_definition(cuda_device, FR, item_tiles, skew):
    from torch_rgcn_b200 import _lib
    from torch_rgcn_b200.graph import GraphPlan
    from torch_rgcn_b200.synthetic import random_triples
    N, R, E = 700, 5, 6000
    t = random_triples(N, R, E, seed=11, rel_dist='zipf' if skew else 'uniform', node_skew=skew)
    tp = orc.add_inverse_and_self(t.numpy(), N, R)
    Rp = 2 * R + 1
    plan = GraphPlan(torch.as_tensor(tp).to(cuda_device), N, Rp, _lib.NORM_ROW, fuse_rows=FR, fuse_item_tiles=item_tiles,
                     fuse_dirs=3)
    val = plan.val[:plan.nnz].cpu().numpy()
    for d in (0, 1):
        exp = _expected_lists(tp, N, Rp, val, FR, item_tiles, backward=bool(d))
        arrs = {k: v.cpu().numpy() for k, v in plan._fused[d].items()}
        n_items, tiles, overflow, split, flagged = arrs['meta'].tolist()[:5]
        assert overflow == 0 and plan.fused_ok[d]
        assert tiles == exp['total'], (d, tiles, exp['total'])
        n = tiles * 16
        np.testing.assert_array_equal(arrs['col'][:n], exp['col'], err_msg=f'col d={d}')
        np.testing.assert_array_equal(arrs['rec'][:tiles], _records(exp), err_msg=f'rec d={d}')
        assert flagged == int(exp['serial'].sum())
        np.testing.assert_array_equal(arrs['blk_tile'], exp['blk_tile'], err_msg=f'blk_tile d={d}')
        assert n_items == len(exp['items']), (d, n_items, len(exp['items']))
        np.testing.assert_array_equal(arrs['items'][:n_items], exp['items'], err_msg=f'items d={d}')
        assert split == exp['split']
        assert np.all(arrs['col'][n:] == -1) and np.all(arrs['rec'][tiles:] == 0)       # untouched padding
        assert plan.c.fuse_items[d] == n_items and plan.c.fuse_split[d] == split and plan.c.fuse_tiles[d] == tiles


def _params_np(layer):
    return {n: p.detach().float().cpu().numpy() for n, p in layer.named_parameters()}


def _graph(kind, N, R, E):
    from torch_rgcn_b200.synthetic import random_triples
    if kind == 'uniform':
        return random_triples(N, R, E, seed=21)
    if kind == 'hub':
        return random_triples(N, R, E, seed=22, rel_dist='zipf', node_skew=True)
    # 'multi': few distinct endpoints per relation -> (s, p) segments with many edges and repeated triples
    t = random_triples(N, R, E, seed=23)
    t[:, 0] = t[:, 0] % 97
    return t


def _run(cuda_device, monkeypatch, fused, N, R, t, vertical, grads, fuse_rows='512', item_tiles='4096', seed=8, tma='gather4'):
    from torch_rgcn_b200.layers import RelationalGraphConvolutionNC
    monkeypatch.setenv('RGCN_FUSED', '2' if fused else '0')
    monkeypatch.setenv('RGCN_FUSE_ROWS', fuse_rows)
    monkeypatch.setenv('RGCN_FUSE_ITEM_TILES', item_tiles)
    monkeypatch.setenv('RGCN_TILE_MB', '0')
    monkeypatch.setenv('RGCN_FUSED_TMA', tma)
    tp = torch.as_tensor(orc.add_inverse_and_self(t.numpy(), N, R))
    torch.manual_seed(seed)
    layer = RelationalGraphConvolutionNC(triples=tp, num_nodes=N, num_relations=2 * R + 1, in_features=64,
                                         out_features=64, decomposition={'type': 'block', 'num_blocks': 4},
                                         vertical_stacking=vertical).to(cuda_device)
    with torch.no_grad():
        layer.bias.copy_(torch.randn(64, device=cuda_device))
    g = torch.Generator(device='cpu').manual_seed(seed + 1)
    feats = torch.randn(N, 64, generator=g).to(cuda_device).to(torch.bfloat16)
    if grads != 'weights_only':
        feats.requires_grad_(True)
    if grads == 'features_only':
        layer.blocks.requires_grad_(False)
        layer.bias.requires_grad_(False)
    out = layer(feats)
    G = torch.randn(N, 64, generator=g).to(cuda_device)
    out.backward(G)
    plan = layer._plan_cache[1]
    assert (plan.fuse_rows > 0) == fused
    if fused:
        assert plan.fused_ok == [True, True]
    return layer, tp, feats, out, G, plan


def _close(got, ref, name, tol=1e-2):
    scale = np.abs(ref).max()
    err = np.abs(got - ref)
    bad = np.argwhere(err > tol * scale + tol * np.abs(ref))
    msg = f'{name}: max err {err.max():.4g} (scale {scale:.4g}), {len(bad)} bad of {err.size}; first bad {bad[:8].tolist()}'
    assert len(bad) == 0, msg


@pytest.mark.parametrize('kind,N,R,E', [('uniform', 3000, 9, 40000), ('hub', 3000, 9, 40000), ('multi', 1500, 4, 30000),
                                        ('uniform', 50, 2, 40)])
@pytest.mark.parametrize('fuse_rows,item_tiles', [('640', '4096'), ('64', '4096'), ('128', '2')])
@pytest.mark.parametrize('grads', ['all', 'features_only'])
def test_fused_matches_oracle(cuda_device, monkeypatch, kind, N, R, E, fuse_rows, item_tiles, grads):
    """bf16 tolerance of the tensor-core paths: 1e-2 of the tensor's scale (features and MMA operands are bf16,
    products and sums fp32)."""
    t = _graph(kind, N, R, E)
    vertical = kind != 'hub'                       # hub graph also exercises the horizontal (permuted) weights
    layer, tp, feats, out, G, plan = _run(cuda_device, monkeypatch, True, N, R, t, vertical, grads, fuse_rows, item_tiles)
    if item_tiles == '2' and E > 1000:
        assert plan.c.fuse_split[0] > 0 and plan.c.fuse_split[1] > 0, 'expected split row blocks'
    ref_out, ref_g = orc.nc_layer(tp.numpy(), N, 2 * R + 1, _params_np(layer), feats.detach().float().cpu().numpy(),
                                  vertical, G.cpu().numpy())
    _close(out.detach().cpu().numpy(), ref_out, 'out')
    assert feats.grad.dtype == torch.bfloat16
    _close(feats.grad.float().cpu().numpy(), ref_g['features'], 'features')
    if grads == 'all':
        _close(layer.blocks.grad.cpu().numpy(), ref_g['blocks'], 'blocks')
        np.testing.assert_allclose(layer.bias.grad.cpu().numpy(), ref_g['bias'], atol=1e-3, rtol=1e-4)


@pytest.mark.parametrize('stages', ['2', '3', '4'])
def test_fused_bulk_copy_variant_and_shallow_pipelines(cuda_device, monkeypatch, stages):
    """The per-row cp.async.bulk build of the kernel and short pipelines (1 - 4 producer warps) give the same sums."""
    N, R, E = 3000, 9, 40000
    t = _graph('uniform', N, R, E)
    monkeypatch.setenv('RGCN_FUSED_STAGES', stages)
    la, tp, fa, out_a, G, _ = _run(cuda_device, monkeypatch, True, N, R, t, True, 'all', fuse_rows='256', tma='gather4')
    lb, _, fb, out_b, _, _ = _run(cuda_device, monkeypatch, True, N, R, t, True, 'all', fuse_rows='256', tma='bulk')
    assert torch.equal(out_a, out_b) and torch.equal(fa.grad, fb.grad)
    ref_out, ref_g = orc.nc_layer(tp.numpy(), N, 2 * R + 1, _params_np(la), fa.detach().float().cpu().numpy(), True,
                                  G.cpu().numpy())
    _close(out_a.detach().cpu().numpy(), ref_out, 'out')
    _close(fa.grad.float().cpu().numpy(), ref_g['features'], 'features')


def test_fused_is_reproducible_and_agrees_with_two_phase(cuda_device, monkeypatch):
    N, R, E = 4000, 7, 20000
    t = _graph('uniform', N, R, E)
    la, _, fa, out_a, _, plan = _run(cuda_device, monkeypatch, True, N, R, t, True, 'all')
    assert plan.c.fuse_split[0] == 0
    lb, _, fb, out_b, _, _ = _run(cuda_device, monkeypatch, True, N, R, t, True, 'all')
    assert torch.equal(out_a, out_b), 'unsplit fused forward must be bit-reproducible'
    assert torch.equal(fa.grad, fb.grad)
    lc, _, fc, out_c, _, _ = _run(cuda_device, monkeypatch, False, N, R, t, True, 'all')
    scale = out_c.abs().max().item()
    assert (out_a - out_c).abs().max().item() <= 1e-2 * scale
    gs = fc.grad.float().abs().max().item()
    assert (fa.grad.float() - fc.grad.float()).abs().max().item() <= 2e-2 * gs


def test_fused_isolated_rows_get_bias(cuda_device, monkeypatch):
    """Rows without edges (whole empty row blocks included) still receive the bias; their gradient is zero."""
    from torch_rgcn_b200.layers import RelationalGraphConvolutionNC
    monkeypatch.setenv('RGCN_FUSED', '2')
    monkeypatch.setenv('RGCN_FUSE_ROWS', '64')
    N, Rp = 1000, 3
    tp = torch.tensor([[5, 0, 7], [5, 1, 900], [900, 2, 5], [999, 0, 0]])
    layer = RelationalGraphConvolutionNC(triples=tp, num_nodes=N, num_relations=Rp, in_features=64, out_features=64,
                                         decomposition={'type': 'block', 'num_blocks': 4},
                                         vertical_stacking=True).to(cuda_device)
    with torch.no_grad():
        layer.bias.copy_(torch.arange(64, device=cuda_device, dtype=torch.float32))
    x = torch.randn(N, 64, device=cuda_device).to(torch.bfloat16).requires_grad_(True)
    out = layer(x)
    out.backward(torch.ones_like(out))
    assert layer._plan_cache[1].fused_ok == [True, True]
    ref_out, ref_g = orc.nc_layer(tp.numpy(), N, Rp, _params_np(layer), x.detach().float().cpu().numpy(), True,
                                  np.ones((N, 64), np.float32))
    _close(out.detach().cpu().numpy(), ref_out, 'out')
    touched = torch.zeros(N, dtype=torch.bool)
    touched[tp[:, 0]] = True
    assert torch.equal(out[~touched.to(cuda_device)], layer.bias.detach().expand((~touched).sum().item(), 64))
    _close(x.grad.float().cpu().numpy(), ref_g['features'], 'features')
