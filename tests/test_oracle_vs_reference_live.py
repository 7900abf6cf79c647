"""CPU, build container only: the oracle against the UNMODIFIED reference layers imported from /root/reference, on
randomised configurations beyond the committed fixtures.  Skipped wherever the reference checkout is absent (the GPU box):
the committed fixtures (tests/golden) remain the portable pin."""
import os
import sys
import warnings

import numpy as np
import pytest
import torch

from oracle import rgcn_oracle as orc

REF = '/root/reference'
pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, 'torch_rgcn')), reason='reference checkout not present')


def _ref():
    if REF not in sys.path:
        sys.path.insert(0, REF)
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        from torch_rgcn import layers, utils
    return layers, utils


def _triples(rng, N, R, E):
    return np.stack([rng.integers(0, N, E), rng.integers(0, R, E), rng.integers(0, N, E)], 1)


NC_CASES = []
for seed in range(12):
    r = np.random.default_rng(1000 + seed)
    kind = ['none', 'basis', 'block', 'diag', 'featureless', 'featureless_basis'][seed % 6]
    # featureless and diagonal layers only run horizontally upstream (layers.py:286-291 multiply the (R'N x N) vertical
    # adjacency with an (R'N, O) operand and raise); the CUDA layers raise for the same combinations
    vertical = bool(r.integers(0, 2)) and kind in ('none', 'basis', 'block')
    NC_CASES.append((seed, kind, vertical, bool(r.integers(0, 2))))


@pytest.mark.parametrize('seed,kind,vertical,shuffle', NC_CASES)
def test_nc_oracle_matches_live_reference(seed, kind, vertical, shuffle):
    layers, utils = _ref()
    rng = np.random.default_rng(seed)
    N, R, E = int(rng.integers(5, 40)), int(rng.integers(1, 6)), int(rng.integers(1, 150))
    nb = int(rng.integers(1, 4))
    in_f, out_f = nb * int(rng.integers(1, 5)), nb * int(rng.integers(1, 5))
    dec = {'basis': {'type': 'basis', 'num_bases': int(rng.integers(1, 4))}, 'featureless_basis': {'type': 'basis', 'num_bases': 2},
           'block': {'type': 'block', 'num_blocks': nb}}.get(kind)
    if kind.startswith('featureless'):
        in_f = None
    tp = utils.add_inverse_and_self(torch.as_tensor(_triples(rng, N, R, E)), N, R)
    if shuffle:
        tp = tp[torch.as_tensor(rng.permutation(tp.size(0)))]
    torch.manual_seed(seed)
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        layer = layers.RelationalGraphConvolutionNC(triples=tp, num_nodes=N, num_relations=2 * R + 1, in_features=in_f,
                                                    out_features=out_f, decomposition=dec, vertical_stacking=vertical,
                                                    diag_weight_matrix=(kind == 'diag'))
        if layer.bias is not None:
            with torch.no_grad():
                layer.bias.normal_()
        x = None if in_f is None else torch.randn(N, in_f, requires_grad=True)
        out = layer(x) if x is not None else layer()
        G = torch.randn(out.shape)
        out.backward(G)
    params = {n: p.detach().numpy() for n, p in layer.named_parameters()}
    got, og = orc.nc_layer(tp.numpy(), N, 2 * R + 1, params, None if x is None else x.detach().numpy(), vertical, G.numpy())
    np.testing.assert_allclose(got, out.detach().numpy(), atol=3e-5, rtol=1e-4)
    if x is not None:
        np.testing.assert_allclose(og['features'], x.grad.numpy(), atol=3e-5, rtol=1e-4)
    for n, p in layer.named_parameters():
        np.testing.assert_allclose(og[n], p.grad.numpy(), atol=3e-5, rtol=1e-4, err_msg=n)


@pytest.mark.parametrize('seed', range(8))
def test_lp_oracle_matches_live_reference(seed):
    layers, _ = _ref()
    rng = np.random.default_rng(50 + seed)
    N, R, E = int(rng.integers(5, 40)), int(rng.integers(1, 6)), int(rng.integers(0, 120))
    nb = int(rng.integers(1, 4))
    in_f, out_f = nb * int(rng.integers(1, 5)), nb * int(rng.integers(1, 5))
    kind = ['none', 'basis', 'block', 'none'][seed % 4]
    vertical = kind != 'block' and bool(rng.integers(0, 2))
    dec = {'basis': {'type': 'basis', 'num_bases': 2}, 'block': {'type': 'block', 'num_blocks': nb}}.get(kind)
    t = torch.as_tensor(_triples(rng, N, R, max(E, 1)))[:E]
    torch.manual_seed(seed)
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        layer = layers.RelationalGraphConvolutionLP(num_nodes=N, num_relations=2 * R + 1, in_features=in_f,
                                                    out_features=out_f, decomposition=dec, vertical_stacking=vertical,
                                                    w_init='glorot-normal', b_init='normal').eval()
        x = torch.randn(N, in_f, requires_grad=True)
        out = layer(t, x)
        G = torch.randn(out.shape)
        out.backward(G)
    params = {n: p.detach().numpy() for n, p in layer.named_parameters()}
    got, og = orc.lp_layer(t.numpy(), N, 2 * R + 1, params, x.detach().numpy(), vertical, G.numpy())
    np.testing.assert_allclose(got, out.detach().numpy(), atol=3e-5, rtol=1e-4)
    np.testing.assert_allclose(og['features'], x.grad.numpy(), atol=3e-5, rtol=1e-4)
    for n, p in layer.named_parameters():
        np.testing.assert_allclose(og[n], p.grad.numpy(), atol=3e-5, rtol=1e-4, err_msg=n)
