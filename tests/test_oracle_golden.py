"""Pin oracle/rgcn_oracle.py: reference unit-test vectors + fixtures produced by the real reference.

CPU only.  Vectors are restated from the reference's tests (file:line cited per test).
"""
import numpy as np
import pytest

from oracle import rgcn_oracle as orc
from conftest import load_golden, golden_names


def test_add_inverse_and_self_reference_vector():
    # reference tests/test_utils.py:5-25 (negative object ids are deliberate there)
    t = np.array([[0, 0, -1], [1, 1, -2], [2, 2, -3]])
    exp = np.array([[0, 0, -1], [1, 1, -2], [2, 2, -3], [-1, 3, 0], [-2, 4, 1], [-3, 5, 2],
                    [0, 6, 0], [1, 6, 1], [2, 6, 2]])
    assert np.array_equal(orc.add_inverse_and_self(t, 3, 3), exp)


STACK_TRIPLES = np.array([[0, 0, 3], [1, 1, 4], [2, 2, 5], [3, 3, 0], [4, 4, 1], [5, 5, 2],
                          [0, 6, 0], [1, 6, 1], [2, 6, 2], [3, 6, 3], [4, 6, 4], [5, 6, 5]])
STACK_VER = np.array([[0, 3], [10, 4], [20, 5], [30, 0], [40, 1], [50, 2], [54, 0], [55, 1], [56, 2],
                      [57, 3], [58, 4], [59, 5]])
STACK_HOR = np.array([[0, 3], [1, 13], [2, 23], [3, 27], [4, 37], [5, 47], [0, 54], [1, 55], [2, 56],
                      [3, 57], [4, 58], [5, 59]])


def test_stack_matrices_reference_vector():
    # reference tests/test_utils.py:28-84
    ind, size = orc.stack_matrices(STACK_TRIPLES, 9, 7, vertical_stacking=True)
    assert np.array_equal(ind, STACK_VER) and size == (63, 9)
    ind, size = orc.stack_matrices(STACK_TRIPLES, 9, 7, vertical_stacking=False)
    assert np.array_equal(ind, STACK_HOR) and size == (9, 63)


SUM_VER_IND = np.array([[0, 0], [0, 1], [0, 2], [4, 1], [8, 2], [7, 2]])
SUM_HOR_IND = np.array([[0, 0], [1, 0], [2, 0], [3, 0], [1, 4], [2, 8], [2, 7]])


def test_sum_sparse_reference_vector():
    # reference tests/test_utils.py:87-123
    v = np.ones(6, np.float32)
    out = v / orc.sum_sparse(SUM_VER_IND, v, (9, 3), True).astype(np.float32)
    assert np.array_equal(out, np.array([1 / 3, 1 / 3, 1 / 3, 1, 1, 1], np.float32))
    v = np.ones(7, np.float32)
    out = v / orc.sum_sparse(SUM_HOR_IND, v, (4, 9), False).astype(np.float32)
    assert np.array_equal(out, np.array([1 / 4, 1 / 4, 1 / 4, 1 / 4, 1, 1, 1], np.float32))


ARR_ROW_IND = np.array([[0, 0], [0, 1], [1, 1], [4, 2], [5, 0], [5, 1], [6, 0], [7, 0], [7, 1], [9, 2],
                        [10, 2], [11, 1], [12, 0], [13, 1], [14, 2]])
ARR_ROW_VAL = np.array([1, 1, 1, 2, 2, 1, 1, 1, 1, 2, 1, 2, 1, 1, 1], np.float64)
ARR_COL_IND = np.array([[0, 0], [0, 1], [1, 1], [2, 3], [2, 4], [1, 5], [0, 6], [1, 6], [1, 7], [2, 10],
                        [1, 11], [0, 11], [0, 12], [1, 13], [2, 14]])
ARR_COL_VAL = np.array([1, 1, 1, 2, 1, 2, 1, 1, 1, 2, 1, 2, 1, 1, 1], np.float64)
ARR_EXPECT = np.array([2, 2, 1, 2, 3, 3, 1, 2, 2, 2, 1, 2, 1, 1, 1], np.float64)


def test_arrange_matrix_reference_vector():
    # reference tests/test_utils.py:170-220 (not collected upstream: its long-valued spmm raises; values as floats)
    assert np.array_equal(orc.sum_sparse(ARR_ROW_IND, ARR_ROW_VAL, (15, 3), True), ARR_EXPECT)
    sums = orc.sum_sparse(ARR_COL_IND, ARR_COL_VAL, (3, 15), False)
    r = (len(ARR_COL_VAL) - 3) // 2
    sums = np.concatenate([sums[r:2 * r], sums[:r], sums[2 * r:]])
    assert np.array_equal(sums, ARR_EXPECT)


def _check(meta, d, params, grads, out, og, atol=2e-5):
    np.testing.assert_allclose(out, d['out'], atol=atol, rtol=1e-5)
    for k, g in grads.items():
        assert k in og, k
        np.testing.assert_allclose(og[k], g, atol=atol, rtol=1e-4, err_msg=k)


@pytest.mark.parametrize('name', golden_names('nc_'))
def test_oracle_matches_reference_nc(name):
    meta, d, params, grads = load_golden(name)
    out, og = orc.nc_layer(d['triples_plus'], meta['N'], meta['num_relations'], params, d.get('features'),
                           meta['vertical'], d['G'])
    _check(meta, d, params, grads, out, og)


@pytest.mark.parametrize('name', golden_names('lp_'))
def test_oracle_matches_reference_lp(name):
    meta, d, params, grads = load_golden(name)
    out, og = orc.lp_layer(d['triples'], meta['N'], meta['num_relations'], params, d['features'],
                           meta['vertical'], d['G'], keep=d.get('keep'), self_mask=d.get('self_mask'))
    _check(meta, d, params, grads, out, og)


# ---- the torch CPU port that bench.py times as the reference's CPU path ---------------------------------
def _port_check(name, fwd):
    import torch
    meta, d, params, grads = load_golden(name)
    tp = {k: torch.tensor(v, requires_grad=True) for k, v in params.items()}
    feats = torch.tensor(d['features'], requires_grad=True) if 'features' in d else None
    out = fwd(meta, d, tp, feats)
    out.backward(torch.tensor(d['G']))
    np.testing.assert_allclose(out.detach().numpy(), d['out'], atol=1e-5, rtol=1e-5)
    for k, g in grads.items():
        got = feats.grad if k == 'features' else tp[k].grad
        np.testing.assert_allclose(got.numpy(), g, atol=1e-5, rtol=1e-4, err_msg=k)


@pytest.mark.parametrize('name', golden_names('nc_'))
def test_port_matches_reference_nc(name):
    import torch
    from oracle import torch_sparse_port as port
    _port_check(name, lambda meta, d, p, x: port.nc_forward(torch.tensor(d['triples_plus']), meta['N'],
                                                            meta['num_relations'], p, x, meta['vertical']))


@pytest.mark.parametrize('name', [n for n in golden_names('lp_') if 'train' not in n])
def test_port_matches_reference_lp(name):
    import torch
    from oracle import torch_sparse_port as port
    _port_check(name, lambda meta, d, p, x: port.lp_forward(torch.tensor(d['triples']), meta['N'],
                                                            meta['num_relations'], p, x, meta['vertical']))


# ---- whole models: the oracle layers composed like reference models.py:192-200 / :288-296 --------------------
def _split(params, prefix):
    return {k[len(prefix):]: v for k, v in params.items() if k.startswith(prefix)}


@pytest.mark.parametrize('name', golden_names('model_'))
def test_oracle_matches_reference_models(name):
    meta, d, params, grads = load_golden(name)
    N, R = meta['N'], meta['R']
    tp = orc.add_inverse_and_self(d['triples'], N, R)
    Rp = 2 * R + 1
    G = d['G']
    if meta['cls'] == 'EmbeddingNodeClassifier':
        emb = params['node_embeddings']
        h, _ = orc.nc_layer(tp, N, Rp, _split(params, 'rgcn_no_hidden.'), emb, False, np.zeros((N, emb.shape[1])))
        a = np.maximum(h, 0)
        out, g1 = orc.nc_layer(tp, N, Rp, _split(params, 'rgc1.'), a, False, G)
        _, g0 = orc.nc_layer(tp, N, Rp, _split(params, 'rgcn_no_hidden.'), emb, False, g1['features'] * (h > 0))
        got = {'rgc1.' + k: v for k, v in g1.items() if k != 'features'}
        got.update({'rgcn_no_hidden.' + k: v for k, v in g0.items() if k != 'features'})
        got['node_embeddings'] = g0['features']
    elif meta['kwargs'].get('nlayers') == 1:
        out, g1 = orc.nc_layer(tp, N, Rp, _split(params, 'rgc1.'), None, False, G)
        got = {'rgc1.' + k: v for k, v in g1.items() if v is not None}
    else:
        h, _ = orc.nc_layer(tp, N, Rp, _split(params, 'rgc1.'), None, False, np.zeros((N, 16)))
        a = np.maximum(h, 0)
        out, g2 = orc.nc_layer(tp, N, Rp, _split(params, 'rgc2.'), a, True, G)
        _, g1 = orc.nc_layer(tp, N, Rp, _split(params, 'rgc1.'), None, False, g2['features'] * (h > 0))
        got = {'rgc2.' + k: v for k, v in g2.items() if k != 'features'}
        got.update({'rgc1.' + k: v for k, v in g1.items() if v is not None})
    np.testing.assert_allclose(out, d['out'], atol=2e-5, rtol=1e-5)
    for k, g in grads.items():
        np.testing.assert_allclose(got[k], g, atol=2e-5, rtol=1e-4, err_msg=k)
