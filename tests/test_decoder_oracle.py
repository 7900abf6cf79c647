"""CPU: the DistMult / negative-sampling oracle against fixtures generated from the unmodified reference
(tests/golden/make_golden.py: make_distmult, make_negsample)."""
import numpy as np
import pytest

from conftest import load_golden, golden_names
from oracle import distmult_oracle as dm


@pytest.mark.parametrize('name', golden_names('distmult_'))
def test_distmult_oracle_matches_reference(name):
    meta, d, params, grads = load_golden(name)
    biases = [params.get(k) for k in ('sbias', 'pbias', 'obias')]
    out = dm.score(d['triples'], d['nodes'], params['relations'], *biases)
    np.testing.assert_allclose(out, d['out'], atol=1e-5, rtol=1e-5)
    g = dm.score_backward(d['triples'], d['nodes'], params['relations'], d['G'], with_bias=biases[0] is not None)
    np.testing.assert_allclose(g['nodes'], grads['nodes'], atol=1e-5, rtol=1e-5)
    for k in params:
        np.testing.assert_allclose(g[k], grads[k], atol=1e-5, rtol=1e-5, err_msg=k)
    np.testing.assert_allclose(dm.penalty(d['triples'], d['nodes'], params['relations']), d['penalty'], rtol=1e-6)
    pg = dm.penalty_backward(d['triples'], d['nodes'], params['relations'], meta['penalty_grad'])
    np.testing.assert_allclose(pg['nodes'], d['pgrad_nodes'], atol=1e-7, rtol=1e-5)
    np.testing.assert_allclose(pg['relations'], d['pgrad_relations'], atol=1e-7, rtol=1e-5)


@pytest.mark.parametrize('name', golden_names('negsample_'))
def test_corruption_oracle_matches_reference(name):
    meta, d, _, _ = load_golden(name)
    out = dm.corrupt(d['batch'], d['head'], d['corruptions'])
    np.testing.assert_array_equal(out, d['out'])
    # exactly one endpoint of every triple was replaced; relations untouched (utils/misc.py:185)
    before = d['batch'].reshape(-1, 3)
    assert np.array_equal(out[:, 1], before[:, 1])
    assert np.all((out[:, 0] == before[:, 0]) | (out[:, 2] == before[:, 2]))


def test_decoder_module_mirrors_reference_surface():
    """Constructor, parameter names / shapes and initial values (same RNG draws) of the drop-in DistMult."""
    import torch
    from torch_rgcn_b200.layers import DistMult
    meta, d, params, _ = load_golden('distmult_bias')
    torch.manual_seed(meta['seed'] + 2)
    dec = DistMult(meta['R'], meta['d'], meta['N'], meta['R'], w_init='standard-normal', b_init=meta['b_init'])
    assert [n for n, _ in dec.named_parameters()] == ['relations', 'sbias', 'obias', 'pbias']
    for n, p in dec.named_parameters():
        np.testing.assert_array_equal(p.detach().numpy(), params[n], err_msg=n)
    plain = DistMult(3, 4, 5, 3)
    assert plain.sbias is None and plain.pbias is None and plain.obias is None
    with pytest.raises(RuntimeError):                      # no CPU fallback
        plain(torch.zeros(2, 3, dtype=torch.long), torch.zeros(5, 4))
