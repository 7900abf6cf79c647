"""CPU: the ranking oracle against what the unmodified reference `evaluate` produced (tests/golden/ranking_*.npz)."""
import numpy as np
import pytest

from conftest import load_golden, golden_names
from oracle import ranking_oracle as ro


@pytest.mark.parametrize('name', golden_names('ranking_'))
def test_ranking_oracle_matches_reference(name):
    meta, d, params, _ = load_golden(name)
    biases = {k: params.get(k) for k in ('sbias', 'pbias', 'obias')}
    for tag, known in (('filtered', d['known']), ('raw', None)):
        r = ro.ranks(d['test'], d['nodes'], params['relations'], known, **biases)
        assert r == d['ranks_' + tag].tolist(), tag
        mrr, hits = ro.metrics(r)
        np.testing.assert_allclose(mrr, d['mrr_' + tag], rtol=1e-12)
        np.testing.assert_allclose(hits, d['hits_' + tag], rtol=1e-12)
    if meta['integer']:
        assert len(set(d['ranks_raw'].tolist())) < len(d['ranks_raw'])      # the integer fixtures do exercise ties
