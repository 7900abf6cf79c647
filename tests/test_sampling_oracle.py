"""CPU: the sampling oracle against the unmodified reference's edge_neighborhood (statistically: the reference draws
from numpy's global generator) and the literal edge-dropout rule."""
import numpy as np

from conftest import load_golden
from oracle import sampling_oracle as so


def _canonical(triples, idx):
    """The reference returns triples, not edge indices; the fixture maps each back to the first not-yet-used equal
    triple (make_golden.make_sampling_hist), so repeated triples are told apart the same way here."""
    out = []
    for e in idx:
        out.append([i for i, t in enumerate(triples) if t == triples[e] and i not in out][0])
    return out


def test_edge_neighborhood_distribution_matches_reference():
    meta, d, _, _ = load_golden('sampling_hist')
    ref = d['hist'].astype(np.float64)
    runs = 40000
    rng = np.random.default_rng(5)
    got = np.zeros_like(ref)
    for _ in range(runs):
        idx = so.edge_neighborhood(d['triples'], meta['N'], meta['S'], rng.random((meta['S'], 2), dtype=np.float32))
        got[tuple(_canonical(d['triples'].tolist(), idx))] += 1
    assert ((ref > 0) == (got > 0)).mean() > 0.98                     # same support up to very rare outcomes
    p, q = ref / ref.sum(), got / got.sum()
    # two-sample chi-square over ordered samples with enough mass; the statistic's mean is its dof
    m = (ref + got) >= 40
    chi2 = (((ref[m] * np.sqrt(runs / meta['runs']) - got[m] * np.sqrt(meta['runs'] / runs)) ** 2) / (ref[m] + got[m])).sum()
    dof = int(m.sum()) - 1
    assert chi2 < dof + 5 * np.sqrt(2 * dof), (chi2, dof)
    assert 0.5 * np.abs(p - q).sum() < 0.05                           # total variation
    # first pick: a vertex uniform over the vertices with edges (nothing seen yet), then one of its entries
    first_ref, first_got = ref.sum((1, 2)) / ref.sum(), got.sum((1, 2)) / got.sum()
    np.testing.assert_allclose(first_got, first_ref, atol=0.01)


def test_edge_neighborhood_is_a_partial_permutation_and_follows_the_frontier():
    rng = np.random.default_rng(1)
    N, E = 50, 200
    t = np.stack([rng.integers(0, N, E), rng.integers(0, 3, E), rng.integers(0, N, E)], 1)
    idx = so.edge_neighborhood(t, N, E, rng.random((E, 2), dtype=np.float32))
    assert sorted(idx.tolist()) == list(range(E))                      # sample_size == |E|: every edge exactly once
    # after the first pick every later edge touches a vertex seen so far, unless the seen component is exhausted
    idx = so.edge_neighborhood(t, N, 20, rng.random((20, 2), dtype=np.float32))
    seen = set(t[idx[0], [0, 2]].tolist())
    for e in idx[1:]:
        assert t[e, 0] in seen or t[e, 2] in seen
        seen.update(t[e, [0, 2]].tolist())


def test_edge_dropout_rule():
    perm = np.random.default_rng(0).permutation(11)
    np.testing.assert_array_equal(so.edge_dropout_rows(11, 0.5, perm), perm[6:])   # round(5.5) = 6 (banker's)
    np.testing.assert_array_equal(so.edge_dropout_rows(10, 0.8, perm[:10]), perm[8:10])
