"""GPU parity of the per-step link-prediction graph construction (torch_rgcn_b200/sampling.py -> csrc/sampling.cu):
the edge-neighbourhood kernel against the oracle pick for pick on shared uniforms, and the dropout / uniform rules."""
import numpy as np
import pytest
import torch

from conftest import load_golden
from oracle import sampling_oracle as so

pytestmark = pytest.mark.gpu


def _graph(seed, N, E, hub=False, loops=0):
    g = torch.Generator().manual_seed(seed)
    s, o = torch.randint(0, N, (E,), generator=g), torch.randint(0, N, (E,), generator=g)
    if hub:
        s[: E // 3] = 3                                   # a vertex with E / 3 incident edges (adjacency spans many warps-fuls)
    if loops:
        o[-loops:] = s[-loops:]                           # self-loops sit twice in their vertex's list
    t = torch.stack([s, torch.randint(0, 4, (E,), generator=g), o], 1)
    return torch.cat([t, t[:7]], 0)                       # repeated triples


def test_adjacency_matches_reference_order(cuda_device):
    from torch_rgcn_b200.sampling import EdgeNeighborhoodSampler
    t = _graph(0, 40, 150, loops=5)
    sm = EdgeNeighborhoodSampler(t.to(cuda_device), 40)
    adj = so.adjacency(t.numpy(), 40)
    ptr = sm.adj_ptr.cpu().numpy()
    assert ptr.tolist() == np.concatenate([[0], np.cumsum([len(a) for a in adj])]).tolist()
    flat = [x for a in adj for x in a]
    assert sm.adj_edge.cpu().numpy()[: len(flat)].tolist() == [e for e, _ in flat]
    assert sm.adj_other.cpu().numpy()[: len(flat)].tolist() == [o for _, o in flat]


@pytest.mark.parametrize('N,E,S,hub,loops', [
    (20, 60, 67, False, 4),            # no upper tree level (N <= 32), sample == every edge (incl. the repeats)
    (700, 2500, 900, True, 20),        # one upper level, a 800-entry adjacency list
    (5000, 8000, 2500, False, 30),     # two upper levels
    (40943, 30000, 6000, False, 50),   # WN18-sized node set: three upper levels, state in shared memory
    (40943, 30000, 3000, True, 0),
    (300000, 20000, 2500, False, 10),  # counts do not fit in shared memory: workspace placement
])
def test_edge_neighborhood_matches_oracle_pick_for_pick(cuda_device, N, E, S, hub, loops):
    from torch_rgcn_b200.sampling import EdgeNeighborhoodSampler
    t = _graph(N + E, N, E, hub, loops)
    u = torch.rand(S, 2, generator=torch.Generator().manual_seed(S))
    sm = EdgeNeighborhoodSampler(t.to(cuda_device), N)
    got = sm.sample_indices(S, u.to(cuda_device))
    sm.check()
    want = so.edge_neighborhood(t.numpy(), N, S, u.numpy())
    assert got.cpu().numpy().tolist() == want.tolist()
    rows = sm.sample(S, u.to(cuda_device))
    assert torch.equal(rows.cpu(), t[torch.as_tensor(want)])


def test_edge_neighborhood_distribution_matches_reference(cuda_device):
    """The reference's own histogram (tests/golden/sampling_hist.npz) against the kernel driven by torch's generator."""
    from torch_rgcn_b200.sampling import EdgeNeighborhoodSampler
    meta, d, _, _ = load_golden('sampling_hist')
    triples = d['triples'].tolist()
    sm = EdgeNeighborhoodSampler(torch.as_tensor(d['triples']).to(cuda_device), meta['N'])
    runs, S = 20000, meta['S']
    torch.manual_seed(3)
    u = torch.rand(runs, S, 2, device=cuda_device)
    idx = torch.stack([sm.sample_indices(S, u[r]) for r in range(runs)]).cpu().numpy()
    ref = d['hist'].astype(np.float64)
    got = np.zeros_like(ref)
    for row in idx:
        canon = []
        for e in row.tolist():                              # same tie-breaking of repeated triples as the fixture
            canon.append([i for i, tt in enumerate(triples) if tt == triples[e] and i not in canon][0])
        got[tuple(canon)] += 1
    p, q = ref / ref.sum(), got / got.sum()
    assert 0.5 * np.abs(p - q).sum() < 0.06
    np.testing.assert_allclose(got.sum((1, 2)) / runs, ref.sum((1, 2)) / ref.sum(), atol=0.012)


def test_sampling_api_and_dropout_rule(cuda_device):
    from torch_rgcn_b200 import sampling as sp
    t = _graph(5, 500, 3000).to(cuda_device)
    ent = {i: i for i in range(500)}
    torch.manual_seed(0)
    a = sp.select_sampling('edge-neighborhood')(t, sample_size=800, entities=ent)
    torch.manual_seed(0)
    b = sp.edge_neighborhood(t, 800, ent)
    assert a.shape == (800, 3) and a.is_cuda and torch.equal(a, b)          # reproducible under torch.manual_seed
    rows = {tuple(r) for r in t.cpu().tolist()}
    assert all(tuple(r) in rows for r in a.cpu().tolist())
    u = sp.select_sampling('uniform')(t, sample_size=1000)
    assert u.shape == (1000, 3) and all(tuple(r) in rows for r in u.cpu().tolist())
    with pytest.raises(ValueError):
        sp.uniform_sampling(t, t.size(0) + 1)
    with pytest.raises(ValueError):
        sp.EdgeNeighborhoodSampler(t, 500).sample(t.size(0) + 1)
    with pytest.raises(NotImplementedError):
        sp.select_sampling('snowball')
    # dropout: graph[perm][round(keep_prob * n):]
    perm = torch.randperm(t.size(0), generator=torch.Generator().manual_seed(1))
    for rate in (0.5, 0.2):
        got = sp.edge_dropout(t, rate, perm=perm.to(cuda_device))
        want = t.cpu()[torch.as_tensor(so.edge_dropout_rows(t.size(0), 1 - rate, perm.numpy()))]
        assert torch.equal(got.cpu(), want)
    assert sp.edge_dropout(t, 0.0) is t or torch.equal(sp.edge_dropout(t, 0.0), t)
    # the whole epoch's inputs (predict_links.py:123-148)
    sm = sp.EdgeNeighborhoodSampler(t, 500)
    graph, batch, lbl = sp.training_step_inputs(sm, 500, graph_batch_size=600, neg_sample_rate=3, edge_dropout_rate=0.5)
    assert graph.shape == (600 - round(0.5 * 600), 3) and batch.shape == (600 * 4, 3) and lbl.shape == (2400,)
    assert lbl[:600].all() and not lbl[600:].any()
    pos, neg = batch[:600], batch[600:].view(600, 3, 3)
    assert torch.equal(neg[:, :, 1], pos[:, None, 1].expand(600, 3))         # relations are never corrupted
    assert ((neg[:, :, 0] == pos[:, None, 0]) | (neg[:, :, 2] == pos[:, None, 2])).all()   # one end stays
    with pytest.raises(IndexError):
        bad = torch.tensor([0, t.size(0)], device=cuda_device)
        _, st = sp.take_triples(t, bad)
        if int(st.item()):
            raise IndexError
