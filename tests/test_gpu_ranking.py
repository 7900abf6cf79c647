"""GPU parity of the filtered ranking evaluation (torch_rgcn_b200/evaluation.py -> csrc/ranking.cu)."""
import numpy as np
import pytest
import torch

from conftest import load_golden, golden_names
from oracle import ranking_oracle as ro

pytestmark = pytest.mark.gpu


class _Model:
    """What evaluate() needs from a link-prediction model: encode(graph) and scoring_function."""

    def __init__(self, dec, x):
        self.scoring_function, self.x = dec, x

    def encode(self, graph):
        return self.x


def _decoder(meta, params, dev):
    from torch_rgcn_b200.layers import DistMult
    dec = DistMult(meta['R'], meta['d'], meta['N'], meta['R'], b_init=meta['b_init']).to(dev)
    with torch.no_grad():
        for n, p in dec.named_parameters():
            p.copy_(torch.as_tensor(params[n]))
    return dec


@pytest.mark.parametrize('name', golden_names('ranking_'))
def test_evaluate_matches_reference_fixture(cuda_device, name):
    """Integer-valued fixtures: every score is exact in fp32, so ranks (ties included) must be identical.  Float
    fixture: scores differ from the reference's in the last bits (q = p * o first, fma chain), which can only move a
    rank when two scores agree to ~1e-7 relative; none do in this fixture."""
    from torch_rgcn_b200.evaluation import evaluate, TrueTripleFilter
    meta, d, params, _ = load_golden(name)
    model = _Model(_decoder(meta, params, cuda_device), torch.as_tensor(d['nodes']).to(cuda_device))
    true = ro.true_dicts(d['known'])                                        # the reference's dictionary pair
    for tag, filt in (('filtered', True), ('raw', False)):
        mrr, hits, ranks = evaluate(model, None, torch.as_tensor(d['test']), true, meta['N'], filter_candidates=filt,
                                    verbose=False)
        assert ranks == d['ranks_' + tag].tolist(), tag
        np.testing.assert_allclose(mrr, d['mrr_' + tag], rtol=1e-12)
        np.testing.assert_allclose(hits, d['hits_' + tag], rtol=1e-12)
    filt = TrueTripleFilter(torch.as_tensor(d['known']), meta['N'], meta['R'])
    _, _, ranks = evaluate(model, None, torch.as_tensor(d['test']), filt, meta['N'], verbose=False)
    assert ranks == d['ranks_filtered'].tolist()


@pytest.mark.parametrize('N,R,dim,T,bias', [(3000, 11, 128, 700, False), (1000, 5, 50, 300, True), (130, 3, 7, 90, False)])
def test_ranks_match_oracle_at_size(cuda_device, N, R, dim, T, bias):
    """Integer embeddings (exact scores, many ties) at sizes that span several candidate / query tiles, odd widths and
    a width that is not a multiple of the k-chunk."""
    from torch_rgcn_b200.evaluation import rank_triples, TrueTripleFilter
    g = torch.Generator().manual_seed(9)
    known = torch.stack([torch.randint(0, N, (6 * T,), generator=g), torch.randint(0, R, (6 * T,), generator=g),
                         torch.randint(0, N, (6 * T,), generator=g)], 1)
    known[:T, 1:] = known[0, 1:]                     # one (p, o) pair with T known heads
    known = torch.cat([known, known[:50]], 0)
    test = known[torch.randperm(known.size(0), generator=g)[:T]]
    x = torch.randint(-2, 3, (N, dim), generator=g).float()
    rel = torch.randint(-2, 3, (R, dim), generator=g).float()
    b = [torch.randint(-3, 4, (n,), generator=g).float() for n in (N, R, N)] if bias else [None, None, None]
    filt = TrueTripleFilter(known, N, R)
    dev = cuda_device
    cu = lambda t: None if t is None else t.to(dev)      # noqa: E731
    got = []
    for head in (True, False):
        got += rank_triples(test.to(dev), x.to(dev), rel.to(dev), head, filt, cu(b[0]), cu(b[1]), cu(b[2])).tolist()
    want = ro.ranks(test.numpy(), x.numpy(), rel.numpy(), known.numpy(), *[None if t is None else t.numpy() for t in b])
    assert got == want
    raw = rank_triples(test.to(dev), x.to(dev), rel.to(dev), True, None, cu(b[0]), cu(b[1]), cu(b[2])).tolist()
    assert raw == ro.ranks(test.numpy(), x.numpy(), rel.numpy(), None, *[None if t is None else t.numpy() for t in b])[:T]


def test_float_scores_rank_like_the_oracle(cuda_device):
    """Random fp32 embeddings at WN18-like width: ranks agree with the fp64 oracle except where two scores are within
    rounding of each other (allowed: a handful of ranks off by one)."""
    from torch_rgcn_b200.evaluation import rank_triples
    g = torch.Generator().manual_seed(4)
    N, R, dim, T = 5000, 18, 128, 400
    test = torch.stack([torch.randint(0, N, (T,), generator=g), torch.randint(0, R, (T,), generator=g),
                        torch.randint(0, N, (T,), generator=g)], 1)
    x, rel = torch.randn(N, dim, generator=g), torch.randn(R, dim, generator=g)
    got = rank_triples(test.to(cuda_device), x.to(cuda_device), rel.to(cuda_device), False).cpu().numpy()
    want = np.array(ro.ranks(test.numpy(), x.numpy(), rel.numpy(), None, dtype=np.float64)[T:])
    assert np.abs(got - want).max() <= 1 and (got != want).mean() < 0.02


def test_rank_bad_indices_raise(cuda_device):
    from torch_rgcn_b200.evaluation import rank_triples
    x, rel = torch.randn(10, 4, device=cuda_device), torch.randn(2, 4, device=cuda_device)
    with pytest.raises(IndexError):
        rank_triples(torch.tensor([[0, 2, 1]], device=cuda_device), x, rel, True)
    assert rank_triples(torch.zeros(0, 3, dtype=torch.long, device=cuda_device), x, rel, True).numel() == 0


def test_compression_relation_predictor_matches_reference(cuda_device):
    """c-rgcn model (reference models.py:208-245, eval mode, node_embedding == hidden1_size — the one configuration in
    which the shipped class runs): scores, penalty and every parameter gradient against the reference's fixture."""
    from torch_rgcn_b200.models import CompressionRelationPredictor
    meta, d, params, grads = load_golden('lpmodel_crp')
    torch.manual_seed(0)
    model = CompressionRelationPredictor(nnodes=meta['N'], nrel=meta['R'], encoder_config=meta['encoder'],
                                         decoder_config=meta['decoder']).eval()
    assert sorted(n for n, _ in model.named_parameters()) == sorted(params)
    model.to(cuda_device)
    with torch.no_grad():
        for n, p in model.named_parameters():
            p.copy_(torch.as_tensor(params[n]))
    graph, batch = torch.as_tensor(d['graph']).to(cuda_device), torch.as_tensor(d['batch']).to(cuda_device)
    scores, penalty = model(graph, batch)
    ((scores * torch.as_tensor(d['G']).to(cuda_device)).sum() + meta['penalty_weight'] * penalty).backward()
    np.testing.assert_allclose(scores.detach().cpu().numpy(), d['out'], atol=1e-4, rtol=1e-4)
    np.testing.assert_allclose(penalty.detach().cpu().numpy(), d['penalty'], atol=1e-5, rtol=1e-5)
    for n, p in model.named_parameters():
        np.testing.assert_allclose(p.grad.cpu().numpy(), grads[n], atol=1e-4, rtol=1e-4, err_msg=n)
    # evaluate() drives the same model through encode() once
    from torch_rgcn_b200.evaluation import evaluate
    mrr, hits, ranks = evaluate(model, graph, batch[:16], ro.true_dicts(d['batch']), meta['N'], verbose=False)
    x = model.encode(graph).detach().cpu().numpy()
    want = ro.ranks(d['batch'][:16], x, params['scoring_function.relations'], d['batch'], dtype=np.float64)
    assert np.abs(np.array(ranks) - np.array(want)).max() <= 1
