"""Filtered ranking evaluation of link-prediction models on the device.

Device counterparts of reference utils/misc.py:29-110: `generate_true_dict` (-> `TrueTripleFilter`, sorted key lists
instead of Python dictionaries), `filter_scores` + `evaluate` (-> `rank_triples` / `evaluate`).  The reference scores
a (batch, num_nodes, 3) tensor of candidate triples through the WHOLE model for every evaluation batch (the encoder
runs 2 * ceil(|test| / batch) times, utils/misc.py:86); here the encoder runs once and the candidate scores are never
materialised (torch_rgcn_b200/csrc/ranking.cu).  No CPU fallback.
"""
import torch

from . import _lib


class TrueTripleFilter:
    """Known true completions of every (p, o) and (s, p) pair, for filtered ranking.

    Built from all known triples (train + valid + test) like reference `generate_true_dict` (utils/misc.py:29-37);
    also accepts that function's `(heads, tails)` dictionary pair."""

    def __init__(self, all_triples, num_nodes, num_rels, device=None):
        if isinstance(all_triples, tuple) and len(all_triples) == 2 and isinstance(all_triples[0], dict):
            heads, _ = all_triples                      # {(p, o): [s, ...]} holds every triple once per occurrence
            all_triples = [(s, p, o) for (p, o), ss in heads.items() for s in ss]
        t = torch.as_tensor(all_triples, dtype=torch.long)
        if device is None:
            _lib.require_cuda()
            device = t.device if t.is_cuda else torch.device('cuda', torch.cuda.current_device())
        t = t.to(device).reshape(-1, 3).contiguous()
        self.num_nodes, self.num_rels, self.num_true, self.device = int(num_nodes), int(num_rels), t.size(0), t.device
        M = max(self.num_true, 1)
        self.keys = [torch.empty(M, dtype=torch.int64, device=t.device) for _ in range(2)]     # [tail lists, head lists]
        self.vals = [torch.empty(M, dtype=torch.int32, device=t.device) for _ in range(2)]
        status = torch.zeros(1, dtype=torch.int32, device=t.device)
        ws_bytes = _lib.lib.rgcn_rank_filter_workspace_bytes(self.num_true)
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=t.device)
        with torch.cuda.device(t.device):
            for head in (0, 1):
                _lib.check(_lib.lib.rgcn_rank_build_filter(_lib.ptr(t), self.num_true, self.num_nodes, self.num_rels, head,
                                                           _lib.ptr(self.keys[head]), _lib.ptr(self.vals[head]),
                                                           _lib.ptr(status), _lib.ptr(ws), ws_bytes, _lib.stream_ptr()))
        bad = int(status.item())
        assert bad == 0, f'{bad // 2} known triples index a node >= {num_nodes} or a relation >= {num_rels}'


def rank_triples(queries, nodes, relations, head, true_filter=None, sbias=None, pbias=None, obias=None):
    """Rank of the true head (head=True) or tail of every query triple among all num_nodes completions, int64 (T,).

    rank = #{scores > true} + (#{scores == true} - 1) // 2 + 1 after removing the other known true completions
    (utils/misc.py:39-58, :91-101).  Scores are DistMult scores of `nodes` (N, d) and `relations` (R, d)."""
    _lib.require_cuda(queries, nodes, relations, sbias, pbias, obias)
    assert queries.dtype == torch.long and queries.dim() == 2 and queries.size(1) == 3
    q = queries.contiguous()
    f32 = lambda x: None if x is None else x.detach().to(torch.float32).contiguous()     # noqa: E731
    nodes, relations, sbias, pbias, obias = (f32(x) for x in (nodes, relations, sbias, pbias, obias))
    dev = nodes.device
    T, N, R, d = q.size(0), nodes.size(0), relations.size(0), nodes.size(1)
    ranks = torch.empty(T, dtype=torch.int64, device=dev)
    status = torch.zeros(1, dtype=torch.int32, device=dev)
    keys = vals = None
    M = 0
    if true_filter is not None:
        assert true_filter.num_nodes == N and true_filter.device == dev
        keys, vals, M = true_filter.keys[1 if head else 0], true_filter.vals[1 if head else 0], true_filter.num_true
    with torch.cuda.device(dev):
        for lo in range(0, T, 1 << 20):                 # bounded workspace (the query matrix is T x d floats)
            hi = min(T, lo + (1 << 20))
            ws_bytes = _lib.lib.rgcn_rank_workspace_bytes(hi - lo, d)
            ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
            _lib.check(_lib.lib.rgcn_rank_triples(_lib.ptr(q[lo:hi]), hi - lo, 1 if head else 0, _lib.ptr(nodes), N,
                                                  _lib.ptr(relations), R, d, _lib.ptr(sbias), _lib.ptr(pbias),
                                                  _lib.ptr(obias), _lib.ptr(keys), _lib.ptr(vals), M, _lib.ptr(ranks[lo:hi]),
                                                  _lib.ptr(status), _lib.ptr(ws), ws_bytes, _lib.stream_ptr()))
    bad = int(status.item())
    if bad:
        raise IndexError(f'{bad} query triples index a node >= {N} or a relation >= {R}')
    return ranks


def evaluate(model, graph, test_set, true_triples, num_nodes, batch_size=16, hits_at_k=[1, 3, 10], filter_candidates=True,
             verbose=True):
    """Drop-in for reference utils/misc.py:60-110: returns (mrr, hits, ranks) with head ranks first, then tail ranks.

    `model` must expose `encode(graph) -> (num_nodes, d)` node embeddings and `scoring_function` (a DistMult).  The
    encoder runs once (the reference re-runs it per batch; `batch_size` and `verbose` are accepted and ignored).
    `true_triples` is a TrueTripleFilter or the reference's `(heads, tails)` dictionary pair."""
    dec = model.scoring_function
    with torch.no_grad():
        x = model.encode(graph)
        test_set = torch.as_tensor(test_set, dtype=torch.long).to(x.device)
        filt = None
        if filter_candidates:
            filt = true_triples if isinstance(true_triples, TrueTripleFilter) else \
                TrueTripleFilter(true_triples, num_nodes, dec.relations.size(0), device=x.device)
        ranks = []
        for head in (True, False):
            ranks.extend(rank_triples(test_set, x, dec.relations, head, filt, dec.sbias, dec.pbias, dec.obias).tolist())
    mrr = sum(1.0 / r for r in ranks) / len(ranks)
    hits = tuple(sum(1.0 if r <= k else 0.0 for r in ranks) / len(ranks) for k in hits_at_k)
    return mrr, hits, ranks
