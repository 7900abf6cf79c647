"""CPU restatement (numpy) of the reference's filtered ranking evaluation, utils/misc.py:29-110.

TEST INFRASTRUCTURE ONLY (tests/, smoke(), bench cpu_baseline).  Pinned against tests/golden/ranking_*.npz, which the
unmodified reference `evaluate` produced (generator: tests/golden/make_golden.py: make_ranking).
"""
import numpy as np

from . import distmult_oracle as dm


def true_dicts(all_triples):
    """utils/misc.py:29-37"""
    heads, tails = {}, {}
    for s, p, o in np.asarray(all_triples).tolist():
        heads.setdefault((p, o), []).append(s)
        tails.setdefault((s, p), []).append(o)
    return heads, tails


def ranks(test, nodes, relations, all_triples=None, sbias=None, pbias=None, obias=None, dtype=np.float32):
    """utils/misc.py:60-101: head ranks of every test triple, then tail ranks.  Scores in `dtype` like the reference
    (fp32: ties are decided at fp32 resolution)."""
    test = np.asarray(test)
    N = nodes.shape[0]
    heads, tails = true_dicts(all_triples) if all_triples is not None else (None, None)
    nodes, relations = np.asarray(nodes, dtype), np.asarray(relations, dtype)
    out = []
    for head in (True, False):
        for s, p, o in test.tolist():
            cand = np.arange(N)
            trip = np.stack([cand, np.full(N, p), np.full(N, o)], 1) if head else np.stack([np.full(N, s), np.full(N, p), cand], 1)
            sc = (nodes[trip[:, 0]] * relations[trip[:, 1]] * nodes[trip[:, 2]]).sum(-1)      # layers.py:92
            if sbias is not None:
                sc = sc + (np.asarray(sbias, dtype)[trip[:, 0]] + np.asarray(pbias, dtype)[trip[:, 1]] + np.asarray(obias, dtype)[trip[:, 2]])
            target = s if head else o
            if heads is not None:
                for c in (heads[p, o] if head else tails[s, p]):                                 # filter_scores :47-52
                    if c != target:
                        sc[c] = -np.inf
            true = sc[target]
            out.append(int((sc > true).sum() + ((sc == true).sum() - 1) // 2 + 1))
    return out


def metrics(rank_list, hits_at_k=(1, 3, 10)):
    """utils/misc.py:103-109"""
    mrr = sum(1.0 / r for r in rank_list) / len(rank_list)
    return mrr, tuple(sum(1.0 if r <= k else 0.0 for r in rank_list) / len(rank_list) for k in hits_at_k)
