"""CPU restatement (numpy / pure Python) of the reference's per-step graph construction for link prediction:
edge-neighbourhood sampling (utils/misc.py:125-172), uniform sampling (:121-123) and general edge dropout
(experiments/predict_links.py:143-148).

TEST INFRASTRUCTURE ONLY (tests/, smoke(), bench cpu_baseline).

The reference draws from numpy's global generator (`np.random.choice`), which no device kernel can replay, so the
restatement takes the randomness as an explicit array of uniforms, two per pick:
    u[i, 0] picks the vertex      (inverse CDF over the integer weights, what np.random.choice(p=...) does with floats)
    u[i, 1] picks the incident edge (the j-th not-yet-picked entry of the vertex's adjacency list; the reference
            redraws from the whole list until it hits an unpicked entry, utils/misc.py:155-158, which is the same
            distribution)
Pinned statistically: tests/golden/sampling_hist.npz holds the histogram of ordered samples the unmodified reference
function produced over many runs on a small graph (generator: tests/golden/make_golden.py: make_sampling_hist); the
restatement must reproduce it within sampling error (tests/test_sampling_oracle.py).  The CUDA kernel is then compared
with this restatement pick for pick on shared uniforms.
"""
import numpy as np


def adjacency(triples, num_nodes):
    """utils/misc.py:129-132: adj[v] = [(edge index, other end)] in edge order, the subject's entry first."""
    adj = [[] for _ in range(num_nodes)]
    for i, (s, _, o) in enumerate(np.asarray(triples).tolist()):
        adj[s].append((i, o))
        adj[o].append((i, s))
    return adj


def edge_neighborhood(triples, num_nodes, sample_size, uniforms):
    """utils/misc.py:125-172 with explicit uniforms (sample_size, 2) in [0, 1).  Returns the picked edge indices."""
    triples = np.asarray(triples)
    adj = adjacency(triples, num_nodes)
    counts = np.array([len(a) for a in adj], dtype=np.int64)          # sample_counts, :139
    picked = np.zeros(len(triples), dtype=bool)
    seen = np.zeros(num_nodes, dtype=bool)
    u = np.asarray(uniforms, dtype=np.float32).astype(np.float64).reshape(-1, 2)
    out = np.zeros(sample_size, dtype=np.int64)
    for i in range(sample_size):
        w = counts * seen                                             # :144
        if w.sum() == 0:                                              # :146-148
            w = (counts > 0).astype(np.int64)
        total = int(w.sum())
        assert total > 0, 'sample_size exceeds the number of edges'
        target = min(int(u[i, 0] * total), total - 1)
        v = int(np.searchsorted(np.cumsum(w), target, side='right'))  # :151
        seen[v] = True
        c = int(counts[v])
        j = min(int(u[i, 1] * c), c - 1)
        free = [(e, other) for e, other in adj[v] if not picked[e]]
        # a self-loop (s == o) sits twice in adj[v] and counts twice, like in the reference's redraw loop
        e, other = free[j]
        out[i] = e
        picked[e] = True
        counts[v] -= 1
        counts[other] -= 1
        seen[other] = True
    return out


def edge_dropout_rows(num_edges, keep_prob, perm):
    """predict_links.py:145-148: graph[perm][round(keep_prob * n):] — the rows that stay, given the permutation."""
    return np.asarray(perm)[int(round(keep_prob * num_edges)):]
