"""CPU restatement (numpy, fp64 accumulation) of the reference's DistMult decoder and negative sampling.

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg, never by the
product.  Pinned against fixtures generated from the unmodified reference (tests/golden/distmult_*.npz,
tests/golden/negsample_*.npz; generator tests/golden/make_golden.py).

Reference: torch_rgcn/layers.py:77-98 (DistMult.s_penalty / forward), torch_rgcn/utils.py:201-206 (split_spo),
utils/misc.py:174-189 (negative_sampling).
"""
import numpy as np


def _spo(triples):
    t = np.asarray(triples).reshape(-1, 3)                 # split_spo handles (B, 3) and (B, K, 3) alike
    return t[:, 0], t[:, 1], t[:, 2], np.asarray(triples).shape[:-1]


def score(triples, nodes, relations, sbias=None, pbias=None, obias=None):
    """layers.py:86-98: (s * p * o).sum(-1) (+ sbias[s] + pbias[p] + obias[o])"""
    s, p, o, lead = _spo(triples)
    nodes, relations = np.asarray(nodes, np.float64), np.asarray(relations, np.float64)
    out = (nodes[s] * relations[p] * nodes[o]).sum(-1)
    if sbias is not None:
        out = out + (np.asarray(sbias, np.float64)[s] + np.asarray(pbias, np.float64)[p] + np.asarray(obias, np.float64)[o])
    return out.reshape(lead)


def score_backward(triples, nodes, relations, grad, with_bias=False):
    """Closed-form gradients of `score` (autograd upstream)."""
    s, p, o, _ = _spo(triples)
    nodes, relations = np.asarray(nodes, np.float64), np.asarray(relations, np.float64)
    g = np.asarray(grad, np.float64).reshape(-1, 1)
    g_nodes = np.zeros_like(nodes)
    g_rel = np.zeros_like(relations)
    np.add.at(g_nodes, s, g * relations[p] * nodes[o])
    np.add.at(g_nodes, o, g * nodes[s] * relations[p])
    np.add.at(g_rel, p, g * nodes[s] * nodes[o])
    out = {'nodes': g_nodes, 'relations': g_rel}
    if with_bias:
        for name, idx, n in (('sbias', s, nodes.shape[0]), ('pbias', p, relations.shape[0]), ('obias', o, nodes.shape[0])):
            b = np.zeros(n)
            np.add.at(b, idx, g[:, 0])
            out[name] = b
    return out


def penalty(triples, nodes, relations):
    """layers.py:77-84: s.pow(2).mean() + p.pow(2).mean() + o.pow(2).mean()"""
    s, p, o, _ = _spo(triples)
    nodes, relations = np.asarray(nodes, np.float64), np.asarray(relations, np.float64)
    return (nodes[s] ** 2).mean() + (relations[p] ** 2).mean() + (nodes[o] ** 2).mean()


def penalty_backward(triples, nodes, relations, grad=1.0):
    s, p, o, _ = _spo(triples)
    nodes, relations = np.asarray(nodes, np.float64), np.asarray(relations, np.float64)
    scale = 2.0 * float(grad) / (len(s) * nodes.shape[1])
    g_nodes = np.zeros_like(nodes)
    g_rel = np.zeros_like(relations)
    np.add.at(g_nodes, s, scale * nodes[s])
    np.add.at(g_nodes, o, scale * nodes[o])
    np.add.at(g_rel, p, scale * relations[p])
    return {'nodes': g_nodes, 'relations': g_rel}


def corrupt(batch, head_mask, corruptions):
    """utils/misc.py:181-187: mask = cat([head, 0, ~head], dim=2); batch[mask] = corruptions (row-major order)."""
    b = np.array(batch).reshape(-1, 3).copy()
    head = np.asarray(head_mask).reshape(-1).astype(bool)
    c = np.asarray(corruptions).reshape(-1)
    b[head, 0] = c[head]
    b[~head, 2] = c[~head]
    return b
