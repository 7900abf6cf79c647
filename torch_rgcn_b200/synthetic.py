"""Seeded synthetic relational graphs with the shapes of BASELINE.json's configs.

No datasets are available offline, so every benchmark and large parity test runs on these.
"""
import torch

# name: (num_nodes, raw relations R, raw triples E)   — standard dataset sizes, used only to size the graphs
SHAPES = {
    'aifb': (8285, 45, 29043),
    'mutag': (23644, 23, 74227),
    'am': (1666764, 133, 5988321),
    'wn18': (40943, 18, 141442),
    'fb15k237': (14541, 237, 272115),
    'syn': (5000000, 128, 100000000),
}


def random_triples(num_nodes, num_rels, num_triples, seed=0, device='cpu', rel_dist='uniform', node_skew=False):
    """(E, 3) int64 triples.  rel_dist: 'uniform' | 'zipf';  node_skew: cubic skew through a fixed permutation."""
    g = torch.Generator(device=device).manual_seed(seed)
    kw = dict(generator=g, device=device)
    if node_skew:
        perm = torch.randperm(num_nodes, **kw)
        s = perm[(num_nodes * torch.rand(num_triples, **kw) ** 3).long().clamp_(max=num_nodes - 1)]
        o = perm[(num_nodes * torch.rand(num_triples, **kw) ** 3).long().clamp_(max=num_nodes - 1)]
    else:
        s = torch.randint(0, num_nodes, (num_triples,), **kw)
        o = torch.randint(0, num_nodes, (num_triples,), **kw)
    if rel_dist == 'zipf':
        w = 1.0 / torch.arange(1, num_rels + 1, dtype=torch.float, device=device)
        p = torch.multinomial(w, num_triples, replacement=True, generator=g)
    else:
        p = torch.randint(0, num_rels, (num_triples,), **kw)
    return torch.stack([s, p, o], dim=1)
