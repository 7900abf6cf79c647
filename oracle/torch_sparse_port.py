"""CPU port of the reference's *algorithm* (not its closed form) — the timed CPU baseline.  TEST/BENCH INFRASTRUCTURE ONLY.

`rgcn_oracle.py` restates WHAT the reference computes in O(nnz) numpy; this file restates HOW it computes
it, op for op on torch CPU tensors, so that `bench.py`'s `cpu_baseline` / `--impl reference` time the same
work the reference does: stacked COO adjacency, `sum_sparse` through a sparse x ones product, the dense
(R', N, d) temporaries of the horizontal / vertical branches, and plain autograd for the backward.
The reference checkout itself is pure Python and does not travel to the GPU box, hence a port
(`cpu_baseline.kind = "port"`).  The ops are device-agnostic, so the same port on CUDA tensors gives the reference's
GPU path (`bench.py --impl reference-gpu`, informational).  Pinned by tests/test_oracle_golden.py::test_port_* against the golden
fixtures (reference outputs and autograd gradients).

Only bench.py and tests/ import this file.  References: torch_rgcn/utils.py:71-97, :143-196;
torch_rgcn/layers.py:222-308 (NC), :450-565 (LP).
"""
import torch


def _coo(indices, values, size):
    # the reference uses the legacy torch.sparse.FloatTensor ctor (layers.py:279); same uncoalesced COO tensor
    return torch.sparse_coo_tensor(indices.t(), values, size=size, check_invariants=False)


def stack(triples, n, r, vertical):                       # utils.py:143-166
    fr, to = triples[:, 0], triples[:, 2]
    off = triples[:, 1] * n
    if vertical:
        fr = off + fr
    else:
        to = off + to
    size = (r * n, n) if vertical else (n, r * n)
    idx = torch.cat([fr[:, None], to[:, None]], dim=1)
    assert idx[:, 0].max() < size[0] and idx[:, 1].max() < size[1]
    return idx, size


def sums_per_entry(indices, values, size, row_normalisation):   # utils.py:71-97
    if not row_normalisation:
        indices = torch.cat([indices[:, 1:2], indices[:, 0:1]], dim=1)
        size = (size[1], size[0])
    ones = torch.ones((size[1], 1), device=indices.device)
    sums = torch.sparse.mm(_coo(indices, values, size), ones)
    return sums[indices[:, 0], 0]


def block_diag(m):                                        # utils.py:168-196
    r, nb, bi, bo = m.shape
    eye = torch.eye(nb, device=m.device).view(1, nb, 1, nb, 1)
    return (m.unsqueeze(-2) * eye).reshape(r, nb * bi, nb * bo)


def adjacency(triples_plus, n_nodes, n_rels, vertical, n, i):   # layers.py:255-279 / :490-516
    idx, size = stack(triples_plus, n_nodes, n_rels, vertical)
    vals = torch.ones(idx.size(0), device=idx.device)
    sums = sums_per_entry(idx, vals, size, vertical)
    if not vertical:
        sums = torch.cat([sums[n:2 * n], sums[:n], sums[-i:]], dim=0)
    return _coo(idx, vals / sums, size)


def nc_forward(triples_plus, n_nodes, n_rels, params, features=None, vertical=False):
    """RelationalGraphConvolutionNC.forward, op for op (layers.py:222-308).  params: dict of tensors."""
    n = int((triples_plus.size(0) - n_nodes) / 2)
    if 'bases' in params:
        weights = torch.einsum('rb, bio -> rio', params['comps'], params['bases'])
    elif 'blocks' in params:
        weights = block_diag(params['blocks'])
    else:
        weights = params['weights']
    adj = adjacency(triples_plus, n_nodes, n_rels, vertical, n, n_nodes)
    if features is None:
        out = torch.mm(adj, weights.view(n_rels * n_nodes, -1))
    elif weights.dim() == 2:                              # diag (layers.py:289-292)
        fw = torch.einsum('ij,kj->kij', features, weights).reshape(n_rels * n_nodes, -1)
        out = torch.mm(adj, fw)
    elif vertical:
        af = torch.sparse.mm(adj, features).view(n_rels, n_nodes, -1)
        out = torch.einsum('rio, rni -> no', weights, af)
    else:
        fw = torch.einsum('ni, rio -> rno', features, weights).contiguous()
        out = torch.mm(adj, fw.view(n_rels * n_nodes, -1))
    if params.get('bias') is not None:
        out = out + params['bias']
    return out


def lp_forward(triples, n_nodes, n_rels, params, features, vertical=False):
    """RelationalGraphConvolutionLP.forward in eval mode, op for op (layers.py:450-565)."""
    r = int((n_rels - 1) / 2)
    inv = torch.cat([triples[:, 2, None], triples[:, 1, None] + r, triples[:, 0, None]], dim=1)
    ids = torch.arange(n_nodes, device=triples.device)[:, None]
    loops = torch.cat([ids, torch.full_like(ids, 2 * r), ids], dim=1)
    self_part = torch.cat([triples, loops], dim=0)        # utils.py:124
    tp = torch.cat([triples, inv, self_part], dim=0)
    adj = adjacency(tp, n_nodes, n_rels, vertical, triples.size(0), self_part.size(0))
    if 'blocks' in params and not vertical:               # layers.py:534-548
        nb = params['blocks'].size(1)
        bf = features.view(n_nodes, nb, -1)
        fw = torch.einsum('nbi, rbio -> rnbo', bf, params['blocks']).contiguous().view(n_rels - 1, n_nodes, -1)
        self_fw = torch.einsum('ni, io -> no', features, params['blocks_self'])[None]
        out = torch.mm(adj, torch.cat([fw, self_fw], dim=0).view(n_rels * n_nodes, -1))
    else:
        weights = torch.einsum('rb, bio -> rio', params['comps'], params['bases']) if 'bases' in params \
            else params['weights']
        if vertical:
            af = torch.sparse.mm(adj, features).view(n_rels, n_nodes, -1)
            out = torch.einsum('rio, rni -> no', weights, af)
        else:
            fw = torch.einsum('ni, rio -> rno', features, weights).contiguous()
            out = torch.mm(adj, fw.view(n_rels * n_nodes, -1))
    if params.get('bias') is not None:
        out = out + params['bias']
    return out
