"""CPU oracle for the RGCN per-relation aggregation path.  TEST INFRASTRUCTURE ONLY.

This file is the *checker*: a numpy (fp64) restatement of what
thiviyanT/torch-rgcn computes on its hot path.  Only `tests/`,
`__graft_entry__.smoke()` and `bench.py`'s `cpu_baseline` / `--impl reference`
legs may import it.  Nothing under `torch_rgcn_b200/` imports it, and the
product path raises when the CUDA library is missing instead of falling back
to this code.

Parity status: PINNED.  `tests/test_oracle_golden.py` checks every function
here against (1) the known-answer vectors of the reference's own unit tests
(reference `tests/test_utils.py:5-25, 28-84, 87-123, 170-220`) and (2) the
fixtures under `tests/golden/*.npz`, which `tests/golden/make_golden.py`
produced by importing the unmodified reference layers in the build container.

Every function cites the reference lines it restates (paths relative to the
reference checkout).  The reference's arithmetic lives in PyTorch
(`torch.sparse` COO `mm`/`spmm`, `einsum`); here the same sums are written as
explicit scatter-adds in float64.
"""
import numpy as np


def _f64(a):
    """float64 ndarray view / copy (np.float64(a) turns one-element arrays into scalars)."""
    return np.asarray(a, dtype=np.float64)


I64 = np.int64


# --------------------------------------------------------------------------
# a9: triple augmentation
# --------------------------------------------------------------------------
def generate_inverses(triples, num_rels):
    """(s,p,o) -> (o, p+R, s).  torch_rgcn/utils.py:100-107."""
    t = np.asarray(triples, dtype=I64).reshape(-1, 3)
    return np.stack([t[:, 2], t[:, 1] + num_rels, t[:, 0]], axis=1)


def self_loops(num_nodes, num_rels, keep=None):
    """(v, 2R, v) for every kept node v.  torch_rgcn/utils.py:113-122."""
    ids = np.arange(num_nodes, dtype=I64)
    if keep is not None:
        ids = ids[np.asarray(keep, dtype=bool)]
    return np.stack([ids, np.full_like(ids, 2 * num_rels), ids], axis=1)


def add_inverse_and_self(triples, num_nodes, num_rels):
    """[orig; inverse; self].  torch_rgcn/utils.py:127-141."""
    t = np.asarray(triples, dtype=I64).reshape(-1, 3)
    return np.concatenate([t, generate_inverses(t, num_rels), self_loops(num_nodes, num_rels)], axis=0)


def lp_triples_plus(triples, num_nodes, num_rels, keep=None):
    """Edge list the LP layer walks: [T; inverse(T); T; self].

    `generate_self_loops` returns cat([triples, self_loops])
    (torch_rgcn/utils.py:124) and the layer concatenates that after
    [T; inverse] (torch_rgcn/layers.py:483-487), so T appears twice.
    Returns (triples_plus, n, i) with the (n, i) the layer later uses for the
    horizontal permutation (torch_rgcn/layers.py:507-508).
    """
    t = np.asarray(triples, dtype=I64).reshape(-1, 3)
    sl = np.concatenate([t, self_loops(num_nodes, num_rels, keep)], axis=0)
    tp = np.concatenate([t, generate_inverses(t, num_rels), sl], axis=0)
    return tp, t.shape[0], sl.shape[0]


# --------------------------------------------------------------------------
# a4/a5/a6: adjacency stacking and normalisation
# --------------------------------------------------------------------------
def stack_matrices(triples, num_nodes, num_rels, vertical_stacking=True):
    """COO coordinates of the stacked adjacency.  torch_rgcn/utils.py:143-166."""
    t = np.asarray(triples, dtype=I64).reshape(-1, 3)
    n, r = num_nodes, num_rels
    size = (r * n, n) if vertical_stacking else (n, r * n)
    fr, to = t[:, 0].copy(), t[:, 2].copy()
    off = t[:, 1] * n
    if vertical_stacking:
        fr = fr + off
    else:
        to = to + off
    idx = np.stack([fr, to], axis=1)
    if idx.shape[0]:
        assert idx[:, 0].max() < size[0] and idx[:, 1].max() < size[1]
    return idx, size


def sum_sparse(indices, values, size, row_normalisation=True):
    """Row (or column) sums redistributed to every nnz.  torch_rgcn/utils.py:71-97."""
    idx = np.asarray(indices, dtype=I64).reshape(-1, 2)
    vals = np.asarray(values, dtype=np.float64)
    key = idx[:, 0] if row_normalisation else idx[:, 1]
    length = size[0] if row_normalisation else size[1]
    sums = np.bincount(key, weights=vals, minlength=int(length))
    return sums[key]


def edge_values(triples_plus, num_nodes, num_rels, vertical_stacking, n, i):
    """Per-edge weight exactly as the layers derive it (float32).

    NC: torch_rgcn/layers.py:255-273 with n=(nnz-N)//2, i=N.
    LP: torch_rgcn/layers.py:490-510 with n=|T|, i=|T|+|self|.
    Horizontal mode takes column sums and then permutes them with
    cat([sums[n:2n], sums[:n], sums[-i:]]) (layers.py:271 / :509).
    """
    idx, size = stack_matrices(triples_plus, num_nodes, num_rels, vertical_stacking)
    ones = np.ones(idx.shape[0], dtype=np.float32)
    sums = sum_sparse(idx, ones, size, row_normalisation=vertical_stacking).astype(np.float32)
    if not vertical_stacking:
        sums = np.concatenate([sums[n:2 * n], sums[:n], sums[-i:]], axis=0)
    assert sums.shape[0] == idx.shape[0], "permutation does not cover the edge list"
    return (ones / sums).astype(np.float32)


def nc_edge_values(triples_plus, num_nodes, num_rels, vertical_stacking):
    nnz = np.asarray(triples_plus).shape[0]
    return edge_values(triples_plus, num_nodes, num_rels, vertical_stacking,
                       int((nnz - num_nodes) / 2), num_nodes)   # layers.py:235-236


# --------------------------------------------------------------------------
# a3: weight decompositions -> effective dense (R', I, O)
# --------------------------------------------------------------------------
def block_diag(blocks):
    """(R, nb, bi, bo) -> (R, nb*bi, nb*bo).  torch_rgcn/utils.py:168-196."""
    b = np.asarray(blocks, dtype=np.float64)
    r, nb, bi, bo = b.shape
    out = np.zeros((r, nb * bi, nb * bo))
    for k in range(nb):
        out[:, k * bi:(k + 1) * bi, k * bo:(k + 1) * bo] = b[:, k]
    return out


def effective_weights(params):
    """Dense (R', I, O) float64 from a dict of parameter arrays.

    none:  weights                       layers.py:240 / :467
    basis: einsum('rb,bio->rio')         layers.py:242 / :469
    block: block_diag(blocks) [+ blocks_self appended for LP]   layers.py:244 / :521-522 / :540-547
    diag:  weights (R', I) read as diag  layers.py:147-151, 290-291
    """
    if 'bases' in params:
        return np.einsum('rb,bio->rio', _f64(params['comps']), _f64(params['bases']))
    if 'blocks' in params:
        w = block_diag(params['blocks'])
        if 'blocks_self' in params:
            w = np.concatenate([w, _f64(params['blocks_self'])[None]], axis=0)
        return w
    w = _f64(params['weights'])
    if w.ndim == 2:
        out = np.zeros((w.shape[0], w.shape[1], w.shape[1]))
        ar = np.arange(w.shape[1])
        out[:, ar, ar] = w
        return out
    return w


# --------------------------------------------------------------------------
# a1/a2/a7/a8: message passing, closed form
# --------------------------------------------------------------------------
def propagate(triples_plus, val, W, X=None, bias=None, num_nodes=None, self_mask=None, mask_rel=None):
    """out[s] = bias + sum_e val_e * (X[o_e] @ W[p_e])   (featureless: X = I_N).

    Closed form of layers.py:286-301 / :518-551: both stackings multiply the
    same normalised adjacency with the same per-relation transformed features.
    `self_mask` (N, O) restates the 'schlichtkrull-dropout' branch
    (layers.py:545-546): the transformed features of relation `mask_rel` are
    multiplied element-wise by the dropout mask before aggregation.
    """
    t = np.asarray(triples_plus, dtype=I64).reshape(-1, 3)
    val = _f64(val)
    W = _f64(W)
    N = num_nodes if num_nodes is not None else X.shape[0]
    O = W.shape[2]
    out = np.zeros((N, O))
    s, p, o = t[:, 0], t[:, 1], t[:, 2]
    for r in np.unique(p):
        m = p == r
        if X is None:
            msg = W[r][o[m]]
        else:
            msg = _f64(X)[o[m]] @ W[r]
        if self_mask is not None and r == mask_rel:
            msg = msg * _f64(self_mask)[o[m]]
        np.add.at(out, s[m], msg * val[m, None])
    if bias is not None:
        out = out + _f64(bias)
    return out


def propagate_backward(triples_plus, val, W, G, X=None, self_mask=None, mask_rel=None):
    """Closed-form gradients (the reference has only autograd).

    gX[o] += val * (G[s] @ W_p^T);  gW_p += val * X[o]^T G[s]
    featureless: gW[p, o, :] += val * G[s].
    Returns (gX or None, gW (R', I, O)).
    """
    t = np.asarray(triples_plus, dtype=I64).reshape(-1, 3)
    val = _f64(val)
    W = _f64(W)
    G = _f64(G)
    s, p, o = t[:, 0], t[:, 1], t[:, 2]
    gW = np.zeros_like(W)
    gX = None if X is None else np.zeros((X.shape[0], W.shape[1]))
    for r in np.unique(p):
        m = p == r
        g = G[s[m]] * val[m, None]
        if self_mask is not None and r == mask_rel:
            g = g * _f64(self_mask)[o[m]]
        if X is None:
            np.add.at(gW[r], o[m], g)
        else:
            x = _f64(X)[o[m]]
            gW[r] += x.T @ g
            np.add.at(gX, o[m], g @ W[r].T)
    return gX, gW


def project_weight_grad(gW, params):
    """Map the dense gW (R', I, O) onto the parameters of each decomposition."""
    out = {}
    if 'bases' in params:
        out['comps'] = np.einsum('rio,bio->rb', gW, _f64(params['bases']))
        out['bases'] = np.einsum('rb,rio->bio', _f64(params['comps']), gW)
    elif 'blocks' in params:
        r, nb, bi, bo = params['blocks'].shape
        gb = np.zeros((r, nb, bi, bo))
        for k in range(nb):
            gb[:, k] = gW[:r, k * bi:(k + 1) * bi, k * bo:(k + 1) * bo]
        out['blocks'] = gb
        if 'blocks_self' in params:
            out['blocks_self'] = gW[r]
    elif np.asarray(params['weights']).ndim == 2:
        ar = np.arange(gW.shape[1])
        out['weights'] = gW[:, ar, ar]
    else:
        out['weights'] = gW
    return out


# --------------------------------------------------------------------------
# whole-layer restatements
# --------------------------------------------------------------------------
def nc_layer(triples_plus, num_nodes, num_rels, params, X=None, vertical_stacking=False, G=None):
    """RelationalGraphConvolutionNC.forward (layers.py:222-308) + autograd, closed form.

    Returns out, or (out, grads dict incl. 'features') when G is given.
    """
    val = nc_edge_values(triples_plus, num_nodes, num_rels, vertical_stacking)
    W = effective_weights(params)
    out = propagate(triples_plus, val, W, X, params.get('bias'), num_nodes)
    if G is None:
        return out
    gX, gW = propagate_backward(triples_plus, val, W, G, X)
    grads = project_weight_grad(gW, params)
    if params.get('bias') is not None:
        grads['bias'] = _f64(G).sum(0)
    grads['features'] = gX
    return out, grads


def lp_layer(triples, num_nodes, num_rels_total, params, X, vertical_stacking=False, G=None,
             keep=None, self_mask=None):
    """RelationalGraphConvolutionLP.forward (layers.py:450-565) + autograd, closed form.

    `num_rels_total` is the layer's num_relations (2R+1).  `keep` is the
    bernoulli self-loop keep mask (utils.py:120-122), `self_mask` the
    (N, O) dropout mask of the 'schlichtkrull-dropout' block branch.
    """
    R = int((num_rels_total - 1) / 2)          # layers.py:460
    tp, n, i = lp_triples_plus(triples, num_nodes, R, keep)
    val = edge_values(tp, num_nodes, num_rels_total, vertical_stacking, n, i)
    W = effective_weights(params)
    out = propagate(tp, val, W, X, params.get('bias'), num_nodes, self_mask, num_rels_total - 1)
    if G is None:
        return out
    gX, gW = propagate_backward(tp, val, W, G, X, self_mask, num_rels_total - 1)
    grads = project_weight_grad(gW, params)
    if params.get('bias') is not None:
        grads['bias'] = _f64(G).sum(0)
    grads['features'] = gX
    return out, grads
