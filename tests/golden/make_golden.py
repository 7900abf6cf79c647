"""Generate tests/golden/*.npz by running the UNMODIFIED reference layers.

Run in the build container only (the reference checkout does not travel to the GPU box):

    python tests/golden/make_golden.py [/root/reference]

Each fixture stores the seeded inputs (triples, features, every parameter, the
upstream gradient) and what the reference produced (output, autograd gradients of
every parameter and of the features).  The fixtures are the committed evidence that
`oracle/rgcn_oracle.py` and the CUDA path reproduce the reference; nothing here is
imported by the product.
"""
import json
import os
import sys

import numpy as np
import torch

REF = sys.argv[1] if len(sys.argv) > 1 else '/root/reference'
sys.path.insert(0, REF)
from torch_rgcn.layers import RelationalGraphConvolutionNC, RelationalGraphConvolutionLP, DistMult  # noqa: E402
from torch_rgcn.utils import add_inverse_and_self  # noqa: E402
from torch_rgcn.models import NodeClassifier, EmbeddingNodeClassifier, CompressionRelationPredictor  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def rand_triples(gen, n_nodes, n_rels, n_edges, dup=0.25):
    s = torch.randint(0, n_nodes, (n_edges,), generator=gen)
    p = torch.randint(0, n_rels, (n_edges,), generator=gen)
    o = torch.randint(0, n_nodes, (n_edges,), generator=gen)
    t = torch.stack([s, p, o], dim=1)
    k = int(n_edges * dup)                       # repeat some rows / share (s,p) so segments have >1 edge
    if k:
        t[-k:, 0] = t[:k, 0]
        t[-k:, 1] = t[:k, 1]
        h = max(k // 2, 1)
        t[-h:, 2] = t[:h, 2]                      # and some exact duplicates
    return t.long()


def grads_of(layer, feats):
    g = {n: p.grad.detach().numpy().copy() for n, p in layer.named_parameters() if p.grad is not None}
    if feats is not None and feats.grad is not None:
        g['features'] = feats.grad.detach().numpy().copy()
    return g


def save(name, meta, arrays):
    arrays = {k: np.asarray(v) for k, v in arrays.items()}
    np.savez_compressed(os.path.join(HERE, name + '.npz'), meta=np.array(json.dumps(meta)), **arrays)
    print('wrote', name, {k: v.shape for k, v in arrays.items() if k in ('out', 'triples')})


def make_nc(name, seed, N, R, E, in_f, out_f, decomposition=None, vertical=False, diag=False, bias=True,
            shuffle=False):
    gen = torch.Generator().manual_seed(seed)
    triples = rand_triples(gen, N, R, E)
    tp = add_inverse_and_self(triples, N, R)
    if shuffle:                                   # non-canonical row order: horizontal normalisation depends on it
        tp = tp[torch.randperm(tp.size(0), generator=gen)]
    torch.manual_seed(seed + 2)
    layer = RelationalGraphConvolutionNC(triples=tp, num_nodes=N, num_relations=2 * R + 1, in_features=in_f,
                                         out_features=out_f, bias=bias, decomposition=decomposition,
                                         vertical_stacking=vertical, diag_weight_matrix=diag)
    if layer.bias is not None:                    # zero bias would not exercise the add
        with torch.no_grad():
            layer.bias.normal_(0, 1, generator=gen)
    feats = None
    if in_f is not None:
        feats = torch.randn(N, in_f, generator=gen).requires_grad_(True)
    out = layer(feats) if feats is not None else layer()
    G = torch.randn(out.shape, generator=gen)
    out.backward(G)
    arrays = {'triples_plus': tp.numpy(), 'out': out.detach().numpy(), 'G': G.numpy()}
    if feats is not None:
        arrays['features'] = feats.detach().numpy()
    for n, p in layer.named_parameters():
        arrays['param_' + n] = p.detach().numpy()
    for n, g in grads_of(layer, feats).items():
        arrays['grad_' + n] = g
    meta = dict(kind='nc', N=N, R=R, num_relations=2 * R + 1, in_features=in_f, out_features=layer.out_features,
                decomposition=decomposition, vertical=vertical, diag=diag, bias=bias)
    save(name, meta, arrays)


def make_lp(name, seed, N, R, E, in_f, out_f, decomposition=None, vertical=False, b_init=None, train=None):
    gen = torch.Generator().manual_seed(seed)
    triples = rand_triples(gen, N, R, E)
    torch.manual_seed(seed + 2)
    edo = None
    if train == 'bernoulli':
        edo = {'general': 0.5, 'self_loop': 0.4, 'self_loop_type': 'plain'}
    elif train == 'schlichtkrull':
        edo = {'general': 0.5, 'self_loop': 0.2, 'self_loop_type': 'schlichtkrull-dropout'}
    layer = RelationalGraphConvolutionLP(num_nodes=N, num_relations=2 * R + 1, in_features=in_f, out_features=out_f,
                                         edge_dropout=edo, decomposition=decomposition, vertical_stacking=vertical,
                                         w_init='glorot-normal', b_init=b_init)
    if layer.bias is not None:
        with torch.no_grad():
            layer.bias.normal_(0, 1, generator=gen)
    feats = torch.randn(N, in_f, generator=gen).requires_grad_(True)
    arrays = {}
    layer.train(train is not None)
    rng_seed = seed + 7
    if train == 'bernoulli':
        # first RNG draw inside forward is the self-loop keep mask (reference utils.py:120-121)
        torch.manual_seed(rng_seed)
        arrays['keep'] = torch.bernoulli(torch.empty(N).fill_(1 - edo['self_loop'])).bool().numpy()
    elif train == 'schlichtkrull':
        # forward draws bernoulli(N) with p=1 first, then F.dropout on a (1,N,O) tensor (reference layers.py:545-546)
        torch.manual_seed(rng_seed)
        torch.bernoulli(torch.empty(N).fill_(1.0))
        arrays['self_mask'] = torch.nn.functional.dropout(torch.ones(1, N, out_f), p=edo['self_loop'],
                                                          training=True)[0].numpy()
    torch.manual_seed(rng_seed)
    out = layer(triples, feats)
    G = torch.randn(out.shape, generator=gen)
    out.backward(G)
    arrays.update({'triples': triples.numpy(), 'out': out.detach().numpy(), 'G': G.numpy(),
                   'features': feats.detach().numpy()})
    for n, p in layer.named_parameters():
        arrays['param_' + n] = p.detach().numpy()
    for n, g in grads_of(layer, feats).items():
        arrays['grad_' + n] = g
    meta = dict(kind='lp', N=N, R=R, num_relations=2 * R + 1, in_features=in_f, out_features=out_f,
                decomposition=decomposition, vertical=vertical, b_init=b_init, train=train, edge_dropout=edo)
    save(name, meta, arrays)


def make_model(name, seed, cls, N, R, E, nclass, **kw):
    """Whole reference model (models.py:137-296): logits and autograd gradients of every parameter."""
    gen = torch.Generator().manual_seed(seed)
    triples = rand_triples(gen, N, R, E)
    torch.manual_seed(seed + 2)
    model = cls(triples=triples.tolist(), nnodes=N, nrel=R, nclass=nclass, **kw)
    with torch.no_grad():
        for n, p in model.named_parameters():
            if n.endswith('bias'):
                p.normal_(0, 1, generator=gen)
    out = model()
    G = torch.randn(out.shape, generator=gen)
    out.backward(G)
    arrays = {'triples': triples.numpy(), 'out': out.detach().numpy(), 'G': G.numpy()}
    for n, p in model.named_parameters():
        arrays['param_' + n] = p.detach().numpy()
        arrays['grad_' + n] = p.grad.detach().numpy()
    meta = dict(kind='model', cls=cls.__name__, N=N, R=R, nclass=nclass, kwargs=kw,
                state_keys=sorted(model.state_dict().keys()))
    save(name, meta, arrays)


def make_distmult(name, seed, N, R, d, B, b_init=None, three_d=False):
    """Reference DistMult (layers.py:9-98): scores, s_penalty and the autograd gradients of both."""
    gen = torch.Generator().manual_seed(seed)
    shape = (B // 4, 4) if three_d else (B,)
    triples = torch.stack([torch.randint(0, N, shape, generator=gen), torch.randint(0, R, shape, generator=gen),
                           torch.randint(0, N, shape, generator=gen)], dim=-1)
    flat = triples.view(-1, 3)
    flat[-B // 4:, 0] = flat[:B // 4, 0]                   # repeated subjects / relations: gradients accumulate
    flat[-B // 8:, 1] = flat[:B // 8, 1]
    torch.manual_seed(seed + 2)
    dec = DistMult(R, d, N, R, w_init='standard-normal', b_init=b_init)
    nodes = torch.randn(N, d, generator=gen, requires_grad=True)
    scores = dec(triples, nodes)
    G = torch.randn(scores.shape, generator=gen)
    scores.backward(G)
    arrays = {'triples': triples.numpy(), 'nodes': nodes.detach().numpy(), 'G': G.numpy(), 'out': scores.detach().numpy(),
              'grad_nodes': nodes.grad.numpy().copy()}
    for n, p in dec.named_parameters():
        arrays['param_' + n] = p.detach().numpy().copy()
        arrays['grad_' + n] = p.grad.numpy().copy()
    nodes.grad = None
    dec.zero_grad()
    pen = dec.s_penalty(triples, nodes)
    (pen * 3.0).backward()
    arrays['penalty'] = pen.detach().numpy()
    arrays['pgrad_nodes'] = nodes.grad.numpy().copy()
    arrays['pgrad_relations'] = dec.relations.grad.numpy().copy()
    save(name, dict(kind='distmult', seed=seed, N=N, R=R, d=d, B=B, b_init=b_init, penalty_grad=3.0), arrays)


def make_negsample(name, seed, N, bs, ns, head_prob):
    """Reference negative_sampling (utils/misc.py:174-189).  utils/misc.py imports sacred at module level (absent
    here), so the import runs with a stub module in its place; the function itself is executed unmodified.  The two
    random draws are re-drawn from the same seed to record the corruptions and the head mask it used."""
    import types
    for m in ('sacred', 'sacred.observers'):
        sys.modules.setdefault(m, types.ModuleType(m))
    sys.modules['sacred'].Experiment = object
    sys.modules['sacred.observers'].MongoObserver = object
    from utils.misc import negative_sampling
    gen = torch.Generator().manual_seed(seed)
    pos = torch.stack([torch.randint(0, N, (bs,), generator=gen), torch.randint(0, 5, (bs,), generator=gen),
                       torch.randint(0, N, (bs,), generator=gen)], dim=1)
    batch = pos.clone()[:, None, :].expand(bs, ns, 3).contiguous()       # experiments/predict_links.py:132
    before = batch.numpy().copy()
    torch.manual_seed(seed + 1)
    out = negative_sampling(batch, N, head_prob)
    torch.manual_seed(seed + 1)
    corruptions = torch.randint(size=(bs * ns,), low=0, high=N, dtype=torch.long)
    head = torch.bernoulli(torch.empty(size=(bs, ns, 1), dtype=torch.float).fill_(head_prob)).to(torch.bool)
    save(name, dict(kind='negsample', seed=seed, N=N, bs=bs, ns=ns, head_prob=head_prob),
         {'batch': before, 'corruptions': corruptions.numpy(), 'head': head.numpy().reshape(-1), 'out': out.numpy()})


def make_crp(name, seed, N, R, E, nemb, B):
    """Reference CompressionRelationPredictor (models.py:208-245) where it runs as shipped: node_embedding ==
    hidden1_size (the encoder layer is built for the embedding width, models.py:224) and a glorot initialiser.  eval()
    mode: no self-loop dropout draws."""
    gen = torch.Generator().manual_seed(seed)
    graph = rand_triples(gen, N, R, E)
    batch = torch.stack([torch.randint(0, N, (B,), generator=gen), torch.randint(0, R, (B,), generator=gen),
                         torch.randint(0, N, (B,), generator=gen)], 1)
    enc = {'node_embedding': nemb, 'hidden1_size': nemb, 'num_layers': 1, 'weight_init': 'glorot-normal',
           'bias_init': 'zeros', 'edge_dropout': {'general': 0.5, 'self_loop': 0.2, 'self_loop_type': 'schlichtkrull-dropout'}}
    dec = {'l2_penalty_type': 'schlichtkrull-l2', 'l2_penalty': 0.01, 'weight_init': 'standard-normal'}
    torch.manual_seed(seed + 2)
    model = CompressionRelationPredictor(nnodes=N, nrel=R, encoder_config=enc, decoder_config=dec).eval()
    scores, penalty = model(graph, batch)
    G = torch.randn(scores.shape, generator=gen)
    ((scores * G).sum() + 0.5 * penalty).backward()
    arrays = {'graph': graph.numpy(), 'batch': batch.numpy(), 'G': G.numpy(), 'out': scores.detach().numpy(),
              'penalty': penalty.detach().numpy()}
    for n, p in model.named_parameters():
        arrays['param_' + n] = p.detach().numpy().copy()
        arrays['grad_' + n] = p.grad.numpy().copy()
    save(name, dict(kind='crp', seed=seed, N=N, R=R, nemb=nemb, encoder=enc, decoder=dec, penalty_weight=0.5), arrays)


def make_ranking(name, seed, N, R, d, n_known, n_test, integer=False, bias=False, batch_size=7):
    """Reference evaluate (utils/misc.py:60-110, with filter_scores :39-58 and generate_true_dict :29-37) on fixed node
    embeddings.  `integer` draws small-integer embeddings: every score is then exact in fp32 whatever the summation
    order, ties are frequent, and the ranks must match an implementation exactly."""
    _stub_sacred()
    from utils.misc import evaluate, generate_true_dict
    gen = torch.Generator().manual_seed(seed)
    known = torch.stack([torch.randint(0, N, (n_known,), generator=gen), torch.randint(0, R, (n_known,), generator=gen),
                         torch.randint(0, N, (n_known,), generator=gen)], 1)
    known[: n_known // 3, 1] = known[0, 1]                   # many completions of a few (p, o) / (s, p) pairs
    known[: n_known // 3, 2] = known[0, 2]
    known[n_known // 3: 2 * n_known // 3, 0] = known[1, 0]
    known[n_known // 3: 2 * n_known // 3, 1] = known[1, 1]
    known = torch.cat([known, known[:5]], 0)                # repeated known triples
    test = known[torch.randperm(known.size(0), generator=gen)[:n_test]]
    torch.manual_seed(seed + 2)
    dec = DistMult(R, d, N, R, b_init='normal' if bias else None)
    if integer:
        x = torch.randint(-2, 3, (N, d), generator=gen).float()
        with torch.no_grad():
            dec.relations.copy_(torch.randint(-2, 3, (R, d), generator=gen).float())
            if bias:
                for b in (dec.sbias, dec.pbias, dec.obias):
                    b.copy_(torch.randint(-3, 4, b.shape, generator=gen).float())
    else:
        x = torch.randn(N, d, generator=gen)
    model = lambda graph, triples: (dec(triples, x), None)        # noqa: E731  what LinkPredictor.forward returns
    true = generate_true_dict(known.tolist())
    arrays = {'known': known.numpy(), 'test': test.numpy(), 'nodes': x.numpy()}
    for n, p in dec.named_parameters():
        arrays['param_' + n] = p.detach().numpy().copy()
    for filt in (True, False):
        with torch.no_grad():
            mrr, hits, ranks = evaluate(model, None, test, true, N, batch_size=batch_size, filter_candidates=filt,
                                        verbose=False)
        tag = 'filtered' if filt else 'raw'
        arrays['ranks_' + tag] = np.array(ranks)
        arrays['mrr_' + tag] = np.array(mrr)
        arrays['hits_' + tag] = np.array(hits)
    save(name, dict(kind='ranking', seed=seed, N=N, R=R, d=d, integer=integer, b_init='normal' if bias else None), arrays)


def _stub_sacred():
    import types
    for m in ('sacred', 'sacred.observers'):
        sys.modules.setdefault(m, types.ModuleType(m))
    sys.modules['sacred'].Experiment = object
    sys.modules['sacred.observers'].MongoObserver = object


def make_sampling_hist(name, seed, runs):
    """Reference edge_neighborhood (utils/misc.py:125-172) run `runs` times on a small graph with a self-loop, a
    repeated edge and an isolated node: histogram of the ORDERED samples it returns.  The sampler draws from numpy's
    global generator, so parity of a restatement is statistical."""
    _stub_sacred()
    from utils.misc import edge_neighborhood
    triples = [[0, 0, 1], [1, 1, 2], [2, 0, 0], [3, 1, 3], [0, 1, 1], [4, 0, 2], [1, 0, 4], [0, 0, 1], [2, 1, 3]]
    N, S = 6, 3                                   # node 5 has no edges
    entities = {'n%d' % i: i for i in range(N)}
    np.random.seed(seed)
    E = len(triples)
    hist = np.zeros((E,) * S, dtype=np.int64)
    for _ in range(runs):
        picked = edge_neighborhood(triples, sample_size=S, entities=entities)
        idx = []
        for t in picked:                          # the function returns triples; map back to edge indices (first free match)
            cands = [i for i, tt in enumerate(triples) if tt == t and i not in idx]
            idx.append(cands[0])
        hist[tuple(idx)] += 1
    save(name, dict(kind='sampling_hist', seed=seed, N=N, S=S, runs=runs), {'triples': np.array(triples), 'hist': hist})


def main():
    make_sampling_hist('sampling_hist', 71, 40000)
    make_crp('lpmodel_crp', 61, 30, 3, 90, 16, 64)
    make_ranking('ranking_integer', 51, 60, 4, 8, 150, 40, integer=True)
    make_ranking('ranking_integer_bias', 52, 45, 3, 6, 120, 30, integer=True, bias=True)
    # the reference's filter_scores raises when a batch has nothing to filter (utils/misc.py:56-58 indexes an empty 1-D
    # tensor with [:, 0]); one batch over the whole test set avoids that here
    make_ranking('ranking_float', 53, 80, 5, 16, 200, 50, batch_size=50)
    make_distmult('distmult_plain', 31, 40, 5, 16, 64)
    make_distmult('distmult_bias', 32, 40, 5, 16, 64, b_init='normal')
    make_distmult('distmult_3d_odd', 33, 30, 4, 10, 48, b_init='uniform', three_d=True)
    make_distmult('distmult_wide', 34, 50, 6, 128, 96)
    make_negsample('negsample_half', 41, 50, 12, 10, 0.5)
    make_negsample('negsample_heads', 42, 50, 7, 3, 1.0)
    N, R, E = 24, 3, 70
    basis = {'type': 'basis', 'num_bases': 3}
    block = {'type': 'block', 'num_blocks': 2}
    # --- node-classification layer (reference layers.py:101-308)
    make_nc('nc_none_h_feat', 0, N, R, E, 12, 8)
    make_nc('nc_none_v_feat', 1, N, R, E, 12, 8, vertical=True)
    make_nc('nc_none_h_featureless', 2, N, R, E, None, 8)
    make_nc('nc_basis_h_feat', 3, N, R, E, 12, 8, decomposition=basis)
    make_nc('nc_basis_v_feat', 4, N, R, E, 12, 8, decomposition=basis, vertical=True)
    make_nc('nc_basis_h_featureless', 5, N, R, E, None, 8, decomposition=basis)
    make_nc('nc_block_h_feat', 6, N, R, E, 12, 8, decomposition=block)
    make_nc('nc_block_v_feat', 7, N, R, E, 12, 8, decomposition=block, vertical=True)
    make_nc('nc_block_h_featureless', 8, N, R, E, None, 8, decomposition=block)
    make_nc('nc_diag_h', 9, N, R, E, 12, 12, diag=True)
    make_nc('nc_none_h_feat_odd', 10, 23, R, E, 10, 3)                      # dims not multiples of 4
    make_nc('nc_basis_h_featureless_odd', 11, 23, R, E, None, 10, decomposition=basis)
    make_nc('nc_block_v_feat_odd', 12, 23, R, E, 10, 6, decomposition=block, vertical=True)   # 5x3 blocks
    make_nc('nc_none_h_feat_nobias', 13, N, R, E, 12, 8, bias=False)
    make_nc('nc_none_h_feat_shuffled', 14, N, R, E, 12, 8, shuffle=True)
    make_nc('nc_none_v_feat_shuffled', 15, N, R, E, 12, 8, vertical=True, shuffle=True)
    make_nc('nc_none_h_feat_wide', 16, 40, 5, 200, 64, 32)
    make_nc('nc_block_h_feat_wide', 17, 40, 5, 200, 64, 64, decomposition={'type': 'block', 'num_blocks': 4})
    # --- link-prediction layer (reference layers.py:311-565)
    make_lp('lp_none_h', 20, N, R, E, 12, 8, b_init='zeros')
    make_lp('lp_none_v', 21, N, R, E, 12, 8)
    make_lp('lp_basis_h', 22, N, R, E, 12, 8, decomposition=basis, b_init='normal')
    make_lp('lp_basis_v', 23, N, R, E, 12, 8, decomposition=basis)
    make_lp('lp_block_h', 24, N, R, E, 12, 8, decomposition=block, b_init='zeros')
    make_lp('lp_none_h_train_bernoulli', 25, N, R, E, 12, 8, b_init='zeros', train='bernoulli')
    make_lp('lp_block_h_train_schlichtkrull', 26, N, R, E, 12, 8, decomposition=block, b_init='zeros',
            train='schlichtkrull')
    make_lp('lp_none_h_wn18like', 27, 60, 9, 300, 16, 16, b_init='zeros')
    # --- whole node-classification models (reference models.py:137-296)
    make_model('model_nc_none', 30, NodeClassifier, 40, 4, 160, 3, nhid=16, nlayers=2)
    make_model('model_nc_basis', 31, NodeClassifier, 40, 4, 160, 3, nhid=16, nlayers=2,
               decomposition={'type': 'basis', 'num_bases': 5})
    make_model('model_nc_block', 32, NodeClassifier, 40, 4, 160, 4, nhid=16, nlayers=2,
               decomposition={'type': 'block', 'num_blocks': 2})
    make_model('model_nc_1layer', 33, NodeClassifier, 40, 4, 160, 3, nlayers=1)
    make_model('model_ernn', 34, EmbeddingNodeClassifier, 40, 4, 160, 3, nemb=32, nlayers=2)


if __name__ == '__main__':
    main()
