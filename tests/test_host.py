"""CPU-side checks: the C-ABI library loads and exports the header's symbols, host logic, loud failure."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

from conftest import ROOT, load_golden, golden_names


def test_library_exports_every_declared_symbol():
    from torch_rgcn_b200 import _lib
    header = open(os.path.join(ROOT, 'include', 'rgcn_b200.h')).read()
    declared = sorted(set(re.findall(r'\b(rgcn_[a-z_]+)\s*\(', header)))
    assert len(declared) >= 16
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for name in declared:
        assert hasattr(lib, name), f'{name} declared in include/rgcn_b200.h but not exported'
    assert sorted(_lib.EXPORTS) == declared
    assert _lib.lib.rgcn_abi_version() == int(re.search(r'RGCN_ABI_VERSION (\d+)', header).group(1))


def test_struct_layouts_match_header():
    """ctypes mirrors of rgcn_graph / rgcn_params / rgcn_grads have one field per header member, in order."""
    from torch_rgcn_b200 import _lib
    header = open(os.path.join(ROOT, 'include', 'rgcn_b200.h')).read()
    for cname, cls in (('rgcn_graph', _lib.Graph), ('rgcn_params', _lib.Params), ('rgcn_grads', _lib.Grads),
                       ('rgcn_tiling', _lib.Tiling), ('rgcn_fused', _lib.Fused)):
        body = re.search(r'typedef struct %s \{(.*?)\} %s;' % (cname, cname), header, re.S).group(1)
        body = re.sub(r'/\*.*?\*/', '', body, flags=re.S)
        members = re.findall(r'(\w+)\s*(?:\[\d+\])?\s*;', body)
        assert members == [f[0] for f in cls._fields_], cname


def test_shard_plan_balances_and_is_deterministic():
    from torch_rgcn_b200 import _lib
    rng = np.random.RandomState(0)
    counts = (rng.zipf(1.5, 267) % 100000).astype(np.int64)
    counts[-1] = 1_666_764                      # the self-loop relation
    for world in (1, 2, 4, 8):
        owner = np.empty(267, np.int32)
        _lib.check(_lib.lib.rgcn_shard_plan(counts.ctypes.data_as(ctypes.c_void_p), 267, world,
                                            owner.ctypes.data_as(ctypes.c_void_p)))
        assert owner.min() >= 0 and owner.max() < world
        load = np.bincount(owner, weights=counts, minlength=world)
        # LPT guarantee: max load <= mean + largest item
        assert load.max() <= counts.sum() / world + counts.max()
        owner2 = np.empty(267, np.int32)
        _lib.lib.rgcn_shard_plan(counts.ctypes.data_as(ctypes.c_void_p), 267, world,
                                 owner2.ctypes.data_as(ctypes.c_void_p))
        assert np.array_equal(owner, owner2)


def test_bad_arguments_return_errors_without_a_gpu():
    from torch_rgcn_b200 import _lib
    with pytest.raises(_lib.RgcnError):
        _lib.check(_lib.lib.rgcn_shard_plan(None, 3, 2, None))
    with pytest.raises(_lib.RgcnError):
        _lib.check(_lib.lib.rgcn_block_diag(None, 1, 0, 1, 1, None, None))
    assert _lib.lib.rgcn_graph_workspace_bytes(1000, 10, 3, 0, 0) > 1000 * 40


@pytest.mark.parametrize('name', ['nc_none_h_feat', 'nc_basis_h_featureless', 'nc_block_v_feat_odd', 'nc_diag_h'])
def test_nc_constructor_matches_reference_init(name):
    """Same seed -> same initial parameters as the reference constructor (RNG draw order, shapes, names)."""
    from torch_rgcn_b200.layers import RelationalGraphConvolutionNC
    meta, d, params, _ = load_golden(name)
    seed = {'nc_none_h_feat': 0, 'nc_basis_h_featureless': 5, 'nc_block_v_feat_odd': 12, 'nc_diag_h': 9}[name]
    torch.manual_seed(seed + 2)                 # tests/golden/make_golden.py seeds the constructor with seed+2
    layer = RelationalGraphConvolutionNC(triples=torch.as_tensor(d['triples_plus']), num_nodes=meta['N'],
                                         num_relations=meta['num_relations'], in_features=meta['in_features'],
                                         out_features=meta['out_features'], bias=meta['bias'],
                                         decomposition=meta['decomposition'], vertical_stacking=meta['vertical'],
                                         diag_weight_matrix=meta['diag'])
    got = dict(layer.named_parameters())
    assert sorted(got) == sorted(params)
    for n, p in got.items():
        if n != 'bias':                         # the fixture re-draws the bias after construction
            np.testing.assert_array_equal(p.detach().numpy(), params[n], err_msg=n)
    assert layer.out_features == meta['out_features']


@pytest.mark.parametrize('name,seed', [('lp_none_h', 20), ('lp_basis_h', 22), ('lp_block_h', 24)])
def test_lp_constructor_matches_reference_init(name, seed):
    from torch_rgcn_b200.layers import RelationalGraphConvolutionLP
    meta, d, params, _ = load_golden(name)
    torch.manual_seed(seed + 2)
    layer = RelationalGraphConvolutionLP(num_nodes=meta['N'], num_relations=meta['num_relations'],
                                         in_features=meta['in_features'], out_features=meta['out_features'],
                                         decomposition=meta['decomposition'], vertical_stacking=meta['vertical'],
                                         w_init='glorot-normal', b_init=meta['b_init'])
    got = dict(layer.named_parameters())
    assert sorted(got) == sorted(params)
    for n, p in got.items():
        if n != 'bias':
            np.testing.assert_array_equal(p.detach().cpu().numpy(), params[n], err_msg=n)


def test_constructor_error_behaviour():
    from torch_rgcn_b200.layers import RelationalGraphConvolutionNC, RelationalGraphConvolutionLP
    t = torch.zeros(3, 3, dtype=torch.long)
    with pytest.raises(NotImplementedError):
        RelationalGraphConvolutionNC(triples=t, num_nodes=3, num_relations=1, in_features=4, out_features=4,
                                     decomposition={'type': 'tucker'})
    with pytest.raises(AssertionError):
        RelationalGraphConvolutionNC(triples=t, num_nodes=3, num_relations=1, in_features=5, out_features=4,
                                     decomposition={'type': 'block', 'num_blocks': 2})
    with pytest.raises(NotImplementedError):
        RelationalGraphConvolutionNC(triples=t, num_nodes=3, num_relations=1, in_features=4, out_features=4,
                                     reset_mode='bogus')
    with pytest.raises(TypeError):             # reference layers.py:444-447: schlichtkrull-normal lacks `shape` here
        RelationalGraphConvolutionLP(num_nodes=3, num_relations=3, in_features=4, out_features=4,
                                     w_init='schlichtkrull-normal')
    diag = RelationalGraphConvolutionNC(triples=t, num_nodes=3, num_relations=1, in_features=4, out_features=9,
                                        diag_weight_matrix=True)
    assert diag.out_features == 4 and diag.bias is None and diag.weights.shape == (1, 4)


@pytest.mark.skipif(torch.cuda.is_available(), reason='checks the no-GPU failure mode')
def test_forward_without_cuda_fails_loudly():
    from torch_rgcn_b200.layers import RelationalGraphConvolutionNC
    t = torch.tensor([[0, 0, 1], [1, 1, 0], [0, 2, 0], [1, 2, 1]])
    layer = RelationalGraphConvolutionNC(triples=t, num_nodes=2, num_relations=3, in_features=4, out_features=4)
    with pytest.raises(RuntimeError, match='CUDA'):
        layer(torch.randn(2, 4))


@pytest.mark.skipif(torch.cuda.is_available(), reason='checks the no-GPU failure mode')
def test_link_prediction_helpers_without_cuda_fail_loudly():
    """Decoder, ranking evaluation, sampling and the LP models have no CPU path either."""
    from torch_rgcn_b200 import evaluation, sampling
    from torch_rgcn_b200.decoder import DistMult
    from torch_rgcn_b200.models import CompressionRelationPredictor
    t = torch.tensor([[0, 0, 1], [1, 1, 0], [2, 0, 2]])
    x = torch.randn(3, 4)
    with pytest.raises(RuntimeError, match='CUDA'):
        DistMult(2, 4, 3, 2)(t, x)
    with pytest.raises(RuntimeError, match='CUDA'):
        evaluation.rank_triples(t, x, torch.randn(2, 4), True)
    with pytest.raises(RuntimeError, match='CUDA'):
        evaluation.TrueTripleFilter(t, 3, 2)
    with pytest.raises(RuntimeError, match='CUDA'):
        sampling.EdgeNeighborhoodSampler(t, 3)
    for fn in (sampling.uniform_sampling, sampling.edge_neighborhood):
        with pytest.raises(RuntimeError, match='CUDA'):
            fn(t, 2, {0: 0, 1: 1, 2: 2})
    with pytest.raises(RuntimeError, match='CUDA'):
        sampling.edge_dropout(t, 0.5)
    enc = {'node_embedding': 4, 'hidden1_size': 4, 'num_layers': 1, 'weight_init': 'glorot-normal', 'bias_init': 'zeros',
           'edge_dropout': {'general': 0.5, 'self_loop': 0.2, 'self_loop_type': 'schlichtkrull-dropout'}}
    dec = {'l2_penalty_type': 'schlichtkrull-l2', 'l2_penalty': 0.01, 'weight_init': 'standard-normal'}
    model = CompressionRelationPredictor(nnodes=3, nrel=2, encoder_config=enc, decoder_config=dec)
    assert sorted(n for n, _ in model.named_parameters()) == sorted(
        ['node_embeddings', 'node_embeddings_bias', 'rgc1.weights', 'rgc1.bias', 'scoring_function.relations',
         'encoding_layer.weight', 'encoding_layer.bias', 'decoding_layer.weight', 'decoding_layer.bias'])
    with pytest.raises(RuntimeError, match='CUDA'):
        model(t, t)


def test_select_sampling_names():
    """reference utils/misc.py:112-119: the two method names, case-insensitive, anything else NotImplementedError"""
    from torch_rgcn_b200 import sampling
    assert sampling.select_sampling('Uniform') is sampling.uniform_sampling
    assert sampling.select_sampling('edge-neighborhood') is sampling.edge_neighborhood
    with pytest.raises(NotImplementedError):
        sampling.select_sampling('snowball')


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, 'torch_rgcn_b200')
    for fn in os.listdir(pkg):
        if fn.endswith('.py'):
            src = open(os.path.join(pkg, fn)).read()
            assert 'oracle' not in src.replace('no CPU', ''), fn


def test_layer_copies_and_pickles_without_its_plan_cache():
    """copy.deepcopy(model) / pickling (the best-checkpoint pattern) work after a forward has cached a GraphPlan, whose
    ctypes structs cannot be pickled: the copy drops the cache and rebuilds it lazily."""
    import copy
    import pickle
    import torch
    from torch_rgcn_b200 import _lib
    from torch_rgcn_b200.layers import RelationalGraphConvolutionNC
    t = torch.tensor([[0, 0, 1], [1, 1, 2]])
    layer = RelationalGraphConvolutionNC(triples=t, num_nodes=3, num_relations=2, in_features=4, out_features=4)
    layer._plan_cache = ('key', _lib.Graph())                 # what a forward leaves behind
    with pytest.raises(Exception):
        pickle.dumps(_lib.Graph())
    clone = copy.deepcopy(layer)
    assert clone._plan_cache is None and layer._plan_cache is not None
    assert torch.equal(clone.weights, layer.weights) and clone.weights is not layer.weights
    again = pickle.loads(pickle.dumps(layer))
    assert again._plan_cache is None and torch.equal(again.weights, layer.weights)
