"""ctypes binding of librgcn_b200.so (the C ABI declared in include/rgcn_b200.h).

The library is the product: there is no Python / PyTorch fallback.  If it cannot be loaded the
import of this module raises, and every layer call needs a CUDA device.
"""
import ctypes as C
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, 'librgcn_b200.so')
CSRC = os.path.join(_HERE, 'csrc')

# enum values of include/rgcn_b200.h
NORM_ROW, NORM_COL_SWAPPED, NORM_EXPLICIT = 0, 1, 2
W_DENSE, W_BASIS, W_BLOCK, W_DIAG = 0, 1, 2, 3
F32, BF16 = 0, 1
FUSE_REC_WORDS = 36          # RGCN_FUSE_REC_WORDS

_p = C.c_void_p
_i64 = C.c_int64


class Tiling(C.Structure):
    _fields_ = [('tilerow', _p), ('row', _p), ('col', _p), ('rel', _p), ('slot', _p), ('val', _p),
                ('stepptr', _p), ('slotneed', _p), ('items', _p)]


class Fused(C.Structure):
    _fields_ = [('col', _p), ('rec', _p), ('blk_tile', _p), ('items', _p), ('meta', _p)]


class Graph(C.Structure):
    _fields_ = [('num_nodes', _i64), ('num_rels', _i64), ('nnz', _i64),
                ('d_rowptr', _p), ('d_src', _p), ('d_rel', _p), ('d_val', _p),
                ('s_rowptr', _p), ('s_dst', _p), ('s_rel', _p), ('s_val', _p),
                ('r_relptr', _p), ('r_dst', _p), ('r_src', _p), ('r_val', _p),
                ('r_dslot', _p), ('r_sslot', _p), ('r_chunkptr', _p),
                ('val', _p), ('status', _p), ('d_long', _p), ('s_long', _p), ('num_long_dst', _i64), ('num_long_src', _i64), ('max_rel_edges', _i64),
                ('tile_edges', _i64), ('num_tiles', _i64), ('tile_capacity', _i64), ('ring_depth', _i64),
                ('ft', Tiling), ('bt', Tiling),
                ('fuse_rows', _i64), ('fuse_cap', _i64), ('fuse_item_tiles', _i64), ('fuse_dirs', _i64),
                ('fuse_items', _i64 * 2), ('fuse_split', _i64 * 2), ('fuse_tiles', _i64 * 2), ('ff', Fused), ('fb', Fused)]


class Params(C.Structure):
    _fields_ = [('form', C.c_int32), ('featureless', C.c_int32),
                ('in_dim', _i64), ('out_dim', _i64), ('num_bases', _i64), ('num_blocks', _i64),
                ('num_block_rels', _i64),
                ('weights', _p), ('bases', _p), ('comps', _p), ('blocks', _p), ('blocks_self', _p),
                ('bias', _p), ('self_mask', _p), ('out_dtype', C.c_int32), ('pad_', C.c_int32),
                ('row_lo', _i64), ('row_hi', _i64), ('peer_out', _p * 8), ('num_peer_out', C.c_int32),
                ('pad2_', C.c_int32)]


class Grads(C.Structure):
    _fields_ = [('features', _p), ('weights', _p), ('bases', _p), ('comps', _p), ('blocks', _p),
                ('blocks_self', _p), ('bias', _p), ('features_dtype', C.c_int32)]


def build(force=False):
    """Compile the library in-tree with nvcc for sm_100a (cross-compiles without a GPU)."""
    if force:
        subprocess.run(['make', '-C', CSRC, 'clean'], check=True, stdout=subprocess.DEVNULL)
    subprocess.run(['make', '-C', CSRC, '-j4'], check=True)
    return LIB_PATH


def _load():
    if not os.path.exists(LIB_PATH):
        try:
            build()
        except Exception as exc:  # noqa: BLE001
            raise RuntimeError(
                f'torch_rgcn_b200: {LIB_PATH} is missing and could not be built ({exc}). '
                'The engine has no CPU or PyTorch fallback; run `make -C torch_rgcn_b200/csrc`.') from exc
    lib = C.CDLL(LIB_PATH)
    sigs = {
        'rgcn_last_error': (C.c_char_p, []),
        'rgcn_abi_version': (C.c_int, []),
        'rgcn_launch_count': (_i64, []),
        'rgcn_add_inverse_and_self': (C.c_int, [_p, _i64, _i64, _i64, _p, _p]),
        'rgcn_generate_inverses': (C.c_int, [_p, _i64, _i64, _p, _p]),
        'rgcn_lp_triples_plus': (C.c_int, [_p, _i64, _i64, _p, _i64, _p, _p]),
        'rgcn_stack_matrices': (C.c_int, [_p, _i64, _i64, _i64, C.c_int, _p, _p, _p]),
        'rgcn_sum_sparse': (C.c_int, [_p, _p, _i64, _i64, _i64, C.c_int, _p, _p, _p]),
        'rgcn_block_diag': (C.c_int, [_p, _i64, _i64, _i64, _i64, _p, _p]),
        'rgcn_graph_workspace_bytes': (C.c_size_t, [_i64, _i64, _i64, _i64, _i64]),
        'rgcn_fused_items_bound': (_i64, [_i64, _i64, _i64, _i64]),
        'rgcn_tile_items_bound': (_i64, [_i64, _i64, _i64, _i64]),
        'rgcn_tile_steps_len': (_i64, [_i64, _i64, _i64]),
        'rgcn_graph_build': (C.c_int, [_p, _i64, _i64, _i64, C.c_int, _i64, _i64, _p, C.POINTER(Graph), _p,
                                       C.c_size_t, _p]),
        'rgcn_forward_workspace_bytes': (C.c_size_t, [C.POINTER(Graph), C.POINTER(Params), C.c_int]),
        'rgcn_forward': (C.c_int, [C.POINTER(Graph), C.POINTER(Params), _p, C.c_int, _p, _p, C.c_size_t, _p]),
        'rgcn_backward_workspace_bytes': (C.c_size_t, [C.POINTER(Graph), C.POINTER(Params), C.c_int]),
        'rgcn_backward': (C.c_int, [C.POINTER(Graph), C.POINTER(Params), _p, C.c_int, _p, C.POINTER(Grads), _p,
                                    C.c_size_t, _p]),
        'rgcn_shard_plan': (C.c_int, [_p, _i64, C.c_int32, _p]),
        'rgcn_widen_rows': (C.c_int, [_p, _i64, _p, _p]),
        'rgcn_distmult_forward': (C.c_int, [_p, _i64, _p, _i64, _p, _i64, _i64, _p, _p, _p, _p, _p, _p]),
        'rgcn_distmult_backward': (C.c_int, [_p, _i64, _p, _i64, _p, _i64, _i64, _p, _p, _p, _p, _p, _p, _p]),
        'rgcn_distmult_penalty_workspace_bytes': (C.c_size_t, [_i64, _i64]),
        'rgcn_distmult_penalty': (C.c_int, [_p, _i64, _p, _i64, _p, _i64, _i64, _p, _p, _p, C.c_size_t, _p]),
        'rgcn_distmult_penalty_backward': (C.c_int, [_p, _i64, _p, _i64, _p, _i64, _i64, _p, _p, _p, _p]),
        'rgcn_corrupt_triples': (C.c_int, [_p, _p, _p, _i64, _p]),
        'rgcn_rank_filter_workspace_bytes': (C.c_size_t, [_i64]),
        'rgcn_rank_build_filter': (C.c_int, [_p, _i64, _i64, _i64, C.c_int, _p, _p, _p, _p, C.c_size_t, _p]),
        'rgcn_rank_workspace_bytes': (C.c_size_t, [_i64, _i64]),
        'rgcn_rank_triples': (C.c_int, [_p, _i64, C.c_int, _p, _i64, _p, _i64, _i64, _p, _p, _p, _p, _p, _i64, _p, _p, _p,
                                        C.c_size_t, _p]),
        'rgcn_sampler_build_workspace_bytes': (C.c_size_t, [_i64]),
        'rgcn_sampler_build': (C.c_int, [_p, _i64, _i64, _p, _p, _p, _p, C.c_size_t, _p]),
        'rgcn_sample_workspace_bytes': (C.c_size_t, [_i64, _i64]),
        'rgcn_sample_edge_neighborhood': (C.c_int, [_p, _p, _i64, _i64, _p, _i64, _p, _p, _p, C.c_size_t, _p]),
        'rgcn_take_triples': (C.c_int, [_p, _i64, _p, C.c_int, _i64, _p, _p, _p]),
    }
    for name, (res, args) in sigs.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    return lib


lib = _load()
EXPORTS = ['rgcn_last_error', 'rgcn_abi_version', 'rgcn_launch_count', 'rgcn_add_inverse_and_self',
           'rgcn_generate_inverses', 'rgcn_lp_triples_plus', 'rgcn_stack_matrices', 'rgcn_sum_sparse',
           'rgcn_block_diag', 'rgcn_graph_workspace_bytes', 'rgcn_tile_items_bound', 'rgcn_tile_steps_len', 'rgcn_fused_items_bound', 'rgcn_graph_build', 'rgcn_forward_workspace_bytes',
           'rgcn_forward', 'rgcn_backward_workspace_bytes', 'rgcn_backward', 'rgcn_shard_plan', 'rgcn_widen_rows',
           'rgcn_distmult_forward', 'rgcn_distmult_backward', 'rgcn_distmult_penalty_workspace_bytes',
           'rgcn_distmult_penalty', 'rgcn_distmult_penalty_backward', 'rgcn_corrupt_triples',
           'rgcn_rank_filter_workspace_bytes', 'rgcn_rank_build_filter', 'rgcn_rank_workspace_bytes', 'rgcn_rank_triples',
           'rgcn_sampler_build_workspace_bytes', 'rgcn_sampler_build', 'rgcn_sample_workspace_bytes',
           'rgcn_sample_edge_neighborhood', 'rgcn_take_triples']


class RgcnError(RuntimeError):
    pass


def check(rc):
    if rc != 0:
        raise RgcnError(f'rgcn_b200 error {rc}: {lib.rgcn_last_error().decode()}')


def ptr(t):
    """Device pointer of a tensor (None -> NULL)."""
    return None if t is None else C.c_void_p(t.data_ptr())


def stream_ptr():
    import torch
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def require_cuda(*tensors):
    import torch
    if not torch.cuda.is_available():
        raise RuntimeError('torch_rgcn_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback')
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise RuntimeError('torch_rgcn_b200: expected CUDA tensors; move the module and its inputs to the GPU')
