"""B200-native RGCN relational message-passing engine — drop-in for thiviyanT/torch-rgcn's RGCN layers.

    from torch_rgcn_b200.layers import RelationalGraphConvolutionNC, RelationalGraphConvolutionLP

Importing this package loads librgcn_b200.so (hand-written sm_100a CUDA behind the C ABI in
include/rgcn_b200.h).  There is no CPU or PyTorch fallback.
"""
from . import _lib                                   # noqa: F401  (fails loudly if the library is missing)
from .graph import GraphPlan                         # noqa: F401
from .functional import rgcn_propagate               # noqa: F401
from .layers import RelationalGraphConvolutionNC, RelationalGraphConvolutionLP   # noqa: F401
from .decoder import DistMult, negative_sampling      # noqa: F401

__all__ = ['GraphPlan', 'rgcn_propagate', 'RelationalGraphConvolutionNC', 'RelationalGraphConvolutionLP', 'DistMult',
           'negative_sampling']
