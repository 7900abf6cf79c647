"""The reference's OWN unit tests (tests/test_nn.py, tests/test_utils.py of thiviyanT/torch-rgcn), executed unchanged
against the drop-in: `torch_rgcn.layers` / `torch_rgcn.utils` are patched the way INTEGRATION.md documents, then every
`test_*` function of the untouched test modules is called.

The test files are not part of this repo: __graft_entry__.build() copies them next to the pip-installed reference
(baseline/_ref/reference_tests, git-ignored) so that they travel to the GPU box; without that directory the tests skip.
The reference tests build their layers from CPU tensors and never move them, so the patched classes place themselves
on the CUDA device at construction (a user would call `.cuda()`); nothing else differs from a user's import swap."""
import importlib.util
import os
import sys

import pytest
import torch

from conftest import ROOT

pytestmark = pytest.mark.gpu
REF = os.path.join(ROOT, 'baseline', '_ref')
REF_TESTS = os.path.join(REF, 'reference_tests')


@pytest.fixture()
def patched_reference(cuda_device, monkeypatch):
    if not os.path.isdir(REF_TESTS):
        pytest.skip('baseline/_ref/reference_tests missing (run __graft_entry__.build() where /root/reference exists)')
    monkeypatch.syspath_prepend(REF)
    import torch_rgcn.layers as ref_layers                     # the unmodified reference package
    import torch_rgcn.utils as ref_utils
    import torch_rgcn_b200.layers as fast
    import torch_rgcn_b200.utils as fast_utils

    class NCOnDevice(fast.RelationalGraphConvolutionNC):
        def __init__(self, *a, **k):
            super().__init__(*a, **k)
            self.to(cuda_device)

    class LPOnDevice(fast.RelationalGraphConvolutionLP):
        def __init__(self, *a, **k):
            super().__init__(*a, **k)
            self.to(cuda_device)

    monkeypatch.setattr(ref_layers, 'RelationalGraphConvolutionNC', NCOnDevice)
    monkeypatch.setattr(ref_layers, 'RelationalGraphConvolutionLP', LPOnDevice)
    for name in ('add_inverse_and_self', 'generate_inverses', 'generate_self_loops', 'stack_matrices', 'sum_sparse',
                 'block_diag', 'drop_edges'):
        monkeypatch.setattr(ref_utils, name, getattr(fast_utils, name))
    return REF_TESTS


def _load(path, name):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


@pytest.mark.parametrize('module', ['test_nn', 'test_utils'])
def test_reference_unit_tests_pass_unchanged(patched_reference, module):
    mod = _load(os.path.join(patched_reference, module + '.py'), 'reference_' + module)
    tests = [n for n in dir(mod) if n.startswith('test_') and callable(getattr(mod, n))]
    assert tests, f'no tests found in the reference {module}.py'
    for name in tests:
        torch.manual_seed(0)
        getattr(mod, name)()                                    # raises on any failed assertion
