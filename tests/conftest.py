import glob
import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN_DIR = os.path.join(ROOT, 'tests', 'golden')


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box with -m gpu)')


def load_golden(name):
    z = np.load(os.path.join(GOLDEN_DIR, name + '.npz'))
    d = {k: z[k] for k in z.files if k != 'meta'}
    meta = json.loads(str(z['meta']))
    params = {k[len('param_'):]: v for k, v in d.items() if k.startswith('param_')}
    grads = {k[len('grad_'):]: v for k, v in d.items() if k.startswith('grad_')}
    return meta, d, params, grads


def golden_names(prefix=''):
    return sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN_DIR, prefix + '*.npz')))


@pytest.fixture(scope='session')
def cuda_device():
    import torch
    if not torch.cuda.is_available():
        pytest.skip('no CUDA device')
    return torch.device('cuda:0')
