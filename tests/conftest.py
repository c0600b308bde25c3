import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, 'tests', 'golden')


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box)')


def load_golden(name):
    """-> (params dict, arrays dict); the 'input' indirection lets fixtures share one waveform"""
    with np.load(os.path.join(GOLDEN, name + '.npz')) as z:
        arrays = {k: z[k] for k in z.files if k != 'params'}
        params = json.loads(str(z['params']))
    if 'input' in params:
        with np.load(os.path.join(GOLDEN, params.pop('input') + '.npz')) as z:
            arrays['x'] = z['x']
    if isinstance(params.get('window'), list):
        params['window'] = tuple(params['window'])
    return params, arrays


@pytest.fixture(scope='session')
def cuda_device():
    import torch

    if not torch.cuda.is_available():
        pytest.skip('no CUDA device')
    return torch.device('cuda', 0)
