import os
import sys

# test_gpu_sharded runs all ranks of a virtual world on ONE device, one stream per rank, with
# device-side barriers between them: the ranks' streams must map to distinct hardware queues or a
# spinning barrier kernel serialises the rank it waits for (default: 8 queues for 32 pool streams).
os.environ.setdefault('CUDA_DEVICE_MAX_CONNECTIONS', '32')

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, 'oracle')):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN = os.path.join(ROOT, 'tests', 'golden')


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a B200 (run with -m gpu on the GPU box)')


@pytest.fixture(scope='session')
def golden_small():
    import numpy as np
    return np.load(os.path.join(GOLDEN, 'reference_small.npz'))


@pytest.fixture(scope='session')
def golden_s8():
    import numpy as np
    return np.load(os.path.join(GOLDEN, 'reference_s8.npz'))


def rel_fro(a, b):
    """Relative Frobenius error of a against b (b is the reference)."""
    import numpy as np
    denom = np.linalg.norm(b)
    return float(np.linalg.norm(a - b) / (denom if denom > 0 else 1.0))
