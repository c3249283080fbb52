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
    with np.load(os.path.join(GOLDEN, name + '.npz'), allow_pickle=False) as z:
        return {k: z[k] for k in z.files}


def rel_err(a, r):
    """max|a - r| / max|r| per tensor."""
    a = np.asarray(a, dtype=np.float64)
    r = np.asarray(r, dtype=np.float64)
    assert a.shape == r.shape, (a.shape, r.shape)
    den = np.max(np.abs(r))
    return float(np.max(np.abs(a - r)) / (den if den > 0 else 1.0))


@pytest.fixture
def golden():
    return load_golden
