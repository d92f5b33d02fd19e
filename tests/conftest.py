import json
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, 'tests', 'golden')


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box with -m gpu)')


@pytest.fixture(scope='session')
def golden_norms():
    with open(os.path.join(GOLDEN, 'error_norms.json')) as f:
        return json.load(f)


@pytest.fixture(scope='session')
def golden_iterations():
    with open(os.path.join(GOLDEN, 'amg1_iterations.json')) as f:
        return json.load(f)


@pytest.fixture(scope='session')
def golden_example2():
    with open(os.path.join(GOLDEN, 'example2_tables.json')) as f:
        return json.load(f)


@pytest.fixture(scope='session')
def cuda():
    import torch
    if not torch.cuda.is_available():
        pytest.fail('GPU test selected but no CUDA device is visible')
    return torch
