import json
import os
import sys

import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if REPO not in sys.path:
    sys.path.insert(0, REPO)
GOLDEN = os.path.join(REPO, 'tests', 'golden')


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box with -m gpu)')


def load_golden(name):
    with open(os.path.join(GOLDEN, name)) as f:
        return json.load(f)['cases']


def unhex(xs):
    return [float.fromhex(x) for x in xs]


@pytest.fixture(scope='session')
def golden_stencils():
    return load_golden('stencils.json')


@pytest.fixture(scope='session')
def golden_history():
    return load_golden('history.json')


@pytest.fixture(scope='session')
def golden_cfg1():
    return load_golden('cfg1_api.json')
