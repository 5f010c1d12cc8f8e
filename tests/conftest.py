import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests", "golden")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden():
    return dict(np.load(os.path.join(ROOT, "tests", "golden", "golden.npz")))


@pytest.fixture(scope="session")
def golden_configs():
    """outputs of the unmodified reference on the named BASELINE configs (tests/golden/make_golden_configs.py)"""
    return dict(np.load(os.path.join(ROOT, "tests", "golden", "golden_configs.npz")))
