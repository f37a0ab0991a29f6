import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu under gpurun)")


@pytest.fixture(scope="session")
def dunk():
    import cubesat_apds_b200 as d
    return d


@pytest.fixture(scope="session")
def ctx(dunk):
    """GPU tests only: a real context.  Fails loudly (no fallback) without a B200."""
    c = dunk.Context(0, 4)
    yield c
    c.close()


@pytest.fixture(scope="session")
def match_golden():
    return np.load(os.path.join(GOLDEN, "match_golden.npz"))
