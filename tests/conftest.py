import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def refdata():
    return np.load(os.path.join(GOLDEN, "reference_data.npz"))


@pytest.fixture(scope="session")
def vectors():
    return np.load(os.path.join(GOLDEN, "oracle_vectors.npz"))


@pytest.fixture(scope="session")
def oracle_mod():
    import oracle

    oracle.build()
    return oracle


def relerr(a, b):
    """max|a-b| / max|b| -- the parity measure of SURVEY.md section 8(d)."""
    a, b = np.asarray(a), np.asarray(b)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))
