import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def example_sce():
    """Bundled example_sce (BASELINE config 1): cells x genes counts and genes x clones copy number."""
    Y = np.load(os.path.join(GOLDEN, "example_sce_counts.npy")).astype(np.float64)
    L = np.load(os.path.join(GOLDEN, "example_sce_cn.npy")).astype(np.float64)
    return Y, L


@pytest.fixture(scope="session")
def golden_c1():
    return dict(np.load(os.path.join(GOLDEN, "golden_c1.npz")))
