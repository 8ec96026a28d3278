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


@pytest.fixture(scope="module")
def emulated_library():
    """Point the ctypes loader at the CPU-EMULATED build of the C-ABI (tests/cuda_emul/build.py) for one test module.

    Test infrastructure: the real core.cu + every kernel except the tcgen05/TMA ones are compiled for the host against
    a fiber-based emulation of the CUDA execution model, so host code, ABI and kernels can be checked against the oracle
    without a GPU.  The product loader is restored afterwards (it only ever knows clonealign_b200/libclonealign_b200.so).
    """
    from clonealign_b200 import _lib
    sys.path.insert(0, os.path.join(ROOT, "tests", "cuda_emul"))
    try:
        import build as emul_build
    finally:
        sys.path.pop(0)
    path = emul_build.build()
    saved = (_lib.LIB_PATH, _lib._lib)
    _lib.LIB_PATH, _lib._lib = path, None
    try:
        yield path
    finally:
        _lib.LIB_PATH, _lib._lib = saved
