import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def ora():
    """The CPU oracle (test infrastructure; oracle/esvio_oracle.h)."""
    from oracle import oracle
    oracle.build()
    return oracle


@pytest.fixture(scope="session")
def golden_lk():
    return np.load(os.path.join(GOLDEN, "lk_golden.npz"))


@pytest.fixture(scope="session")
def golden_fmat():
    return np.load(os.path.join(GOLDEN, "fmat_golden.npz"))


@pytest.fixture(scope="session")
def golden_misc():
    return np.load(os.path.join(GOLDEN, "misc_golden.npz"))


@pytest.fixture(scope="session")
def golden_imgops():
    return np.load(os.path.join(GOLDEN, "imgops_golden.npz"))


@pytest.fixture(scope="session")
def golden_frames():
    return np.load(os.path.join(GOLDEN, "frames_golden.npz"))


@pytest.fixture(scope="session")
def capi():
    """libesvio_fe.so through ctypes; the GPU tests must run on the native library."""
    from esvio_b200 import _capi
    if not os.path.exists(_capi.LIB_PATH):
        _capi.build()
    _capi.lib()
    return _capi
