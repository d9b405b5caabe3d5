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
def golden():
    return np.load(os.path.join(GOLDEN, "compression_golden.npz"))


@pytest.fixture(scope="session")
def cases():
    return np.load(os.path.join(GOLDEN, "compression_cases.npz"))


@pytest.fixture(scope="session")
def built():
    """Make sure the native libraries exist (cheap when up to date)."""
    import __graft_entry__ as g
    g.build()
    return True


def checksum_np(wit, ws):
    """numpy version of b3w_checksum_device (include/blake3wit.h)."""
    w = np.ascontiguousarray(wit).reshape(-1, ws * 32).view(np.uint64)
    e = np.arange(ws * 4, dtype=np.uint64)
    with np.errstate(over="ignore"):
        mix = (e + np.uint64(1)) * np.uint64(0x9E3779B97F4A7C15)
        return ((w + np.uint64(1)) * mix[None, :]).sum(axis=1, dtype=np.uint64)
