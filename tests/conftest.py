"""pytest configuration: markers + shared loaders.

`-m "not gpu"` (run in the CPU-only build container) covers the oracle against the golden vectors,
the host-side logic and the C-ABI export check; `-m gpu` (run on a B200) holds the parity tests
proper, which call the CUDA path through the C-ABI.
"""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    config.addinivalue_line("markers", "slow: takes more than a few seconds on CPU")


def bits_equal(a, b):
    """Bit-for-bit equality of two float64 arrays (NaNs compare equal iff both NaN)."""
    a = np.ascontiguousarray(a, dtype=np.float64)
    b = np.ascontiguousarray(b, dtype=np.float64)
    an, bn = np.isnan(a), np.isnan(b)
    return bool(np.array_equal(an, bn) and np.array_equal(a[~an].view(np.int64), b[~bn].view(np.int64)))


@pytest.fixture(scope="session")
def golden_dir():
    return os.path.join(ROOT, "tests", "golden")
