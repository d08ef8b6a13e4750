import os
import sys
import zlib

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def engine():
    """The CUDA engine bound to cuda:0; fails loudly when the extension or the GPU is missing."""
    from lightkurve_b200 import engine as eng
    assert eng.device_count() > 0, "no CUDA device visible: -m gpu tests need the B200 box"
    eng.init(0)
    return eng


def seed_for(test_function_name):
    """Deterministic numpy global-RNG seed per test FUNCTION (parametrised variants share it)."""
    return zlib.crc32(test_function_name.encode()) & 0xFFFFFFFF


@pytest.fixture(autouse=True)
def _seed_numpy_global_rng(request):
    """Several tests ported from the reference draw noise from numpy's global RNG without seeding it.  Seed it per
    test function so that every run - and the CPU re-run of the same bodies on the oracle
    (tests/test_shim_on_oracle.py, which seeds identically) - sees the same data."""
    fn = getattr(request, "function", None)
    if fn is not None:
        np.random.seed(seed_for(fn.__name__))
    yield
