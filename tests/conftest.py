import os
import sys

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
