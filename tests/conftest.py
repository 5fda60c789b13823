import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session", autouse=True)
def _build_native():
    """Build the oracle (CPU checker) and the CUDA library (nvcc cross-compiles without a GPU)."""
    from oracle import binding as ob
    ob.build()
    from ggdmc_b200 import _lib
    _lib.build()
