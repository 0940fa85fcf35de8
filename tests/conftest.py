import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box via gpurun)")


@pytest.fixture(scope="session")
def lib():
    """The built C-ABI library; building is a CPU-only cross-compile."""
    import __graft_entry__ as ge
    ge.build()
    from fof_b200 import _lib
    return _lib


@pytest.fixture(scope="session")
def handle(lib):
    return lib.get_handle(0)
