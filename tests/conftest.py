import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def lib():
    """The C-ABI shared library; built in-tree if missing (nvcc cross-compiles without a GPU)."""
    from mepp_b200 import build
    build.build()
    from mepp_b200 import _lib
    return _lib
