import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def built_lib():
    """Build (if stale) and load the CUDA library; GPU tests must run the native path, never a fallback."""
    sys.path.insert(0, ROOT)
    import __graft_entry__ as ge
    ge.build()
    from fwgym_b200 import _capi
    return _capi.lib()
