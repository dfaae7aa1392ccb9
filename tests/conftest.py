import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def oracle_libs():
    """Compile the oracle restatement (and the reference into oracle/_ref when /root/reference exists)."""
    from oracle import harness
    harness.build(ref=True)
    return harness


@pytest.fixture(scope="session")
def cuda_lib():
    """Compile the CUDA extension if stale and load it. No fallback: a missing library is an error."""
    from daqp_b200 import build
    build.build()
    import daqp_b200
    return daqp_b200.lib()
