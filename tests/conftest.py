import os
import sys

import pytest

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
for p in (ROOT, os.path.dirname(__file__)):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def sim_lib():
    """Path of the kernel-source-on-CPU simulator library (built by __graft_entry__.build())."""
    from common import SIM_LIB, ensure_built
    ensure_built()
    return SIM_LIB
