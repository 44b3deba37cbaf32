import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session", autouse=True)
def _built():
    """The C-ABI library and the oracle are built in-tree; build them if a fresh checkout lacks them."""
    import __graft_entry__ as g
    so = os.path.join(ROOT, "gpvecchia_b200", "libgpvecchia_b200.so")
    orc = os.path.join(ROOT, "oracle", "liboracle.so")
    if not (os.path.exists(so) and os.path.exists(orc)):
        g.build()
    yield
