import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def ob():
    """the CPU oracle bindings (test infrastructure)"""
    from oracle import bindings
    bindings.build()
    return bindings


@pytest.fixture(scope="session")
def api():
    """the product's ctypes mirror; builds the library if it is not there yet"""
    from turner_b200 import api as a
    if not os.path.exists(a.LIB_PATH):
        a.build()
    return a


@pytest.fixture(scope="session")
def scenes():
    from turner_b200 import scenes as s
    return s


@pytest.fixture(scope="session")
def ref_ok(ob):
    if not ob.ref_available():
        pytest.skip("oracle/_ref not built (needs /root/reference)")
    return True
