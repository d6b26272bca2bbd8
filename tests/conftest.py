import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def port():
    """The C restatement oracle (oracle/port), built on demand with gcc."""
    from oracle import pyport
    pyport.lib()
    return pyport


@pytest.fixture(scope="session")
def ref():
    """The reference's own sources compiled by oracle/Makefile (only where oracle/_ref exists)."""
    from oracle import pyref
    if not pyref.available():
        pytest.skip("oracle/_ref/libvvref.so not built (needs /root/reference)")
    os.environ.setdefault("OMP_NUM_THREADS", "1")
    return pyref


@pytest.fixture(scope="session")
def ctx():
    from vvflow_b200 import capi
    c = capi.Context(0)
    yield c
    c.close()


@pytest.fixture(autouse=True)
def _fresh_tree(request):
    """a failed GPU test must not leave the shared context with a built tree"""
    yield
    if "ctx" in request.fixturenames:
        try:
            request.getfixturevalue("ctx").tree_destroy()
        except Exception:
            pass
