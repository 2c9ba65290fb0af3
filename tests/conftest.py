import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def built():
    """C-ABI library + oracle, compiled in-tree (nvcc cross-compiles without a GPU)."""
    import __graft_entry__
    so = os.path.join(ROOT, "gofeature_b200", "libgofeature_b200.so")
    if not os.path.exists(so):
        __graft_entry__.build()
    from oracle import lib as ol
    ol.build()
    return so


@pytest.fixture(scope="session")
def gpu_ctx(built):
    from gofeature_b200 import api
    ctx = api.NewContext(0)  # raises ErrCreateContext without a GPU: no fallback
    yield ctx
    ctx.Close()
