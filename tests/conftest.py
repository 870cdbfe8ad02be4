import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import __graft_entry__ as graft  # noqa: E402

pkg = graft.load_package()


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def lib():
    """The C-ABI library; building it needs nvcc only, loading it needs no GPU."""
    if not os.path.exists(pkg.ttmlblend.LIB_PATH):
        graft.build()
    return pkg.load_library()


@pytest.fixture(scope="session")
def oracle_lib():
    from oracle import oracle
    return oracle.load()


@pytest.fixture(scope="session")
def ctx(lib):
    """One context on cuda:0 for the whole session. No GPU -> hard failure, not a skip:
    the product has no CPU fallback and a GPU test must never pass without the CUDA path."""
    c = pkg.TtmlBlend(0)
    yield c
    c.close()
