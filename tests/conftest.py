import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import __graft_entry__ as entry  # noqa: E402

entry.load_package()
ORACLE = entry.load_oracle()


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def oracle():
    ORACLE.oracle_lib()
    return ORACLE


@pytest.fixture(scope="session")
def capi():
    from graph_embed_b200 import build, capi as c
    if not os.path.exists(c.LIB_PATH):
        build.build_library()
    return c


@pytest.fixture(scope="session")
def graphs():
    from graph_embed_b200 import graphs as g
    return g


@pytest.fixture(scope="session")
def ctx(capi):
    """A device context; GPU tests fail loudly (no fallback) if there is no B200."""
    c = capi.Context(0)
    yield c
    c.close()
