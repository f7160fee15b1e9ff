import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")


@pytest.fixture(scope="session")
def entry():
    import __graft_entry__ as g

    # the shared libraries are built by __graft_entry__.build(); build here if a test run comes first
    if not os.path.exists(os.path.join(g.PKG_DIR, "libgmrfb.so")):
        g.build()
    return g


@pytest.fixture(scope="session")
def pkg(entry):
    return entry.load_pkg()


@pytest.fixture(scope="session")
def orc(entry):
    o = entry.load_oracle()
    o.build()
    return o


@pytest.fixture(scope="session")
def W(pkg):
    return pkg.workloads


@pytest.fixture(scope="session")
def ctx(pkg):
    """A GPU context; gpu-marked tests fail loudly (not skip) if the device or the library is unusable."""
    return pkg.default_context()
