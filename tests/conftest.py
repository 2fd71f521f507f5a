import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def oracle_built():
    import oracle
    oracle.build()
    return True


@pytest.fixture(scope="session")
def ref(oracle_built):
    """The compiled, unmodified reference (oracle/_ref/libref_oracle.so)."""
    import oracle
    if not oracle.have_ref():
        pytest.skip("oracle/_ref/libref_oracle.so not built (needs /root/reference at build time)")
    return oracle.RefOracle()


@pytest.fixture(scope="session")
def restated(oracle_built):
    import oracle
    return oracle.Restated()


@pytest.fixture(scope="session")
def engine():
    """The CUDA engine; GPU tests fail (not skip) if the extension is missing."""
    import gfx_imagecompress_b200 as g
    g.load_library()
    g.init(0)
    return g
