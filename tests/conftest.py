import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session", autouse=True)
def _built():
    """The oracle is test infrastructure and cheap to build; the product library must already exist
    (quartz_b200/build.sh, run by __graft_entry__.build()) — tests never compile CUDA."""
    from tests import oracle_ffi
    oracle_ffi.lib()
    yield
