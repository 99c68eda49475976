import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def oracle_lib():
    from alis_b200 import capi
    return capi.load_oracle()


@pytest.fixture(scope="session")
def engine_lib():
    """The CUDA engine -- fails (does not skip) when the extension is missing."""
    from alis_b200 import capi
    return capi.load_engine()
