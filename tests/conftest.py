import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def gpu():
    """Initialise libdaftb200 on cuda:0; GPU tests fail loudly (not skip) if the native library is absent."""
    import daft_b200
    daft_b200.init(0)
    return daft_b200
