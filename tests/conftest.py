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
    """Build the checker (oracle) and the engine once per session; nvcc cross-compiles without a GPU."""
    from oracle import sv_oracle
    sv_oracle.build(ref=True)
    from rocquantum_b200 import build
    build.build()
    yield
