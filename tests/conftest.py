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
    """Make sure libfft_b200.so and the oracle are compiled (nvcc cross-compiles without a GPU)."""
    import __graft_entry__ as ge
    ge.build()
    return ge


@pytest.fixture(scope="session")
def fft(built):
    return built.load_package()


@pytest.fixture(scope="session")
def oracle(built):
    import oracle as o
    return o
