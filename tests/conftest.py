import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle_mod():
    import oracle
    oracle.build()
    return oracle


@pytest.fixture(scope="session")
def fgq():
    """The product package, with the CUDA library loaded; GPU tests fail loudly if the .so or the
    device is missing (there is no fallback to hide behind)."""
    import fgq_b200
    L = fgq_b200.lib()
    assert L.fgq_device_count() > 0, "no CUDA device visible"
    return fgq_b200
