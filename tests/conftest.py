import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import szload  # noqa: E402,F401


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    config.addinivalue_line("markers", "slow: takes minutes (benchmark-size fields); still part of -m gpu")


@pytest.fixture(scope="session")
def oracle_lib():
    from oracle import szo
    szo.build()
    return szo.oracle()


@pytest.fixture(scope="session")
def product_lib():
    """The CUDA product library; GPU tests fail (not skip) when it cannot be loaded."""
    from subzero_jl_b200 import capi
    return capi.product()
