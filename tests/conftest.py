import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    config.addinivalue_line("markers", "slow: long CPU test, excluded from the default CPU suite")


@pytest.fixture(scope="session")
def ref():
    import zref
    return zref.Ref()


@pytest.fixture(scope="session")
def oracle():
    import zref
    return zref.Oracle()
