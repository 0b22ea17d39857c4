import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu on the GPU box")


@pytest.fixture(scope="session")
def ctx():
    """Device context for -m gpu tests.  No skip-on-missing-library: the CUDA path must be the
    one that runs, so a missing libozl_b200.so or device is a hard failure."""
    import openzl_b200
    c = openzl_b200.Context(0)
    yield c
    c.close()
