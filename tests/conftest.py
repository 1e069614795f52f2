import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")


@pytest.fixture(scope="session")
def golden():
    """Vectors produced by the UNMODIFIED reference (tests/golden/make_golden.py)."""
    return dict(np.load(os.path.join(ROOT, "tests", "golden", "reference_golden.npz")))


@pytest.fixture(scope="session")
def native_lib():
    from wtracker_b200 import build

    build.build()
    from wtracker_b200 import _lib

    return _lib.lib()
import os,sys; sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
