import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden():
    return np.load(os.path.join(ROOT, "tests", "golden", "hotpath_v1.npz"))


@pytest.fixture(scope="session")
def co():
    """C oracle (oracle/liboracle.so), built on demand."""
    from oracle import c_oracle

    c_oracle.build()
    return c_oracle
