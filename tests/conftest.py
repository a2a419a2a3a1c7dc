import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session", autouse=True)
def _oracle_built():
    """The oracle's C restatement is a checker; build it on demand (gcc only)."""
    so = os.path.join(ROOT, "oracle", "_build", "libdf_oracle.so")
    if not os.path.exists(so):
        import subprocess
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "all"])
    yield


def golden(name):
    import numpy as np
    return np.load(os.path.join(GOLDEN, name + ".npz"))
