import os
import sys

import numpy as np
import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if REPO not in sys.path:
    sys.path.insert(0, REPO)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on the B200 box)")


@pytest.fixture(scope="session")
def golden():
    return np.load(os.path.join(REPO, "tests", "golden", "reference_v1_11_0.npz"))


@pytest.fixture(scope="session")
def bed_fixtures():
    return np.load(os.path.join(REPO, "tests", "golden", "reference_bed_fixtures.npz"))


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle as orc
    orc.build(ref=os.path.isdir("/root/reference/rocco"))
    return orc


def rel_err(a, b, floor=1e-3):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.max(np.abs(a - b) / np.maximum(np.abs(b), floor))) if a.size else 0.0
