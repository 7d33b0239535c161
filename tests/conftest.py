import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def load_golden(name):
    return np.load(os.path.join(GOLDEN, name), allow_pickle=False)


@pytest.fixture(scope="session")
def golden_c():
    return load_golden("appendix_c.npz")


@pytest.fixture(scope="session")
def golden_c3():
    return load_golden("c3_like.npz")


@pytest.fixture(scope="session")
def golden_std():
    return load_golden("std_norm.npz")


@pytest.fixture(scope="session")
def golden_c2():
    return load_golden("c2_anchor.npz")


@pytest.fixture(scope="session")
def golden_c2_full():
    return load_golden("c2_full.npz")


@pytest.fixture(scope="session")
def golden_c2_truth():
    return load_golden("c2_truth.npz")


@pytest.fixture(scope="session")
def golden_c2_truth_full():
    return load_golden("c2_truth_full.npz")


@pytest.fixture(scope="session")
def golden_r2():
    return load_golden("r2_options.npz")


def relerr(a, b, floor=1.0):
    a = np.asarray(a, dtype=float)
    b = np.asarray(b, dtype=float)
    return float(np.max(np.abs(a - b) / np.maximum(np.abs(b), floor)))
