import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def rel_l2(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    den = np.linalg.norm(b.ravel())
    return float(np.linalg.norm((a - b).ravel()) / den) if den > 0 else float(np.abs(a).max())


def max_norm_err(a, b):
    """max |a-b| / max |b|  (SURVEY.md 8c parity metric)."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    den = np.abs(b).max()
    return float(np.abs(a - b).max() / den) if den > 0 else float(np.abs(a).max())


def _load(name):
    z = np.load(os.path.join(GOLDEN, name))
    cases = {}
    for key in z.files:
        case, tensor = key.rsplit(".", 1)
        cases.setdefault(case, {})[tensor] = z[key]
    return cases


@pytest.fixture(scope="session")
def corr_golden():
    return _load("corr_golden.npz")


@pytest.fixture(scope="session")
def warp_golden():
    return _load("warp_golden.npz")


@pytest.fixture(scope="session")
def occ_golden():
    return _load("occ_golden.npz")
