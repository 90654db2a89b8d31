"""pytest configuration: the `gpu` marker and shared helpers."""
import glob
import os
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parents[1]
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))
GOLDEN = ROOT / "tests" / "golden"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def golden_eval_files():
    return sorted(glob.glob(str(GOLDEN / "eval_*.npz")))


def rel_err(a, b):
    """max |a - b| / max |b|  (the parity measure of SURVEY.md 8d)."""
    a = np.asarray(a, dtype=float).ravel()
    b = np.asarray(b, dtype=float).ravel()
    assert a.shape == b.shape, (a.shape, b.shape)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))


@pytest.fixture(scope="session")
def oracle():
    from oracle import Oracle
    return Oracle()
