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


def grad_err(a, b, N, D):
    """The parity measure for a flat vector laid out like x = [A (N,D,D) | b (N,D)] -- the gradient
    [dL/dA | dL/db], or x0 itself -- taken PER BLOCK: the larger of the two blocks' rel_err.  (Over the
    concatenated vector the block with the smaller magnitude -- dL/db is 10-25 times smaller than dL/dA
    in the L63 / L96 fixtures -- would be held to a correspondingly looser bound.)"""
    a = np.asarray(a, dtype=float).ravel()
    b = np.asarray(b, dtype=float).ravel()
    na = int(N) * int(D) * int(D)
    assert a.size == b.size == na + int(N) * int(D), (a.size, b.size, N, D)
    return max(rel_err(a[:na], b[:na]), rel_err(a[na:], b[na:]))


def key_err(key, got, ref, N, D):
    """rel_err of one named output; `grad` per block."""
    return grad_err(got, ref, N, D) if key == "grad" else rel_err(got, ref)


@pytest.fixture(scope="session")
def oracle():
    from oracle import Oracle
    return Oracle()


def stop_tolerance(trace_fx, n_ref):
    """How far an SCG run's stopping iteration may be from the reference's.  A few iterations either way
    because of the |f_new - f_old| <= 1e-8 test -- and where the reference's trace ends in a plateau (OU: fx
    unchanged TO THE LAST BIT over the final iterations, SCG only raising lambda until the step falls under
    x_tol) the exit is decided by rounding: allow a quarter of the plateau, at most 5 iterations."""
    tr = np.asarray(trace_fx, dtype=float)
    plateau = int(np.argmax(np.abs(tr[::-1] - tr[-1]) > 1e-13 * max(abs(tr[-1]), 1.0)))
    return max(2, n_ref // 50, min(plateau // 4, 5))
