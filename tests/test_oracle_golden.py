"""
Pins the CPU oracle (oracle/vgpa_oracle.c) to the reference: every array the
reference's hot path produces, recorded by tests/golden/make_golden.py from the
unmodified reference, must be reproduced to 1e-11 (observed: <= 4e-13).
"""
import numpy as np
import pytest

from conftest import golden_eval_files, rel_err
from oracle import Problem

KEYS = ("F", "E0", "Esde", "Eobs", "grad", "mt", "st", "lamt", "psit", "Efx", "Edf",
        "dEsde_dm", "dEsde_ds")
TOL = 1e-11


@pytest.mark.parametrize("path", golden_eval_files(), ids=lambda p: p.split("eval_")[-1][:-4])
def test_oracle_reproduces_reference(oracle, path):
    g = np.load(path)
    prob = Problem.from_golden(g)
    out = oracle.eval(prob, g["x"], full=True)
    for k in KEYS:
        assert rel_err(out[k], g[k]) < TOL, k


@pytest.mark.parametrize("model", ["DW", "OU", "L63", "L96"])
def test_oracle_known_answer_at_x0(oracle, model):
    """F(x0) and |grad F(x0)| at the reference's own initialisation."""
    g = np.load(str(next(p for p in golden_eval_files() if f"eval_{model}_rk2" in p)))
    prob = Problem.from_golden(g)
    F, grad = oracle.eval(prob, g["x0"])
    assert abs(F - float(g["F_x0"])) <= 1e-11 * abs(float(g["F_x0"]))
    assert abs(np.linalg.norm(grad) - float(g["gnorm_x0"])) <= 1e-10 * float(g["gnorm_x0"])


def test_oracle_not_positive_definite_raises(oracle):
    """A covariance that loses positive definiteness is a LinAlgError in the
    reference (utilities.py:211 via variational.py:380)."""
    g = np.load(str(next(p for p in golden_eval_files() if "eval_L96_euler" in p)))
    prob = Problem.from_golden(g)
    x = g["x"].copy()
    N, D = prob.N, prob.D
    x[:N * D * D] = -200.0 * np.tile(np.eye(D).ravel(), N)   # explosive A: S(t) blows up / goes indefinite
    x[:N * D * D] += 50.0 * np.random.default_rng(0).standard_normal(N * D * D)
    with pytest.raises(np.linalg.LinAlgError):
        oracle.eval(prob, x)


def test_oracle_batch_matches_single(oracle):
    g = np.load(str(next(p for p in golden_eval_files() if "eval_L63_heun" in p)))
    prob = Problem.from_golden(g)
    rng = np.random.default_rng(3)
    X = np.stack([g["x"] * (1 + 0.01 * rng.standard_normal(g["x"].size)) for _ in range(4)])
    F, G = oracle.eval_batch([prob] * 4, X, threads=2)
    for i in range(4):
        Fi, Gi = oracle.eval(prob, X[i])
        assert Fi == F[i] and np.array_equal(Gi, G[i])


@pytest.mark.parametrize("model", ["DW", "OU", "L63", "L96"])
def test_oracle_hyper_gradients(oracle, model):
    """dEsde_dtheta, dEsde_dsigma (model.energy) and dEobs_dr (GaussianLikelihood.gradients)
    against the unmodified reference (tests/golden/make_golden_hyper.py)."""
    from pathlib import Path
    gold = Path(__file__).resolve().parent / "golden"
    g = np.load(gold / f"eval_{model}_rk2.npz")
    h = np.load(gold / f"hyper_{model}.npz")
    prob = Problem.from_golden(g)
    dth, dsig = oracle.energy_hyper(prob, g["x"], g["mt"], g["st"])
    assert rel_err(dth, h["dEsde_dtheta"]) < TOL
    assert rel_err(dsig, h["dEsde_dsigma"]) < TOL
    dr = oracle.eobs_dr(prob, g["mt"], g["st"])
    assert dr.shape == h["dEobs_dr"].shape
    assert rel_err(dr, h["dEobs_dr"]) < TOL if np.abs(h["dEobs_dr"]).max() > 0 else not dr.any()


@pytest.mark.parametrize("path", golden_eval_files(), ids=lambda p: p.split("eval_")[-1][:-4])
def test_oracle_initialization_reproduces_reference(oracle, path):
    """VarGP.initialization (cubic splines through the observations) against the x0 the unmodified
    reference produced (goldens: key x0)."""
    g = np.load(path)
    prob = Problem.from_golden(g)
    x0 = oracle.initialization(prob, 0.0)
    assert rel_err(x0, g["x0"]) < TOL
