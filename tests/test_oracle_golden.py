"""
Pins the CPU oracle (oracle/vgpa_oracle.c) to the reference: every array the
reference's hot path produces, recorded by tests/golden/make_golden.py from the
unmodified reference, must be reproduced to 1e-11 (observed: <= 4e-13).
"""
import numpy as np
import pytest

from conftest import golden_eval_files, grad_err, key_err, rel_err
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
        assert key_err(k, out[k], g[k], prob.N, prob.D) < TOL, k


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
    assert grad_err(x0, g["x0"], prob.N, prob.D) < TOL


def load_mid(method):
    """mid_L96_<method>.npz (L96, N = 101, M = 8 observations); the evaluation point is stored once."""
    from pathlib import Path
    gold = Path(__file__).resolve().parent / "golden"
    g = dict(np.load(gold / f"mid_L96_{method}.npz"))
    if "x" not in g:
        g["x"] = np.load(gold / "mid_L96_rk2.npz")["x"]
    return g


MID_3D = ("st", "psit", "Edf", "dEsde_ds")


def mid_errors(out, g):
    """Parity errors of every output of a mid_L96 fixture ((N, D, D) arrays at the stored indices)."""
    N, D, t_idx = int(g["N"]), int(g["D"]), g["t_idx"]
    errs = {}
    for k in KEYS:
        got = np.asarray(out[k])
        if k in MID_3D:
            got = got.reshape(N, D, D)[t_idx]
        errs[k] = key_err(k, got, g[k], N, D)
    return errs


@pytest.mark.parametrize("method", ["rk2", "rk4"])
def test_oracle_l96_eight_observations_all_intermediates(oracle, method):
    """L96 with M = 8 observations: the jump logic of the backward sweep at every observation index and
    the observation-ordinal quirk at the level of lamt / psit, not only through F."""
    g = load_mid(method)
    assert int(g["N"]) == 101 and g["obs_t"].size == 8
    prob = Problem.from_golden(g)
    out = oracle.eval(prob, g["x"], full=True)
    bad = {k: e for k, e in mid_errors(out, g).items() if e >= TOL}
    assert not bad, bad
