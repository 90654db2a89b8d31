"""
Hyper-parameter gradients through the C ABI: dEsde_dtheta, dEsde_dsigma of
model.energy and dEobs_dr of GaussianLikelihood.gradients, against the unmodified reference
(tests/golden/hyper_*.npz, written by make_golden_hyper.py) and against the oracle, to the same
relative 1e-9 as the rest of the path.  Also through the reference-shaped classes.
"""
from pathlib import Path

import numpy as np
import pytest

from conftest import rel_err
from oracle import Problem

pytestmark = pytest.mark.gpu
GOLD = Path(__file__).resolve().parent / "golden"
TOL = 1e-9


def _load(model):
    g = np.load(GOLD / f"eval_{model}_rk2.npz")
    h = np.load(GOLD / f"hyper_{model}.npz")
    N, D = int(g["N"]), int(g["D"])
    x = g["x"]
    A = x[:N * D * D].reshape((N,) if D == 1 else (N, D, D))
    b = x[N * D * D:].reshape((N,) if D == 1 else (N, D))
    return g, h, A, b


@pytest.mark.parametrize("model", ["DW", "OU", "L63", "L96"])
def test_energy_hyper_gradients(oracle, model):
    import vgpa_b200.engine as eng
    g, h, A, b = _load(model)
    prob = Problem.from_golden(g)
    out = eng.model_energy(model, g["theta"], g["sigma"], A, b, g["mt"], g["st"], prob.dt_model, hyper=True)
    Esde, dth, dsig = out[0], out[5], out[6]
    assert abs(Esde - float(g["Esde"])) <= TOL * abs(float(g["Esde"]))
    assert rel_err(dth, h["dEsde_dtheta"]) < TOL
    assert rel_err(dsig, h["dEsde_dsigma"]) < TOL
    o_th, o_sig = oracle.energy_hyper(prob, g["x"], g["mt"], g["st"])
    assert rel_err(dth, o_th) < TOL and rel_err(dsig, o_sig) < TOL
    # without the request the five-output form is unchanged
    assert len(eng.model_energy(model, g["theta"], g["sigma"], A, b, g["mt"], g["st"], prob.dt_model)) == 5


@pytest.mark.parametrize("model", ["DW", "OU", "L63", "L96"])
def test_likelihood_dEobs_dr(oracle, model):
    from vgpa_b200.likelihood import GaussianLikelihood
    g, h, _, _ = _load(model)
    D = int(g["D"])
    obs_y = g["obs_y"].reshape(-1) if D == 1 else g["obs_y"]
    R = float(g["R"][0]) if D == 1 else np.diag(g["R"])
    lik = GaussianLikelihood(obs_y, list(g["obs_t"]), R, None, D == 1)
    jm, js, dr = lik.gradients(g["mt"], g["st"])
    assert dr.shape == h["dEobs_dr"].shape
    if np.abs(h["dEobs_dr"]).max() > 0:
        assert rel_err(dr, h["dEobs_dr"]) < TOL
        assert rel_err(dr, oracle.eobs_dr(Problem.from_golden(g), g["mt"], g["st"])) < TOL
    else:
        assert not dr.any()


@pytest.mark.parametrize("model", ["OU", "L96"])
def test_process_energy_returns_reference_structure(model):
    """StochasticProcess.energy -> Esde, (Ef, Edf), (dEsde_dm, dEsde_ds, dEsde_dtheta, dEsde_dsigma)."""
    from vgpa_b200.simulation import dynamical_systems
    g, h, A, b = _load(model)
    D, N = int(g["D"]), int(g["N"])
    sigma = float(g["sigma"][0]) if D == 1 else list(g["sigma"])
    proc = dynamical_systems[model](sigma, float(g["theta"][0]), 1234)
    proc.time_window = float(g["dt"]) * np.arange(N)
    Esde, (Ef, Edf), (dm, ds, dth, dsig) = proc.energy(A, b, g["mt"], g["st"], list(g["obs_t"]))
    assert abs(Esde - float(g["Esde"])) <= TOL * abs(float(g["Esde"]))
    assert rel_err(Ef, g["Efx"]) < TOL and rel_err(dm, g["dEsde_dm"]) < TOL and rel_err(ds, g["dEsde_ds"]) < TOL
    assert rel_err(dth, h["dEsde_dtheta"]) < TOL and rel_err(dsig, h["dEsde_dsigma"]) < TOL
