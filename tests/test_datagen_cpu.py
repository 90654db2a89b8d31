"""
Data generation on the CPU side: the oracle's restatement of
make_trajectory / collect_obs and the host mirror's numpy path against what the unmodified
reference produced (tests/golden/datagen_*.npz, which also hold the draws it consumed).
Bit-exact: both follow the reference's sequence of IEEE operations.
"""
import numpy as np
import pytest

from conftest import GOLDEN

MODELS = ["DW", "OU", "L63", "L96"]
SEED = 31415926535


def x_init_of(g):
    """State at t0 as the reference builds it from its first draws (DW) or its constant (OU)."""
    model = str(g["model"])
    if model == "DW":            # double_well.py:145-151
        s, dt, th = float(g["sigma"][0]), float(g["dt"]), float(g["theta"][0])
        x0 = +th if float(g["u_start"]) > 0.5 else -th
        return np.array([x0 + np.sqrt(0.5 * s * dt) * float(g["n_start"])])
    if model == "OU":
        return np.array([0.0])
    return None


def theta_of(g):
    return np.array([float(g["theta"][0]), 0.0]) if str(g["model"]) == "OU" else g["theta"]


@pytest.mark.parametrize("model", MODELS)
def test_oracle_paths_and_observations_match_reference(oracle, model):
    g = np.load(GOLDEN / f"datagen_{model}.npz")
    path = oracle.make_trajectory(model, int(g["N"]), float(g["dt"]), theta_of(g), g["sigma"], g["z"], x_init_of(g))
    assert np.array_equal(path, g["path"])
    obs = oracle.collect_obs(g["path"], g["obs_t"], g["R"], g["xi"])
    assert np.array_equal(obs, g["obs_y"])


@pytest.mark.parametrize("model", MODELS)
def test_host_mirror_reproduces_reference_stream(model):
    """Same seed -> same path, observation indices and observations as the reference (this also
    checks that the stored draws are what numpy's Generator yields here)."""
    from vgpa_b200.dynamics import dynamical_systems
    g = np.load(GOLDEN / f"datagen_{model}.npz")
    D = int(g["D"])
    sigma = float(g["sigma"][0]) if D == 1 else list(g["sigma"])
    theta = float(g["theta"][0]) if model != "L63" else list(g["theta"])
    proc = dynamical_systems[model](sigma, theta, SEED)
    proc.make_trajectory(0.0, float(g["tf"]), float(g["dt"]))
    assert np.array_equal(proc.sample_path, g["path"])
    density = {"DW": 2, "OU": 2, "L63": 5, "L96": 8}[model]
    obs_t, obs_y, _ = proc.collect_obs(density, float(g["R"][0]) if D == 1 else g["R"])
    assert np.array_equal(np.asarray(obs_t), g["obs_t"])
    assert np.array_equal(obs_y, g["obs_y"])


@pytest.mark.parametrize("model", ["L63", "L96"])
def test_oracle_path_continues_from_a_given_state(oracle, model):
    """x_init for the n-D models (no burn-in): a path restarted from its own state at index k with the
    remaining draws is the tail of the original path, bit for bit."""
    g = np.load(GOLDEN / f"datagen_{model}.npz")
    N, k = int(g["N"]), 37
    tail = oracle.make_trajectory(model, N - k, float(g["dt"]), theta_of(g), g["sigma"],
                                  np.ascontiguousarray(g["z"][:, k:]), g["path"][k])
    assert np.array_equal(tail, g["path"][k:])
