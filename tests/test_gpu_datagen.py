"""
Batched sample paths and observations on the GPU (csrc/datagen.cu through
vgpa_make_trajectory / vgpa_collect_obs and their _device variants) against the unmodified
reference's outputs (tests/golden/datagen_*.npz) and the oracle.  The kernels perform the
reference's IEEE operations in its order, so the comparison is bit-exact, with 1e-12 relative as
the stated bound (numpy's scalar x ** 2 in the Double-Well drift goes through libm's pow).
"""
import numpy as np
import pytest

from conftest import GOLDEN, rel_err
from test_datagen_cpu import MODELS, SEED, theta_of, x_init_of

pytestmark = pytest.mark.gpu
TOL = 1e-12


@pytest.mark.parametrize("model", MODELS)
def test_single_path_matches_reference(model):
    from vgpa_b200.engine import collect_observations, make_trajectories
    g = np.load(GOLDEN / f"datagen_{model}.npz")
    path = make_trajectories(model, int(g["N"]), float(g["dt"]), theta_of(g), g["sigma"], g["z"][None],
                             x_init_of(g))[0]
    assert path.shape == g["path"].shape
    assert rel_err(path, g["path"]) < TOL
    if model != "DW":
        assert np.array_equal(path, g["path"])
    obs = collect_observations(g["path"], g["obs_t"], g["R"], g["xi"][None])[0]
    assert np.array_equal(obs, g["obs_y"])


@pytest.mark.parametrize("model", MODELS)
def test_ensemble_rows_match_oracle(oracle, model):
    """B paths with their own noise level, drift parameter and draws; rows == per-path oracle, and
    observation sets of one shared path == per-set oracle."""
    from vgpa_b200.engine import collect_observations, make_trajectories
    g = np.load(GOLDEN / f"datagen_{model}.npz")
    D, N, dt = int(g["D"]), int(g["N"]), float(g["dt"])
    rng = np.random.default_rng(17)
    B = 7
    Z = rng.standard_normal((B, N) if D == 1 else (B, D, N))
    th = np.stack([theta_of(g) * (1.0 + 0.05 * rng.uniform(-1, 1)) for _ in range(B)])
    sg = np.stack([g["sigma"] * 2.0 ** rng.uniform(-1, 1) for _ in range(B)])
    xi0 = None if D > 1 else rng.standard_normal((B, 1))
    paths = make_trajectories(model, N, dt, th, sg, Z, xi0)
    for p in range(B):
        ref = oracle.make_trajectory(model, N, dt, th[p], sg[p], Z[p], None if xi0 is None else xi0[p])
        assert rel_err(paths[p], ref) < TOL
    M = g["obs_t"].size
    XI = rng.standard_normal((B, M) if D == 1 else (B, D, M))
    R = np.stack([g["R"] * 2.0 ** rng.uniform(-1, 1) for _ in range(B)])
    sets = collect_observations(g["path"], g["obs_t"], R, XI)          # one path, B observation sets
    own = collect_observations(paths, g["obs_t"], R, XI)               # each path its own set
    for p in range(B):
        assert np.array_equal(sets[p], oracle.collect_obs(g["path"], g["obs_t"], R[p], XI[p]))
        assert np.array_equal(own[p], oracle.collect_obs(paths[p], g["obs_t"], R[p], XI[p]))


def test_device_resident_ensemble_equals_host_call():
    """The _device entry points on torch buffers (paths and observation sets never leave HBM)."""
    import torch
    from vgpa_b200._lib import MODELS as MID, lib, raise_for
    from vgpa_b200.engine import collect_observations, make_trajectories
    g = np.load(GOLDEN / "datagen_L96.npz")
    D, N, M, dt = 40, int(g["N"]), g["obs_t"].size, float(g["dt"])
    rng = np.random.default_rng(3)
    B = 5
    Z, XI = rng.standard_normal((B, D, N)), rng.standard_normal((B, D, M))
    dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    th, sg, R, ot, dZ, dXI = dev(g["theta"]), dev(g["sigma"]), dev(g["R"]), dev(g["obs_t"]), dev(Z), dev(XI)
    path = torch.empty((B, N, D), dtype=torch.float64, device="cuda")
    obs = torch.empty((B, M, D), dtype=torch.float64, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    raise_for(lib.vgpa_make_trajectory_device(0, MID["L96"], N, B, dt, th.data_ptr(), 0, sg.data_ptr(), 0, None, 0,
                                              dZ.data_ptr(), N * D, path.data_ptr(), N * D, st))
    raise_for(lib.vgpa_collect_obs_device(0, D, N, M, B, ot.data_ptr(), R.data_ptr(), 0, path.data_ptr(), N * D,
                                          dXI.data_ptr(), M * D, obs.data_ptr(), M * D, st))
    torch.cuda.synchronize()
    hp = make_trajectories("L96", N, dt, g["theta"], g["sigma"], Z)
    assert np.array_equal(path.cpu().numpy(), hp)
    assert np.array_equal(obs.cpu().numpy(), collect_observations(hp, g["obs_t"], g["R"], XI))


@pytest.mark.parametrize("model", MODELS)
def test_mirror_with_device_option_equals_host_path(model):
    """<Model>(sigma, theta, seed).make_trajectory(..., device=0) / collect_obs(..., device=0)."""
    from vgpa_b200.dynamics import dynamical_systems
    g = np.load(GOLDEN / f"datagen_{model}.npz")
    D = int(g["D"])
    sigma = float(g["sigma"][0]) if D == 1 else list(g["sigma"])
    theta = float(g["theta"][0]) if model != "L63" else list(g["theta"])
    proc = dynamical_systems[model](sigma, theta, SEED)
    proc.make_trajectory(0.0, float(g["tf"]), float(g["dt"]), device=0)
    assert rel_err(proc.sample_path, g["path"]) < TOL
    density = {"DW": 2, "OU": 2, "L63": 5, "L96": 8}[model]
    obs_t, obs_y, _ = proc.collect_obs(density, float(g["R"][0]) if D == 1 else g["R"], device=0)
    assert np.array_equal(np.asarray(obs_t), g["obs_t"])
    assert rel_err(obs_y, g["obs_y"]) < TOL


def test_argument_errors():
    from vgpa_b200.engine import make_trajectories
    with pytest.raises(ValueError):
        make_trajectories("DW", 10, 0.01, [1.0], [0.8], np.zeros((1, 10)))          # x_init missing
    with pytest.raises(ValueError):
        make_trajectories("L63", 10, -0.01, [10.0, 28.0, 2.7], [1.0] * 3, np.zeros((1, 3, 10)))
    with pytest.raises(ValueError):
        make_trajectories("XX", 10, 0.01, [1.0], [0.8], np.zeros((1, 10)), [0.0])


@pytest.mark.parametrize("model", ["L63", "L96"])
def test_path_continues_from_a_given_state(model):
    """x_init for the n-D models (no burn-in), per path: restarted from the states at two different
    indices with the remaining draws, the GPU gives the tails of the reference's path bit for bit."""
    from vgpa_b200.engine import make_trajectories
    g = np.load(GOLDEN / f"datagen_{model}.npz")
    N, D = int(g["N"]), int(g["D"])
    k1, k2 = 37, 90
    n = N - k2
    Z = np.stack([g["z"][:, k1:k1 + n], g["z"][:, k2:k2 + n]])
    X0 = np.stack([g["path"][k1], g["path"][k2]])
    tails = make_trajectories(model, n, float(g["dt"]), theta_of(g), g["sigma"], Z, X0)
    assert np.array_equal(tails[0], g["path"][k1:k1 + n])
    assert np.array_equal(tails[1], g["path"][k2:k2 + n])
