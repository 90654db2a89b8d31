"""
VarGP.initialization on the GPU: the batched cubic-spline starting points
against the x0 the unmodified reference produced (goldens, key x0), against the oracle, and the
device-resident ensemble variant against the host one.  Relative 1e-9 like the rest of the path.
"""
import numpy as np
import pytest

from conftest import golden_eval_files, grad_err, rel_err
from test_gpu_parity import evaluator_from_golden
from oracle import Problem

pytestmark = pytest.mark.gpu
TOL = 1e-9


@pytest.mark.parametrize("path", golden_eval_files(), ids=lambda p: p.split("eval_")[-1][:-4])
def test_initialization_matches_reference(oracle, path):
    g = np.load(path)
    with evaluator_from_golden(g) as ev:
        x0 = ev.initialization(0.0)[0]
    N, D = int(g["N"]), int(g["D"])
    assert grad_err(x0, g["x0"], N, D) < TOL
    assert grad_err(x0, oracle.initialization(Problem.from_golden(g), 0.0), N, D) < TOL


def test_batched_device_initialization_rows():
    """B problems with different observation sets: device rows == host rows == per-problem oracle."""
    import torch
    from oracle import Oracle
    g = np.load(str(next(p for p in golden_eval_files() if "eval_L63_rk2" in p)))
    rng = np.random.default_rng(5)
    B = 5
    obs = np.stack([g["obs_y"] + 0.3 * rng.standard_normal(g["obs_y"].shape) for _ in range(B)])
    with evaluator_from_golden(g, B=B, obs_y=obs) as ev:
        Xh = ev.initialization(0.0)
        Xd = torch.empty((B, ev.n_x), dtype=torch.float64, device="cuda")
        ev.initialization_device(Xd.data_ptr(), ev.n_x, 0.0, torch.cuda.current_stream().cuda_stream)
        assert np.array_equal(Xd.cpu().numpy(), Xh)
    orc = Oracle()
    for p in range(B):
        prob = Problem.from_golden(g)
        prob.obs_y = obs[p]
        assert rel_err(Xh[p], orc.initialization(prob, 0.0)) < TOL


def test_initialization_rejects_observation_at_the_ends():
    """The reference's CubicSpline raises ValueError when a knot repeats (observation at index 0 or N-1)."""
    g = np.load(str(next(p for p in golden_eval_files() if "eval_OU_rk2" in p)))
    ot = g["obs_t"].copy()
    ot[0] = 0
    with evaluator_from_golden(g, obs_t=ot) as ev:
        with pytest.raises(ValueError):
            ev.initialization(0.0)
