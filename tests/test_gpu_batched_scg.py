"""
Device-resident batched SCG (vgpa_b200/batched_scg.py) against
the single-problem host SCG (vgpa_b200/scg.py, itself pinned to the reference's SCG):
for every problem of a batch the fx trace must agree within 1e-6 on the common prefix
and the final free energy within 1e-6.
"""
import numpy as np
import pytest

from conftest import golden_eval_files

pytestmark = pytest.mark.gpu


def _load(name):
    return np.load(str(next(p for p in golden_eval_files() if name in p)))


@pytest.mark.parametrize("name,max_it", [("eval_OU_rk4", 60), ("eval_DW_euler", 60), ("eval_L63_heun", 40),
                                          ("eval_L96_rk2", 12)])
def test_batched_scg_follows_single_problem_scg(name, max_it):
    from oracle import prior_kl0
    from vgpa_b200.batched_scg import BatchedSCG
    from vgpa_b200.engine import BatchEvaluator
    from vgpa_b200.scg import SCG
    g = _load(name)
    D, N, B = int(g["D"]), int(g["N"]), 4
    rng = np.random.default_rng([3, D])
    obs_y = np.stack([g["obs_y"] + 0.05 * rng.standard_normal(g["obs_y"].shape) for _ in range(B)])
    sigma = np.stack([g["sigma"] * (0.9 + 0.1 * i) for i in range(B)])
    E0 = float(prior_kl0(g["m0"], g["s0"], g["mu0"], g["tau0"], D == 1))
    common = dict(model=str(g["model"]), method=str(g["method"]), N=N, dt=float(g["dt"]), theta=g["theta"],
                  R=g["R"], obs_t=g["obs_t"], m0=g["m0"], s0=g["s0"], E0=E0, dt_model=float(g["dt"]))
    opts = {"max_it": max_it, "x_tol": 1.0e-6, "f_tol": 1.0e-8, "display": False}
    X0 = np.stack([g["x0"]] * B)
    with BatchEvaluator(sigma=sigma, obs_y=obs_y, B=B, **common) as ev:
        opt = BatchedSCG(ev, opts)
        X, fx = opt(X0)
        Xh = X.cpu().numpy()
        st = opt.stats
        # the batch needs 2 evaluations per iteration, whatever B is
        assert st["evaluations"] <= 2 * max_it + 1
    for p in range(B):
        with BatchEvaluator(sigma=sigma[p], obs_y=obs_y[p], B=1, **common) as ev1:
            f = lambda x: float(ev1.eval(x, want_grad=False)[0][0])
            df = lambda x, eval_fun=False: ev1.eval(x)[1][0].copy()
            scg = SCG(f, df, opts)
            x1, f1 = scg(g["x0"].copy())
        n = min(int(scg.stats["MaxIt"]), int(st["MaxIt"][p]))
        ref, new = scg.stats["fx"][:n], st["fx"][:n, p]
        assert np.max(np.abs(new - ref) / np.maximum(np.abs(ref), 1.0)) < 1e-6, p
        assert abs(int(scg.stats["MaxIt"]) - int(st["MaxIt"][p])) <= 2, p
        assert abs(fx[p] - f1) <= 1e-6 * max(abs(f1), 1.0), p
        assert np.abs(Xh[p] - x1).max() <= 1e-5 * max(np.abs(x1).max(), 1.0), p
        assert np.allclose(st["beta"][:n, p], scg.stats["beta"][:n])


@pytest.mark.parametrize("model", ["DW", "OU", "L63", "L96"])
def test_batched_scg_matches_reference_scg_trace(model):
    """Pinned to the REFERENCE directly: the problems of tests/golden/scg_<model>.npz (inputs and
    the fx / beta traces recorded by the unmodified reference's own SCG + VarGP,
    src/numerics/optim_scg.py:75-285), optimised on the device as a batch of three copies -- every
    row must follow the reference trace within 1e-6 on the common prefix (BASELINE.json)."""
    from conftest import GOLDEN
    from test_gpu_parity import evaluator_from_golden
    from vgpa_b200.batched_scg import BatchedSCG
    g = np.load(GOLDEN / f"scg_{model}.npz")
    B = 3
    opts = {"max_it": int(g["max_it"]), "x_tol": 1.0e-6, "f_tol": 1.0e-8, "display": False}
    with evaluator_from_golden(g, B=B) as ev:
        opt = BatchedSCG(ev, opts)
        X, fx = opt(np.stack([g["x"]] * B))
        st = opt.stats
    n_ref = int(g["n_it"])
    for p in range(B):
        n_new = int(st["MaxIt"][p])
        n = min(n_ref, n_new)
        assert abs(n_ref - n_new) <= max(2, n_ref // 50), (p, n_ref, n_new)
        ref, new = g["trace_fx"][:n], st["fx"][:n, p]
        assert np.max(np.abs(new - ref) / np.maximum(np.abs(ref), 1.0)) < 1e-6, p
        assert abs(fx[p] - float(g["fx_final"])) <= 1e-6 * max(abs(float(g["fx_final"])), 1.0), p
        nb = min(n, 10)      # beta is a ratio of differences of nearly equal numbers: compare early iterations only
        assert np.allclose(st["beta"][:nb, p], g["trace_beta"][:nb], rtol=1e-4), p
