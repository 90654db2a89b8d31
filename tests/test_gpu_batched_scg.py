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
    from conftest import stop_tolerance
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
        # (a trace that ends in a bit-flat plateau stops where rounding decides: conftest.stop_tolerance)
        n1 = int(scg.stats["MaxIt"])
        assert abs(n1 - int(st["MaxIt"][p])) <= stop_tolerance(scg.stats["fx"][:n1], n1), p
        assert abs(fx[p] - f1) <= 1e-6 * max(abs(f1), 1.0), p
        assert np.abs(Xh[p] - x1).max() <= 1e-5 * max(np.abs(x1).max(), 1.0), p
        assert np.allclose(st["beta"][:n, p], scg.stats["beta"][:n])


@pytest.mark.parametrize("model", ["DW", "OU", "L63", "L96"])
def test_batched_scg_matches_reference_scg_trace(model):
    """Pinned to the REFERENCE directly: the problems of tests/golden/scg_<model>.npz (inputs and
    the fx / beta traces recorded by the unmodified reference's own SCG + VarGP,
    src/numerics/optim_scg.py:75-285), optimised on the device as a batch of three copies -- every
    row must follow the reference trace within 1e-6 on the common prefix (BASELINE.json)."""
    from conftest import GOLDEN, stop_tolerance
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
        assert abs(n_ref - n_new) <= stop_tolerance(g["trace_fx"], n_ref), (p, n_ref, n_new)
        ref, new = g["trace_fx"][:n], st["fx"][:n, p]
        assert np.max(np.abs(new - ref) / np.maximum(np.abs(ref), 1.0)) < 1e-6, p
        assert abs(fx[p] - float(g["fx_final"])) <= 1e-6 * max(abs(float(g["fx_final"])), 1.0), p
        nb = min(n, 10)      # beta is a ratio of differences of nearly equal numbers: compare early iterations only
        assert np.allclose(st["beta"][:nb, p], g["trace_beta"][:nb], rtol=1e-4), p


def test_compacted_launch_equals_masked_launch_bitwise():
    """vgpa_set_active_list (compacted launches of the D = 40 kernels, used by BatchedSCG once part of a
    batch has converged): the listed problems get bit-identical F and gradients to a full evaluation, the
    others are left untouched; the host entry points ignore the list."""
    import torch
    from test_gpu_parity import evaluator_from_golden
    g = _load("eval_L96_rk2")
    B = 7
    rng = np.random.default_rng(5)
    X = np.stack([g["x"] * (1.0 + 0.01 * rng.standard_normal(g["x"].size)) for _ in range(B)])
    sigma = np.stack([g["sigma"] * (0.9 + 0.05 * i) for i in range(B)])
    with evaluator_from_golden(g, B=B, sigma=sigma) as ev:
        F_ref, G_ref = ev.eval(X)
        Xd = torch.from_numpy(X).cuda()
        Fd = torch.full((B,), -1.0, dtype=torch.float64, device="cuda")
        Gd = torch.full_like(Xd, -2.0)
        st = torch.cuda.current_stream().cuda_stream
        pick = np.array([5, 1, 2])                     # any order, any subset
        ev.set_active_list(pick)
        ev.eval_device(Xd.data_ptr(), ev.n_x, Fd.data_ptr(), Gd.data_ptr(), ev.n_x, st)
        ev.sync()
        F, G = Fd.cpu().numpy(), Gd.cpu().numpy()
        rest = np.setdiff1d(np.arange(B), pick)
        assert np.array_equal(F[pick], F_ref[pick]) and np.array_equal(G[pick], G_ref[pick])
        assert np.all(F[rest] == -1.0) and np.all(G[rest] == -2.0)
        F2, G2 = ev.eval(X)                            # host buffers: every problem, list or not
        assert np.array_equal(F2, F_ref) and np.array_equal(G2, G_ref)
        ev.set_active_list(None)
        ev.eval_device(Xd.data_ptr(), ev.n_x, Fd.data_ptr(), Gd.data_ptr(), ev.n_x, st)
        ev.sync()
        assert np.array_equal(Fd.cpu().numpy(), F_ref) and np.array_equal(Gd.cpu().numpy(), G_ref)
    g1 = _load("eval_OU_rk4")
    with evaluator_from_golden(g1, B=3) as ev1:
        with pytest.raises(ValueError):
            ev1.set_active_list([0, 2])                # compacted launches exist for Lorenz-96 only


def test_sharded_batched_scg_single_gpu_sub_batches(tmp_path):
    """ShardedBatchedSCG on one GPU: an ensemble optimised in resident sub-batches gives, problem by problem,
    what one BatchedSCG over the whole ensemble gives; the ensemble file carries the reference's keys."""
    from oracle import prior_kl0
    from vgpa_b200.batched_scg import BatchedSCG, ShardedBatchedSCG
    from vgpa_b200.engine import BatchEvaluator
    from vgpa_b200.simulation import load
    g = _load("eval_L63_heun")
    D, N, B = int(g["D"]), int(g["N"]), 10
    rng = np.random.default_rng([9, D])
    obs_y = np.stack([g["obs_y"] + 0.05 * rng.standard_normal(g["obs_y"].shape) for _ in range(B)])
    E0 = float(prior_kl0(g["m0"], g["s0"], g["mu0"], g["tau0"], False))
    common = dict(model="L63", method="heun", N=N, dt=float(g["dt"]), theta=g["theta"], sigma=g["sigma"], R=g["R"],
                  obs_t=g["obs_t"], m0=g["m0"], s0=g["s0"], E0=E0, dt_model=float(g["dt"]))
    opts = {"max_it": 25, "x_tol": 1.0e-6, "f_tol": 1.0e-8, "display": False}
    make = lambda lo, hi: BatchEvaluator(obs_y=obs_y[lo:hi], B=hi - lo, **common)
    ens = ShardedBatchedSCG(B, make, opts, sub_batch=4)
    res = ens.run(t0=0.0, keep=(0, 7))
    with make(0, B) as ev:
        X0 = ev.initialization(0.0)
        opt = BatchedSCG(ev, opts)
        X, fx = opt(X0)
        Xh = X.cpu().numpy()
    assert np.allclose(res["fx"], fx, rtol=1e-12) and np.array_equal(res["n_it"], opt.stats["MaxIt"])
    assert np.allclose(res["kept"][7], Xh[7], rtol=1e-12, atol=1e-14)
    assert res["sub_batch"] == 4
    # the same sub-batches optimised concurrently (one host thread, CUDA stream and evaluator each): same results
    res2 = ShardedBatchedSCG(B, make, opts, sub_batch=4).run(t0=0.0, keep=(0, 7), concurrent=3)
    assert np.array_equal(res2["fx"], res["fx"]) and np.array_equal(res2["n_it"], res["n_it"])
    assert np.array_equal(res2["kept"][7], res["kept"][7]) and res2["concurrent"] == 3
    import os
    out = ens.save(os.path.join(str(tmp_path), "ens"), N, D)
    z = load(out)
    assert {"fx", "n_it", "f_eval", "problem", "at", "bt"} <= set(z)
    assert z["at"].shape == (2, N, D, D) and z["bt"].shape == (2, N, D) and list(z["problem"]) == [0, 7]
    assert np.array_equal(z["at"][1].ravel(), res["kept"][7][:N * D * D])
