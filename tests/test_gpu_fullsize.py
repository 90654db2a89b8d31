"""
Parity at BASELINE.json's FULL sizes (-m gpu): the ensemble construction of configs[4] (L96 D=40,
N=1001, RK2: observation sets x perturbed starts x system-noise values) row by row against the
oracle at relative 1e-9, and size-independent properties of the path that need no oracle at all
(NOT among them: "grad is the derivative of F".  The reference's gradient is the discretised
continuous adjoint -- optimise-then-discretise, variational.py:202-289 -- and differs from the
derivative of the discrete F at O(dt); central differences disagree with it by ~15 % here, for the
reference exactly as for this path):
  * F = E0 + Esde + Eobs, and E0 enters additively (a per-problem constant);
  * a problem's result does not depend on its position in the batch or on its neighbours (bitwise);
  * the backward sweep is linear in its driving terms (dE/dm, dE/dS, jumps).
Also configs[2] at full size (L63, N=2002, RK2) and configs[1] (OU, N=1001, RK4) against the oracle.
"""
import numpy as np
import pytest

from conftest import golden_eval_files, grad_err, rel_err
from oracle import Problem
from test_gpu_parity import evaluator_from_golden

pytestmark = pytest.mark.gpu
TOL = 1e-9


@pytest.fixture(scope="module")
def l96_family():
    import bench
    fam = bench.l96_problem_family(0)
    iset, arr = bench.shard_arrays(fam, bench.N_STARTS * bench.N_NOISE * 2)   # two observation sets
    return bench, fam, iset, arr


def _pick(bench, fam, iset, arr, idx, seed=7):
    """Problems `idx` of the shard with their C5 starting points x = x0 (1 + 0.02 u)."""
    rng = np.random.default_rng(seed)
    X = np.stack([fam["sets"][iset[p]]["x0"] * (1.0 + 0.02 * rng.uniform(-1, 1, bench.N_X)) for p in idx])
    kw = dict(sigma=arr["sigma"][idx], obs_y=arr["obs_y"][idx], m0=arr["m0"][idx], E0=arr["E0"][idx])
    return X, kw


def _evaluator(bench, fam, kw, B):
    from vgpa_b200.engine import BatchEvaluator
    return BatchEvaluator("L96", "rk2", bench.N_GRID, bench.DT, [8.0], kw["sigma"], np.ones(bench.D), fam["obs_t"],
                          kw["obs_y"], kw["m0"], fam["s0"], kw["E0"], B=B, dt_model=fam["dt_model"])


def test_l96_ensemble_rows_match_oracle_at_full_size(oracle, l96_family):
    bench, fam, iset, arr = l96_family
    idx = np.array([0, 7, 15, 300, 512 + 9, 1023])          # both observation sets, extreme noise values
    X, kw = _pick(bench, fam, iset, arr, idx)
    with _evaluator(bench, fam, kw, len(idx)) as ev:
        F, G = ev.eval(X)
    for r, p in enumerate(idx):
        prob = Problem(model="L96", method="rk2", D=bench.D, N=bench.N_GRID, dt=bench.DT, theta=[8.0],
                       sigma=kw["sigma"][r], R=np.ones(bench.D), obs_t=fam["obs_t"], obs_y=kw["obs_y"][r],
                       m0=kw["m0"][r], s0=fam["s0"], E0=float(kw["E0"][r]), dt_model=fam["dt_model"])
        Fo, Go = oracle.eval(prob, X[r])
        assert abs(F[r] - Fo) <= TOL * abs(Fo), (p, F[r], Fo)
        assert grad_err(G[r], Go, prob.N, prob.D) < TOL, p


def test_l96_batch_position_independence_bitwise(l96_family):
    bench, fam, iset, arr = l96_family
    idx = np.array([2, 40, 41, 600])
    X, kw = _pick(bench, fam, iset, arr, idx, seed=13)
    with _evaluator(bench, fam, kw, 4) as ev:
        F, G = ev.eval(X)
    perm = np.array([3, 0, 2, 1])
    kwp = {k: v[perm] for k, v in kw.items()}
    with _evaluator(bench, fam, kwp, 4) as ev:
        Fp, Gp = ev.eval(X[perm])
    assert np.array_equal(Fp, F[perm]) and np.array_equal(Gp, G[perm])
    kw1 = {k: v[1:2] for k, v in kw.items()}
    with _evaluator(bench, fam, kw1, 1) as ev:
        F1, G1 = ev.eval(X[1])
    assert F1[0] == F[1] and np.array_equal(G1[0], G[1])


def test_free_energy_decomposition_and_additive_E0():
    g = np.load(str(next(p for p in golden_eval_files() if "eval_L96_rk2" in p)))
    with evaluator_from_golden(g) as ev:
        full = ev.eval_full(g["x"])
        F0, G0 = ev.eval(g["x"])
    assert abs(full["F"] - (full["E0"] + full["Esde"] + full["Eobs"])) <= 1e-13 * abs(full["F"])
    with evaluator_from_golden(g, E0=float(full["E0"]) + 123.5) as ev:
        F1, G1 = ev.eval(g["x"])
    assert abs((F1[0] - F0[0]) - 123.5) <= 1e-12 * abs(F0[0]) and np.array_equal(G1, G0)


@pytest.mark.parametrize("name", ["eval_L96_rk2", "eval_L63_rk4", "eval_OU_heun"])
def test_backward_sweep_is_linear_in_its_driving_terms(name):
    import vgpa_b200.engine as eng
    g = np.load(str(next(p for p in golden_eval_files() if name in p)))
    N, D = int(g["N"]), int(g["D"])
    A = g["x"][:N * D * D].reshape((N,) if D == 1 else (N, D, D))
    rng = np.random.default_rng(2)
    shp_v, shp_m = ((N,), (N,)) if D == 1 else ((N, D), (N, D, D))

    def sym(a):
        return a if D == 1 else 0.5 * (a + np.swapaxes(a, 1, 2))
    terms = [(rng.standard_normal(shp_v), sym(rng.standard_normal(shp_m)), rng.standard_normal(shp_v),
              sym(rng.standard_normal(shp_m))) for _ in range(2)]
    sol = [eng.solve_bwd(str(g["method"]), A, *t, float(g["dt"])) for t in terms]
    a, b = 0.7, -1.9
    mix = tuple(a * u + b * v for u, v in zip(*terms))
    lam, psi = eng.solve_bwd(str(g["method"]), A, *mix, float(g["dt"]))
    assert rel_err(lam, a * sol[0][0] + b * sol[1][0]) < 1e-11
    assert rel_err(psi, a * sol[0][1] + b * sol[1][1]) < 1e-11


@pytest.mark.parametrize("model,method,N", [("L63", "rk2", 2002), ("OU", "rk4", 1001)])
def test_small_models_at_full_size_match_oracle(oracle, model, method, N):
    """configs[2] (L63, T=2000, RK2) and configs[1] (OU, T=1000, RK4): a batch of distinct observation
    sets and starting points at the full grid size, row by row against the oracle."""
    from vgpa_b200.engine import BatchEvaluator
    D = 3 if model == "L63" else 1
    rng = np.random.default_rng(17)
    M = 100 if model == "L63" else 20
    obs_t = np.linspace(0, N, M + 2, dtype=int)[1:-1]
    theta = [10.0, 28.0, 2.6667] if model == "L63" else [2.0]
    sig = np.full(D, 10.0 if model == "L63" else 0.8)
    R = np.full(D, 2.0 if model == "L63" else 0.04)
    B = 6
    obs_y = rng.standard_normal((B, M, D)) * (3.0 if model == "L63" else 0.5)
    m0 = rng.standard_normal((B, D))
    s0 = 0.25 * np.eye(D)
    a_diag = 0.5 * sig / 0.25
    x1 = np.concatenate([np.tile(np.diag(a_diag).ravel(), N), np.zeros(N * D)])
    X = np.stack([x1 + 0.05 * rng.standard_normal(x1.size) for _ in range(B)])
    with BatchEvaluator(model, method, N, 0.01, theta, sig, R, obs_t, obs_y, m0, s0, np.zeros(B), B=B) as ev:
        F, G = ev.eval(X)
    for r in range(B):
        prob = Problem(model=model, method=method, D=D, N=N, dt=0.01, theta=theta, sigma=sig, R=R, obs_t=obs_t,
                       obs_y=obs_y[r], m0=m0[r], s0=s0, E0=0.0)
        Fo, Go = oracle.eval(prob, X[r])
        assert abs(F[r] - Fo) <= TOL * abs(Fo)
        assert grad_err(G[r], Go, prob.N, prob.D) < TOL


@pytest.mark.parametrize("model,method,N", [("L63", "rk2", 61), ("L63", "rk4", 30), ("OU", "rk4", 75),
                                            ("DW", "euler", 40), ("OU", "heun", 18), ("OU", "rk2", 75),
                                            ("DW", "rk2", 40)])
def test_large_batches_take_the_staged_forward_sweep(oracle, model, method, N):
    """Large batches run other kernels than small ones.  Lorenz-63: above 8192 problems the staged forward
    sweep (small_dim.cu: blocks of time indices through shared memory) and the one-thread-per-problem backward
    sweep instead of the lane-parallel kernels.  D = 1: F-only evaluations above 16384 problems, and RK2 (whose
    forward sweep is never time-parallel), take the staged forward / sequential backward sweeps; evaluations
    with a gradient run the fused time-parallel kernel at every size.  A problem's result in such a batch
    (ragged last warp, ragged last block) must agree with its result in a small batch to rounding, be bitwise
    independent of its position inside either, and sampled rows are held to the oracle."""
    from vgpa_b200.engine import BatchEvaluator
    D = 3 if model == "L63" else 1
    rng = np.random.default_rng(23)
    M = 5
    obs_t = np.linspace(0, N, M + 2, dtype=int)[1:-1]
    theta = [10.0, 28.0, 2.6667] if model == "L63" else ([2.0] if model == "OU" else [1.0])
    sig = np.full(D, 10.0 if model == "L63" else 0.8)
    R = np.full(D, 2.0 if model == "L63" else 0.04)
    B = 8500 if model == "L63" else 17000          # above the small-batch kernels' limits (8192 / 16384)
    obs_y = rng.standard_normal((B, M, D)) * (3.0 if model == "L63" else 0.5)
    m0 = rng.standard_normal((B, D))
    s0 = 0.25 * np.eye(D)
    a_diag = 0.5 * sig / 0.25
    x1 = np.concatenate([np.tile(np.diag(a_diag).ravel(), N), np.zeros(N * D)])
    X = x1[None, :] + 0.05 * rng.standard_normal((B, x1.size))
    with BatchEvaluator(model, method, N, 0.01, theta, sig, R, obs_t, obs_y, m0, s0, np.zeros(B), B=B) as ev:
        F, G = ev.eval(X)
        F_only, _ = ev.eval(X, want_grad=False)          # D = 1: the staged forward sweep whatever the solver
    assert np.abs(F_only - F).max() <= 1e-12 * np.abs(F).max()
    rows = np.concatenate([np.arange(0, 40), np.arange(B - 40, B)])        # first warps and the ragged last one
    with BatchEvaluator(model, method, N, 0.01, theta, sig, R, obs_t, obs_y[rows], m0[rows], s0,
                        np.zeros(rows.size), B=rows.size) as ev:
        Fs, Gs = ev.eval(X[rows])
    if model == "L63":
        # small Lorenz-63 batches run the lane-parallel kernels (l63_lanes.cu), large ones the one-thread-per-
        # problem kernels: the same operations in the same order, F bit for bit, but the compiler contracts a few
        # products differently around the lane-dependent control flow -- last-bit differences in isolated
        # gradient entries (observed: <= 2e-16 of the largest entry)
        assert np.array_equal(F[rows], Fs)
        assert np.abs(G[rows] - Gs).max() <= 1e-13 * np.abs(Gs).max()
    else:
        # small D = 1 batches run the time-parallel sweeps (small_dim.cu, scan1_*): each lane's run of steps
        # starts from a state composed by a scan, so results differ from the sequential kernels by rounding
        assert np.abs(F[rows] - Fs).max() <= 1e-12 * np.abs(Fs).max()
        assert np.abs(G[rows] - Gs).max() <= 1e-11 * np.abs(Gs).max()
    # inside one kernel family a problem's result does not depend on its position or on its neighbours
    rows2 = rows[::-1].copy()
    with BatchEvaluator(model, method, N, 0.01, theta, sig, R, obs_t, obs_y[rows2], m0[rows2], s0,
                        np.zeros(rows2.size), B=rows2.size) as ev:
        F2, G2 = ev.eval(X[rows2])
    assert np.array_equal(F2[::-1], Fs) and np.array_equal(G2[::-1], Gs)
    for r in (0, 31, 4242, B - 1):
        prob = Problem(model=model, method=method, D=D, N=N, dt=0.01, theta=theta, sigma=sig, R=R, obs_t=obs_t,
                       obs_y=obs_y[r], m0=m0[r], s0=s0, E0=0.0)
        Fo, Go = oracle.eval(prob, X[r])
        assert abs(F[r] - Fo) <= TOL * abs(Fo)
        assert grad_err(G[r], Go, prob.N, prob.D) < TOL
