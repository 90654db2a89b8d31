"""
GPU parity tests proper: the CUDA path, called through the C ABI
(libvgpa_b200.so via vgpa_b200.engine), against
  (i)  the committed reference outputs (tests/golden/*.npz) and
  (ii) the CPU oracle on the same seeded inputs.
Tolerance: relative 1e-9 (max-abs error over max-abs value) on F, on the
gradient and on every intermediate -- BASELINE.json's stated bound.
"""
import numpy as np
import pytest

from conftest import golden_eval_files, grad_err, key_err, rel_err

pytestmark = pytest.mark.gpu

TOL = 1e-9
KEYS = ("F", "E0", "Esde", "Eobs", "grad", "mt", "st", "lamt", "psit", "Efx", "Edf",
        "dEsde_dm", "dEsde_ds")


def _engine():
    from vgpa_b200 import engine
    return engine


def evaluator_from_golden(g, B=1, **over):
    from oracle import prior_kl0
    eng = _engine()
    D, N = int(g["D"]), int(g["N"])
    tk = np.arange(0.0, float(g["tf"]) + float(g["dt"]), float(g["dt"]))
    E0 = prior_kl0(g["m0"], g["s0"], g["mu0"], g["tau0"], D == 1)
    kw = dict(model=str(g["model"]), method=str(g["method"]), N=N, dt=float(g["dt"]),
              theta=g["theta"], sigma=g["sigma"], R=g["R"], obs_t=g["obs_t"], obs_y=g["obs_y"],
              m0=g["m0"], s0=g["s0"], E0=E0, B=B, dt_model=float(abs(tk[1] - tk[0])))
    kw.update(over)
    return eng.BatchEvaluator(**kw)


@pytest.mark.parametrize("path", golden_eval_files(), ids=lambda p: p.split("eval_")[-1][:-4])
def test_eval_full_matches_reference(path):
    g = np.load(path)
    with evaluator_from_golden(g) as ev:
        out = ev.eval_full(g["x"])
    N, D = int(g["N"]), int(g["D"])
    bad = {k: key_err(k, out[k], g[k], N, D) for k in KEYS if key_err(k, out[k], g[k], N, D) >= TOL}
    assert not bad, bad


@pytest.mark.parametrize("path", golden_eval_files(), ids=lambda p: p.split("eval_")[-1][:-4])
def test_eval_matches_oracle_at_x0(oracle, path):
    """Batched entry point (vgpa_eval, host buffers) at the reference's x0."""
    from oracle import Problem
    g = np.load(path)
    prob = Problem.from_golden(g)
    F_o, g_o = oracle.eval(prob, g["x0"])
    with evaluator_from_golden(g) as ev:
        F, G = ev.eval(g["x0"])
    assert abs(F[0] - F_o) <= TOL * abs(F_o)
    assert abs(F[0] - float(g["F_x0"])) <= TOL * abs(float(g["F_x0"]))
    assert grad_err(G[0], g_o, int(g["N"]), int(g["D"])) < TOL
    assert abs(np.linalg.norm(G[0]) - float(g["gnorm_x0"])) <= TOL * float(g["gnorm_x0"])


@pytest.mark.parametrize("model", ["DW", "OU", "L63", "L96"])
@pytest.mark.parametrize("method", ["euler", "rk2", "rk4"])
def test_batch_of_distinct_problems(oracle, model, method):
    """B problems that differ in observations, noises, drift, initial moments and x:
    every row must equal the oracle's single-problem answer, and a shared-x call
    must equal the per-row call with x repeated."""
    from oracle import Problem, prior_kl0
    g = np.load(str(next(p for p in golden_eval_files() if f"eval_{model}_{method}" in p)))
    D, N, B = int(g["D"]), int(g["N"]), 5
    rng = np.random.default_rng([11, D, N])
    base = Problem.from_golden(g)
    obs_y = np.stack([g["obs_y"] + 0.1 * rng.standard_normal(g["obs_y"].shape) for _ in range(B)])
    sigma = np.stack([g["sigma"] * (0.7 + 0.15 * i) for i in range(B)])
    R = np.stack([g["R"] * (0.8 + 0.1 * i) for i in range(B)])
    theta = np.stack([g["theta"] * (1.0 + 0.01 * i) for i in range(B)])
    m0 = np.stack([g["m0"] + 0.05 * rng.standard_normal(D) for _ in range(B)])
    s0 = np.stack([g["s0"] * (1.0 + 0.1 * i) for i in range(B)])
    E0 = np.array([prior_kl0(m0[i], s0[i].reshape(D, D) if D > 1 else s0[i], g["mu0"], g["tau0"], D == 1)
                   for i in range(B)])
    X = np.stack([g["x"] * (1 + 0.01 * rng.standard_normal(g["x"].size)) for _ in range(B)])
    probs = [Problem(model=base.model, method=base.method, D=D, N=N, dt=base.dt, theta=theta[i],
                     sigma=sigma[i], R=R[i], obs_t=g["obs_t"], obs_y=obs_y[i], m0=m0[i], s0=s0[i],
                     E0=float(E0[i]), dt_model=base.dt_model) for i in range(B)]
    with evaluator_from_golden(g, B=B, theta=theta, sigma=sigma, R=R, obs_y=obs_y, m0=m0, s0=s0,
                               E0=E0) as ev:
        F, G = ev.eval(X)
        Fs, Gs = ev.eval(X[2])        # one x shared by all problems
        F1, _ = ev.eval(X, want_grad=False)
    for i in range(B):
        F_o, g_o = oracle.eval(probs[i], X[i])
        assert abs(F[i] - F_o) <= TOL * abs(F_o), i
        assert grad_err(G[i], g_o, N, D) < TOL, i
        F_s, g_s = oracle.eval(probs[i], X[2])
        assert abs(Fs[i] - F_s) <= TOL * abs(F_s), i
        assert grad_err(Gs[i], g_s, N, D) < TOL, i
    assert np.array_equal(F, F1)


@pytest.mark.parametrize("path", [p for p in golden_eval_files()],
                         ids=lambda p: p.split("eval_")[-1][:-4])
def test_operator_level_sweeps_and_energy(oracle, path):
    """FwdOde / BwdOde / model.energy entry points on their own."""
    from oracle import Problem
    eng = _engine()
    g = np.load(path)
    D, N = int(g["D"]), int(g["N"])
    x = g["x"]
    A = x[:N * D * D].reshape((N,) if D == 1 else (N, D, D))
    b = x[N * D * D:].reshape((N,) if D == 1 else (N, D))
    mt, st = eng.solve_fwd(str(g["method"]), A, b, g["m0"], g["s0"], g["sigma"], float(g["dt"]))
    assert rel_err(mt, g["mt"]) < TOL and rel_err(st, g["st"]) < TOL
    # dense jump tables from the reference's own outputs
    prob = Problem.from_golden(g)
    jm = np.zeros(N * D)
    js = np.zeros(N * D * D)
    cp = prob.c_struct()
    import ctypes as C
    dp = C.POINTER(C.c_double)
    oracle.lib.oracle_eobs_grad(C.byref(cp), g["mt"].ravel().ctypes.data_as(dp), jm.ctypes.data_as(dp),
                                js.ctypes.data_as(dp))
    shp_v = (N,) if D == 1 else (N, D)
    shp_m = (N,) if D == 1 else (N, D, D)
    lam, psi = eng.solve_bwd(str(g["method"]), A, g["dEsde_dm"], g["dEsde_ds"], jm.reshape(shp_v),
                             js.reshape(shp_m), float(g["dt"]))
    assert rel_err(lam, g["lamt"]) < TOL and rel_err(psi, g["psit"]) < TOL
    Esde, Ef, Edf, dm, ds = eng.model_energy(str(g["model"]), g["theta"], g["sigma"], A, b, g["mt"], g["st"],
                                             prob.dt_model)
    assert abs(Esde - float(g["Esde"])) <= TOL * abs(float(g["Esde"]))
    for got, key in ((Ef, "Efx"), (Edf, "Edf"), (dm, "dEsde_dm"), (ds, "dEsde_ds")):
        assert rel_err(got, g[key]) < TOL, key


def test_not_positive_definite_is_linalg_error():
    g = np.load(str(next(p for p in golden_eval_files() if "eval_L96_euler" in p)))
    N, D = int(g["N"]), int(g["D"])
    x = g["x"].copy()
    x[:N * D * D] = -200.0 * np.tile(np.eye(D).ravel(), N)
    x[:N * D * D] += 50.0 * np.random.default_rng(0).standard_normal(N * D * D)
    with evaluator_from_golden(g) as ev:
        with pytest.raises(np.linalg.LinAlgError):
            ev.eval(x)
        F, _ = ev.eval(g["x"])        # the handle stays usable
        assert abs(F[0] - float(g["F"])) <= TOL * abs(float(g["F"]))


def test_invalid_arguments_are_value_errors():
    eng = _engine()
    g = np.load(str(next(p for p in golden_eval_files() if "eval_OU_rk4" in p)))
    with pytest.raises(ValueError):
        evaluator_from_golden(g, model="XX")
    with pytest.raises(ValueError):
        evaluator_from_golden(g, method="leapfrog")
    with pytest.raises(ValueError):
        evaluator_from_golden(g, dt=-0.01)
    with pytest.raises(ValueError):
        evaluator_from_golden(g, sigma=np.array([-1.0]))
    with evaluator_from_golden(g) as ev:
        with pytest.raises(ValueError):
            ev.eval(np.zeros(7))


@pytest.mark.parametrize("method", ["rk2", "rk4"])
def test_l96_eight_observations_all_intermediates(method):
    """L96, N = 101, M = 8 observations (tests/golden/mid_L96_*.npz, generated by the unmodified
    reference): every intermediate -- in particular lamt / psit across all eight jumps and the
    observation-ordinal quirk F5 -- at 1e-9, the gradient per block."""
    from test_oracle_golden import load_mid, mid_errors
    g = load_mid(method)
    with evaluator_from_golden(g) as ev:
        out = ev.eval_full(g["x"])
        F, G = ev.eval(g["x"])
    bad = {k: e for k, e in mid_errors(out, g).items() if e >= TOL}
    assert not bad, bad
    assert abs(F[0] - float(g["F"])) <= TOL * abs(float(g["F"]))
    assert grad_err(G[0], g["grad"], int(g["N"]), int(g["D"])) < TOL
