"""
The premise of the D = 1 time-parallel CUDA sweeps (vgpa_b200/csrc/small_dim.cu, scan1_*), checked on the CPU
oracle: for Euler, Heun and RK4 the forward sweep's m(t), S(t) are AFFINE functions of the initial moments
(fwd_ode.py:41-75 integrates m' = -A m + b, S' = -2 A S + sigma), so the composition of solver steps is a
composition of affine maps; RK2 passes S in place of A in its first variance stage (runge_kutta2.py:92), which
makes S(t) a polynomial in s0 -- affine in m0 only.  (The CUDA kernels rely on exactly this split.)
"""
import numpy as np
import pytest

from oracle import Problem


def _sweep(oracle, model, method, x, m0, s0, N):
    M = 6
    obs_t = np.linspace(0, N, M + 2, dtype=int)[1:-1].astype(np.int64)
    prob = Problem(model=model, method=method, D=1, N=N, dt=0.01, theta=np.array([2.0 if model == "OU" else 1.0]),
                   sigma=np.array([0.8]), R=np.array([0.04]), obs_t=obs_t, obs_y=np.zeros((M, 1)),
                   m0=np.array([m0]), s0=np.array([[s0]]), E0=0.0)
    full = oracle.eval(prob, x, full=True)
    return full["mt"], full["st"]


@pytest.mark.parametrize("model", ["OU", "DW"])
@pytest.mark.parametrize("method", ["euler", "heun", "rk2", "rk4"])
def test_forward_moments_are_affine_in_the_initial_moments(oracle, model, method):
    N = 300
    rng = np.random.default_rng([5, method == "rk2", model == "DW"])
    x = np.concatenate([1.6 + 0.2 * rng.standard_normal(N), 0.3 * rng.standard_normal(N)])
    (ma, sa), (mb, sb) = _sweep(oracle, model, method, x, -0.7, 0.2, N), _sweep(oracle, model, method, x, 1.1, 0.6, N)
    mc, sc = _sweep(oracle, model, method, x, 0.25 * -0.7 + 0.75 * 1.1, 0.25 * 0.2 + 0.75 * 0.6, N)
    m_dev = np.abs(mc - (0.25 * ma + 0.75 * mb)).max()
    s_dev = np.abs(sc - (0.25 * sa + 0.75 * sb)).max()
    assert m_dev < 1e-13                              # the mean: affine under every solver
    if method == "rk2":
        assert s_dev > 1e-6                           # the variance under RK2: not affine (the quirk is real)
    else:
        assert s_dev < 1e-13
