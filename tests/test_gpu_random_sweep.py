"""
Randomised sweep of the CUDA path against the oracle (through the C ABI): for every model and solver,
several draws of the grid length, the batch size (odd and even: the Lorenz-63 kernels pack two problems
per warp), the observation set (count, positions -- including index 0 and the last index --, values), the
noises and the evaluation points; every problem of every batch is held to the oracle at 1e-9 (F, and the
gradient per block), and a masked re-evaluation (vgpa_set_active) must leave the masked rows untouched.
Seeds are fixed: the sweep is a regression net, not a fuzzer.
"""
import numpy as np
import pytest

from conftest import golden_eval_files, grad_err

pytestmark = pytest.mark.gpu
TOL = 1e-9


def _load(name):
    return np.load(str(next(p for p in golden_eval_files() if name in p)))


@pytest.mark.parametrize("model", ["DW", "OU", "L63", "L96"])
@pytest.mark.parametrize("method", ["euler", "heun", "rk2", "rk4"])
def test_random_configurations_match_oracle(oracle, model, method):
    import torch
    from oracle import Problem, prior_kl0
    from vgpa_b200.engine import BatchEvaluator
    g = _load(f"eval_{model}_{method}")
    D, N0 = int(g["D"]), int(g["N"])
    rng = np.random.default_rng([17, D, {"euler": 0, "heun": 1, "rk2": 2, "rk4": 3}[method]])
    n_draws = 2 if model == "L96" else 4
    for draw in range(n_draws):
        N = int(rng.integers(2, N0 + 1)) if model != "L96" else int(rng.integers(3, N0 + 1))
        B = int(rng.integers(1, 8))
        M = int(rng.integers(0, min(6, N) + 1))
        obs_t = np.sort(rng.choice(N, size=M, replace=False)).astype(np.int64)
        if M >= 2 and draw % 2 == 0:
            obs_t[0], obs_t[-1] = 0, N - 1                      # both ends of the grid
            obs_t = np.unique(obs_t)
            M = obs_t.size
        scale = float(np.abs(g["obs_y"]).max()) if g["obs_y"].size else 1.0
        obs_y = scale * 0.5 * rng.standard_normal((B, M, D))
        sigma = np.stack([g["sigma"] * rng.uniform(0.6, 1.5) for _ in range(B)])
        R = np.stack([g["R"] * rng.uniform(0.6, 1.5) for _ in range(B)])
        theta = np.stack([g["theta"] * rng.uniform(0.95, 1.05) for _ in range(B)])
        m0 = np.stack([g["m0"] + 0.05 * rng.standard_normal(D) for _ in range(B)])
        s0 = np.stack([g["s0"] * rng.uniform(0.8, 1.3) for _ in range(B)])
        E0 = np.array([prior_kl0(m0[i], s0[i].reshape(D, D) if D > 1 else s0[i], g["mu0"], g["tau0"], D == 1)
                       for i in range(B)])
        # the first N time points of the golden evaluation point, jittered per problem
        x = g["x"]
        xN = np.concatenate([x[:N0 * D * D].reshape(N0, -1)[:N].ravel(), x[N0 * D * D:].reshape(N0, -1)[:N].ravel()])
        X = np.stack([xN * (1.0 + 0.01 * rng.standard_normal(xN.size)) for _ in range(B)])
        dt = float(g["dt"])
        with BatchEvaluator(model, method, N, dt, theta, sigma, R, obs_t, obs_y, m0, s0, E0, B=B, dt_model=dt) as ev:
            F, G = ev.eval(X)
            # masked device evaluation: rows whose flag is 0 keep their sentinel values
            Xd = torch.from_numpy(X).cuda()
            Fd = torch.full((B,), -7.0, dtype=torch.float64, device="cuda")
            Gd = torch.full_like(Xd, -7.0)
            keep = rng.integers(0, 2, size=B).astype(np.int32)
            mask = torch.from_numpy(keep).cuda()
            ev.set_active(mask.data_ptr())
            ev.eval_device(Xd.data_ptr(), ev.n_x, Fd.data_ptr(), Gd.data_ptr(), ev.n_x, torch.cuda.current_stream().cuda_stream)
            ev.sync()
            ev.set_active(None)
            Fm, Gm = Fd.cpu().numpy(), Gd.cpu().numpy()
        for i in range(B):
            prob = Problem(model=model, method=method, D=D, N=N, dt=dt, theta=theta[i], sigma=sigma[i], R=R[i],
                           obs_t=obs_t, obs_y=obs_y[i], m0=m0[i], s0=s0[i], E0=float(E0[i]), dt_model=dt)
            Fo, Go = oracle.eval(prob, X[i])
            tag = (model, method, draw, N, B, M, i)
            assert abs(F[i] - Fo) <= TOL * abs(Fo), tag
            assert grad_err(G[i], Go, N, D) < TOL, tag
            if keep[i]:
                assert Fm[i] == F[i] and np.array_equal(Gm[i], G[i]), tag
            else:
                assert Fm[i] == -7.0 and np.all(Gm[i] == -7.0), tag


@pytest.mark.parametrize("model", ["DW", "OU"])
@pytest.mark.parametrize("method", ["euler", "heun", "rk2", "rk4"])
def test_time_parallel_sweeps_at_run_boundaries(oracle, model, method):
    """The D = 1 time-parallel sweeps (small_dim.cu, scan1_*) cut the N-1 steps into 128 runs: grid lengths
    with fewer steps than threads (2, 3, 100), exactly one step per thread and one more (129, 130), ragged
    last runs (300, 1538), a grid too long for the fused kernel's shared memory but not for the separate
    time-parallel sweeps (3800) and one too long for those as well (4300: the sequential kernels), each with
    observations on both ends of the grid and on neighbouring indices, against the oracle."""
    from oracle import Problem
    from vgpa_b200.engine import BatchEvaluator
    rng = np.random.default_rng([23, {"euler": 0, "heun": 1, "rk2": 2, "rk4": 3}[method], model == "OU"])
    theta = np.array([2.0] if model == "OU" else [1.0])
    B = 3
    for N in (2, 3, 100, 129, 130, 300, 1538, 3800, 4300):
        M = min(N, 12)
        obs_t = np.unique(np.concatenate([[0, N - 1], [N // 2, min(N // 2 + 1, N - 1)],
                                          rng.choice(N, size=M, replace=False)])).astype(np.int64)
        M = obs_t.size
        obs_y = 0.7 * rng.standard_normal((B, M, 1))
        sigma = rng.uniform(0.5, 1.2, size=(B, 1))
        R = rng.uniform(0.03, 0.08, size=(B, 1))
        m0 = rng.standard_normal((B, 1))
        s0 = rng.uniform(0.2, 0.4, size=(B, 1, 1))
        E0 = rng.standard_normal(B)
        X = np.concatenate([1.6 + 0.2 * rng.standard_normal((B, N)), 0.3 * rng.standard_normal((B, N))], axis=1)
        with BatchEvaluator(model, method, N, 0.01, theta, sigma, R, obs_t, obs_y, m0, s0, E0, B=B) as ev:
            F, G = ev.eval(X)
        for i in range(B):
            prob = Problem(model=model, method=method, D=1, N=N, dt=0.01, theta=theta, sigma=sigma[i], R=R[i],
                           obs_t=obs_t, obs_y=obs_y[i], m0=m0[i], s0=s0[i], E0=float(E0[i]))
            Fo, Go = oracle.eval(prob, X[i])
            assert abs(F[i] - Fo) <= TOL * abs(Fo), (N, i)
            assert grad_err(G[i], Go, N, 1) < TOL, (N, i)


_FAMILY_SCRIPT = r"""
import sys, numpy as np
sys.path.insert(0, sys.argv[1])
from vgpa_b200.engine import BatchEvaluator
z = np.load(sys.argv[2])
out = {}
for model in ("DW", "OU"):
    for method in ("euler", "heun", "rk2", "rk4"):
        with BatchEvaluator(model, method, int(z["N"]), 0.01, z["theta_" + model], z["sigma"], z["R"], z["obs_t"], z["obs_y"],
                            z["m0"], z["s0"], z["E0"], B=int(z["B"])) as ev:
            F, G = ev.eval(z["X"])
            F1, _ = ev.eval(z["X"], want_grad=False)
        out[f"F_{model}_{method}"], out[f"G_{model}_{method}"], out[f"F1_{model}_{method}"] = F, G, F1
np.savez(sys.argv[3], **out)
"""


def test_time_parallel_and_sequential_kernel_families_agree(tmp_path):
    """The same D = 1 problems (N = 1001, every model and solver) through the time-parallel kernels and, in a second
    process with VGPA_SEQUENTIAL_D1=1, through the one-thread-per-problem kernels: F and the gradient agree to
    rounding (the run-start states of the time-parallel sweeps come out of a scan), far inside the 1e-9 bar that
    both hold against the oracle."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    rng = np.random.default_rng(31)
    N, B, M = 1001, 5, 20
    obs_t = np.linspace(0, N, M + 2, dtype=int)[1:-1].astype(np.int64)
    inp = dict(N=N, B=B, theta_DW=np.array([1.0]), theta_OU=np.array([2.0]), sigma=rng.uniform(0.5, 1.2, (B, 1)),
               R=rng.uniform(0.03, 0.08, (B, 1)), obs_t=obs_t, obs_y=0.7 * rng.standard_normal((B, M, 1)),
               m0=rng.standard_normal((B, 1)), s0=rng.uniform(0.2, 0.4, (B, 1, 1)), E0=rng.standard_normal(B),
               X=np.concatenate([1.6 + 0.2 * rng.standard_normal((B, N)), 0.3 * rng.standard_normal((B, N))], axis=1))
    fin = os.path.join(str(tmp_path), "in.npz")
    np.savez(fin, **inp)
    res = {}
    for tag, flag in (("scan", "0"), ("seq", "1")):
        fout = os.path.join(str(tmp_path), tag + ".npz")
        env = dict(os.environ, VGPA_SEQUENTIAL_D1=flag)
        subprocess.run([sys.executable, "-c", _FAMILY_SCRIPT, root, fin, fout], check=True, env=env, timeout=600)
        res[tag] = np.load(fout)
    differ = 0
    for key in res["scan"].files:
        a, b = res["scan"][key], res["seq"][key]
        assert np.abs(a - b).max() <= 1e-12 * np.abs(b).max(), key
        differ += int(not np.array_equal(a, b))
    assert differ > 0          # the switch did select another kernel family
