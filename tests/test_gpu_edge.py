"""
Edge cases of the CUDA path against the oracle (through the C ABI): no observations,
observations at the ends of the grid (the jump at the last index is never applied,
SURVEY.md 3.8 item 4), the shortest grids, batches that span several chunks and
staging slots, the device-pointer entry point, and bitwise reproducibility.
"""
import numpy as np
import pytest

from conftest import golden_eval_files, grad_err, rel_err

pytestmark = pytest.mark.gpu
TOL = 1e-9


def _load(name):
    return np.load(str(next(p for p in golden_eval_files() if name in p)))


def _problem_and_evaluator(g, obs_t=None, obs_y=None, N=None, B=1, **kw):
    from oracle import Problem, prior_kl0
    from vgpa_b200.engine import BatchEvaluator
    D = int(g["D"])
    N = int(g["N"]) if N is None else N
    obs_t = g["obs_t"] if obs_t is None else np.asarray(obs_t, dtype=np.int64)
    obs_y = g["obs_y"] if obs_y is None else obs_y
    E0 = float(prior_kl0(g["m0"], g["s0"], g["mu0"], g["tau0"], D == 1))
    prob = Problem(model=str(g["model"]), method=str(g["method"]), D=D, N=N, dt=float(g["dt"]), theta=g["theta"],
                   sigma=g["sigma"], R=g["R"], obs_t=obs_t, obs_y=obs_y, m0=g["m0"], s0=g["s0"], E0=E0,
                   dt_model=float(g["dt"]))
    ev = BatchEvaluator(str(g["model"]), str(g["method"]), N, float(g["dt"]), g["theta"], g["sigma"], g["R"],
                        obs_t, obs_y, g["m0"], g["s0"], E0, B=B, dt_model=float(g["dt"]), **kw)
    return prob, ev


def _x_prefix(g, N):
    """The first N time points of the golden evaluation point."""
    D, N0 = int(g["D"]), int(g["N"])
    x = g["x"]
    return np.concatenate([x[:N0 * D * D].reshape(N0, -1)[:N].ravel(), x[N0 * D * D:].reshape(N0, -1)[:N].ravel()])


@pytest.mark.parametrize("name", ["eval_OU_rk4", "eval_L63_rk2", "eval_L96_rk2", "eval_L96_heun"])
def test_no_observations(oracle, name):
    g = _load(name)
    D = int(g["D"])
    prob, ev = _problem_and_evaluator(g, obs_t=np.zeros(0, dtype=np.int64), obs_y=np.zeros((0, D)))
    with ev:
        F, G = ev.eval(g["x"])
    F_o, g_o = oracle.eval(prob, g["x"])
    assert abs(F[0] - F_o) <= TOL * abs(F_o) and grad_err(G[0], g_o, prob.N, prob.D) < TOL


@pytest.mark.parametrize("name", ["eval_DW_heun", "eval_L63_euler", "eval_L96_rk4", "eval_L96_euler"])
def test_observations_at_both_ends(oracle, name):
    """obs at index 0 (its jump IS applied by the last backward step) and at index N-1
    (never applied: lam[N-1] = Psi[N-1] = 0 by construction)."""
    g = _load(name)
    D, N = int(g["D"]), int(g["N"])
    obs_t = np.array([0, N // 2, N - 1])
    rng = np.random.default_rng(2)
    obs_y = rng.standard_normal((3, D)) + g["m0"]
    prob, ev = _problem_and_evaluator(g, obs_t=obs_t, obs_y=obs_y)
    with ev:
        out = ev.eval_full(g["x"])
    ref = oracle.eval(prob, g["x"], full=True)
    for k in ("F", "Eobs", "grad", "lamt", "psit"):
        assert rel_err(out[k], ref[k]) < TOL, k
    assert np.all(out["lamt"][-1] == 0.0) and np.all(out["psit"][-1] == 0.0)


@pytest.mark.parametrize("name", ["eval_OU_rk4", "eval_L63_heun", "eval_L96_rk2", "eval_L96_rk4", "eval_L96_euler"])
@pytest.mark.parametrize("N", [2, 3, 4])
def test_shortest_grids(oracle, name, N):
    g = _load(name)
    D = int(g["D"])
    obs_t = np.array([1]) if N > 2 else np.zeros(0, dtype=np.int64)
    obs_y = (g["obs_y"][:1] if N > 2 else np.zeros((0, D)))
    prob, ev = _problem_and_evaluator(g, obs_t=obs_t, obs_y=obs_y, N=N)
    x = _x_prefix(g, N)
    with ev:
        F, G = ev.eval(x)
    F_o, g_o = oracle.eval(prob, x)
    assert abs(F[0] - F_o) <= TOL * abs(F_o) and grad_err(G[0], g_o, prob.N, prob.D) < TOL


@pytest.mark.parametrize("name", ["eval_L96_rk2", "eval_L63_rk4"])
def test_many_chunks_equal_one_chunk_bitwise(name):
    """A scratch budget of a few problems forces many passes and both staging slots;
    results must be bit-identical to the single-pass evaluation, and run to run."""
    g = _load(name)
    B = 11
    rng = np.random.default_rng(8)
    X = np.stack([g["x"] * (1.0 + 0.01 * rng.standard_normal(g["x"].size)) for _ in range(B)])
    per_problem = 8 * int(g["N"]) * (2 * int(g["D"]) + 2 * int(g["D"]) ** 2 + 1)
    _, ev1 = _problem_and_evaluator(g, B=B)
    _, ev3 = _problem_and_evaluator(g, B=B, scratch_bytes=3 * per_problem + 64)
    with ev1, ev3:
        assert ev1.chunk_size == B and ev3.chunk_size == 3
        F1, G1 = ev1.eval(X)
        F3, G3 = ev3.eval(X)
        F1b, G1b = ev1.eval(X)
    assert np.array_equal(F1, F3) and np.array_equal(G1, G3)
    assert np.array_equal(F1, F1b) and np.array_equal(G1, G1b)


def test_device_pointer_entry_point_matches_host_entry_point():
    import torch
    g = _load("eval_L96_rk2")
    B = 4
    rng = np.random.default_rng(9)
    X = np.stack([g["x"] * (1.0 + 0.01 * rng.standard_normal(g["x"].size)) for _ in range(B)])
    _, ev = _problem_and_evaluator(g, B=B)
    with ev:
        F_h, G_h = ev.eval(X)
        Xd = torch.from_numpy(X).cuda()
        Fd = torch.empty(B, dtype=torch.float64, device="cuda")
        Gd = torch.empty_like(Xd)
        s = torch.cuda.Stream()
        with torch.cuda.stream(s):
            ev.eval_device(Xd.data_ptr(), ev.n_x, Fd.data_ptr(), Gd.data_ptr(), ev.n_x, s.cuda_stream)
        ev.sync()
        assert ev.launch_count == 8
    assert np.array_equal(Fd.cpu().numpy(), F_h) and np.array_equal(Gd.cpu().numpy(), G_h)


def test_device_evaluations_on_two_streams_share_one_handle_safely():
    """A handle owns one scratch: two vgpa_eval_device calls enqueued back to back on DIFFERENT streams must not
    overlap (the second waits for the first on the device), so both return what a lone evaluation returns."""
    import torch
    g = _load("eval_L96_rk2")
    B = 6
    rng = np.random.default_rng(19)
    Xa = np.stack([g["x"] * (1.0 + 0.01 * rng.standard_normal(g["x"].size)) for _ in range(B)])
    Xb = np.stack([g["x"] * (1.0 + 0.01 * rng.standard_normal(g["x"].size)) for _ in range(B)])
    _, ev = _problem_and_evaluator(g, B=B)
    with ev:
        Fa_h, Ga_h = ev.eval(Xa)
        Fb_h, Gb_h = ev.eval(Xb)
        dev = [(torch.from_numpy(X).cuda(), torch.empty(B, dtype=torch.float64, device="cuda")) for X in (Xa, Xb)]
        grads = [torch.empty_like(d[0]) for d in dev]
        streams = [torch.cuda.Stream(), torch.cuda.Stream()]
        torch.cuda.synchronize()
        for rep in range(3):                      # alternate the streams a few times
            for (Xd, Fd), Gd, s in zip(dev, grads, streams):
                ev.eval_device(Xd.data_ptr(), ev.n_x, Fd.data_ptr(), Gd.data_ptr(), ev.n_x, s.cuda_stream)
        ev.sync()
        torch.cuda.synchronize()
    assert np.array_equal(dev[0][1].cpu().numpy(), Fa_h) and np.array_equal(grads[0].cpu().numpy(), Ga_h)
    assert np.array_equal(dev[1][1].cpu().numpy(), Fb_h) and np.array_equal(grads[1].cpu().numpy(), Gb_h)


def test_scratch_hand_over_between_handles():
    """vgpa_scratch_cache: with the hand-over on, the scratch of a destroyed handle is kept (free device memory
    does not come back) and serves the next handle of the same shape; switching it off releases it; results are
    those of a fresh handle either way."""
    import torch
    from vgpa_b200._lib import lib
    g = _load("eval_L96_rk2")
    B = 200                                   # S(t) and dE/dS scratch of ~54 MB each: above the 32 MB threshold
    X = g["x"]                                # one x shared by all problems
    _, ev = _problem_and_evaluator(g, B=B)
    with ev:
        F0, G0 = ev.eval(X)
    torch.cuda.synchronize()
    free0 = torch.cuda.mem_get_info()[0]
    assert lib.vgpa_scratch_cache(1) == 0
    try:
        _, ev = _problem_and_evaluator(g, B=B)
        with ev:
            F1, G1 = ev.eval(X)
        kept = free0 - torch.cuda.mem_get_info()[0]
        assert kept > (32 << 20)                             # the closed handle's scratch is still ours
        _, ev = _problem_and_evaluator(g, B=B)               # same shape: served from the kept blocks
        with ev:
            assert free0 - torch.cuda.mem_get_info()[0] <= kept + (64 << 20)
            F2, G2 = ev.eval(X)
    finally:
        released = lib.vgpa_scratch_cache(0)
    assert released >= kept - (64 << 20) and released > 0
    assert free0 - torch.cuda.mem_get_info()[0] < (64 << 20)
    assert np.array_equal(F1, F0) and np.array_equal(G1, G0) and np.array_equal(F2, F0) and np.array_equal(G2, G0)


def test_misaligned_device_buffers_are_rejected():
    import torch
    g = _load("eval_L96_rk2")
    _, ev = _problem_and_evaluator(g)
    with ev:
        buf = torch.zeros(ev.n_x + 8, dtype=torch.float64, device="cuda")
        F = torch.zeros(1, dtype=torch.float64, device="cuda")
        with pytest.raises(ValueError):
            ev.eval_device(buf.data_ptr() + 8, ev.n_x, F.data_ptr(), None, None, 0)


@pytest.mark.parametrize("name", ["eval_L96_rk2", "eval_L63_heun", "eval_OU_rk4"])
def test_active_set_skips_problems(name):
    """vgpa_set_active: problems flagged 0 are skipped by every kernel (F and gradient rows untouched),
    the others get bitwise the result of the unmasked evaluation; NULL restores the full batch."""
    import torch
    from conftest import golden_eval_files
    from test_gpu_parity import evaluator_from_golden
    g = np.load(str(next(p for p in golden_eval_files() if name in p)))
    B = 7
    rng = np.random.default_rng(11)
    X = np.stack([g["x"] * (1.0 + 1e-3 * rng.standard_normal(g["x"].size)) for _ in range(B)])
    who = np.array([1, 0, 1, 1, 0, 0, 1], dtype=np.int32)
    with evaluator_from_golden(g, B=B) as ev:
        n = ev.n_x
        Xd = torch.from_numpy(X).cuda()
        st = torch.cuda.current_stream().cuda_stream
        F0 = torch.empty(B, dtype=torch.float64, device="cuda")
        G0 = torch.empty((B, n), dtype=torch.float64, device="cuda")
        ev.eval_device(Xd.data_ptr(), n, F0.data_ptr(), G0.data_ptr(), n, st)
        ev.sync()
        F1 = torch.full((B,), -7.0, dtype=torch.float64, device="cuda")
        G1 = torch.full((B, n), -7.0, dtype=torch.float64, device="cuda")
        act = torch.from_numpy(who).cuda()
        ev.set_active(act.data_ptr())
        ev.eval_device(Xd.data_ptr(), n, F1.data_ptr(), G1.data_ptr(), n, st)
        ev.sync()
        ev.set_active(None)
        F2 = torch.empty(B, dtype=torch.float64, device="cuda")
        G2 = torch.empty((B, n), dtype=torch.float64, device="cuda")
        ev.eval_device(Xd.data_ptr(), n, F2.data_ptr(), G2.data_ptr(), n, st)
        ev.sync()
    on = who.astype(bool)
    F0, G0, F1, G1, F2, G2 = (t.cpu().numpy() for t in (F0, G0, F1, G1, F2, G2))
    assert np.array_equal(F1[on], F0[on]) and np.array_equal(G1[on], G0[on])
    assert np.all(F1[~on] == -7.0) and np.all(G1[~on] == -7.0)
    assert np.array_equal(F2, F0) and np.array_equal(G2, G0)
