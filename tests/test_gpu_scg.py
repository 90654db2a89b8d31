"""
GPU tests through the reference-facing interface (Simulation / VarGP / SCG):
the SCG convergence trace of the reference, recorded in tests/golden/scg_*.npz with
the reference's own optimiser, must be reproduced within 1e-6 on the common prefix
(BASELINE.json), and the known answers at the north-star shape (L96 D=40, N=1001)
must be matched to 1e-9.
"""
import sys

import numpy as np
import pytest

from conftest import GOLDEN, rel_err, stop_tolerance

sys.path.insert(0, str(GOLDEN))
import make_golden as mg  # noqa: E402

pytestmark = pytest.mark.gpu

JOBS = {"DW": ("euler", None), "OU": ("rk4", None), "L63": ("heun", 2.0), "L96": ("rk2", 0.2)}


@pytest.mark.parametrize("model", ["DW", "OU", "L63", "L96"])
def test_scg_trace_matches_reference(model):
    from vgpa_b200 import SCG, Simulation
    g = np.load(GOLDEN / f"scg_{model}.npz")
    method, tf = JOBS[model]
    sim = Simulation("t")
    sim.setup(mg.config(model, method, tf))
    vgpa = sim.build()
    x0 = vgpa.initialization()
    assert np.array_equal(x0, g["x"])
    scg = SCG(vgpa.free_energy, vgpa.gradient,
              {"max_it": int(g["max_it"]), "x_tol": 1.0e-6, "f_tol": 1.0e-8, "display": False})
    x, fx = scg(x0.copy())
    n_ref, n_new = int(g["n_it"]), int(scg.stats["MaxIt"])
    n = min(n_ref, n_new)
    assert abs(n_ref - n_new) <= stop_tolerance(g["trace_fx"], n_ref), (n_ref, n_new)
    ref, new = g["trace_fx"][:n], scg.stats["fx"][:n]
    assert np.max(np.abs(new - ref) / np.maximum(np.abs(ref), 1.0)) < 1e-6
    assert abs(fx - float(g["fx_final"])) <= 1e-6 * max(abs(float(g["fx_final"])), 1.0)
    # one CUDA evaluation per distinct x: fewer than the reference's f_eval count
    assert vgpa.n_eval <= float(g["f_eval"])
    vgpa.close()


@pytest.mark.parametrize("model", ["L63", "L96"])
def test_full_scg_optimisation_matches_reference(model):
    """BASELINE configs[3] (L96 D=40, T=1000, RK2) and configs[2] (L63, T=2000, Heun as shipped): the
    whole SCG optimisation to convergence, against the trace the unmodified reference produced
    (scg_<model>_full.npz; 18 and 3 minutes of CPU there, about a second here)."""
    from vgpa_b200 import SCG, Simulation
    g = np.load(GOLDEN / f"scg_{model}_full.npz")
    sim = Simulation("t")
    sim.setup(mg.config(model, str(g["method"]), float(g["tf"])))
    vgpa = sim.build()
    x0 = vgpa.initialization()
    scg = SCG(vgpa.free_energy, vgpa.gradient,
              {"max_it": int(g["max_it"]), "x_tol": 1.0e-6, "f_tol": 1.0e-8, "display": False})
    x, fx = scg(x0.copy())
    n_ref, n_new = int(g["n_it"]), int(scg.stats["MaxIt"])
    n = min(n_ref, n_new)
    assert abs(n_ref - n_new) <= max(2, n_ref // 50), (n_ref, n_new)
    ref, new = g["trace_fx"][:n], scg.stats["fx"][:n]
    assert np.max(np.abs(new - ref) / np.maximum(np.abs(ref), 1.0)) < 1e-6
    assert abs(fx - float(g["fx_final"])) <= 1e-6 * max(abs(float(g["fx_final"])), 1.0)
    vgpa.close()


def test_l96_north_star_known_answers():
    """L96 D=40, tf=10 (N=1001, T=1000 steps), RK2: F(x0), |grad|, gradient samples."""
    from vgpa_b200 import Simulation
    g = np.load(GOLDEN / "known_L96_N1001.npz")
    sim = Simulation("t")
    sim.setup(mg.config("L96", "rk2", 10.0))
    vgpa = sim.build()
    x0 = vgpa.initialization()
    assert x0.size == 1001 * 40 * 41
    assert np.array_equal(x0[g["x0_idx"]], g["x0_samples"]) and x0.sum() == float(g["x0_sum"])
    F = vgpa.free_energy(x0)
    grad = vgpa.gradient(x0)
    assert abs(F - float(g["F_x0"])) <= 1e-9 * abs(float(g["F_x0"]))
    assert abs(np.linalg.norm(grad) - float(g["gnorm_x0"])) <= 1e-9 * float(g["gnorm_x0"])
    assert np.abs(grad[g["g_idx"]] - g["g_samples"]).max() <= 1e-9 * float(g["g_absmax"])
    out = vgpa.arg_out
    assert out["mt"].shape == (1001, 40) and out["psit"].shape == (1001, 40, 40)
    assert np.all(out["lamt"][-1] == 0.0) and np.all(out["psit"][-1] == 0.0)
    vgpa.close()


def test_reference_operator_objects_on_gpu():
    """The level-3 operator objects (FwdOde, BwdOde, model.energy, likelihood) chained
    by hand as VarGP.free_energy does (variational.py:169-181) give the fused answer."""
    from vgpa_b200 import Simulation
    g = np.load(GOLDEN / "eval_L63_rk4.npz")
    sim = Simulation("t")
    sim.setup(mg.config("L63", "rk4", 2.0))
    vgpa = sim.build()
    N, D = vgpa.dim_n, vgpa.dim_d
    x = g["x"]
    A, b = x[:N * D * D].reshape(N, D, D), x[N * D * D:].reshape(N, D)
    md = sim.m_data
    mt, st = vgpa.fwd_ode(A, b, md["m0"], md["s0"], md["model"].sigma)
    Eobs = vgpa.likelihood(mt, st)
    Esde, (Efx, Edf), (dm, ds, *_) = md["model"].energy(A, b, mt, st, vgpa.obs_t)
    jm, js, *_ = vgpa.likelihood.gradients(mt, st)
    lam, psi = vgpa.bwd_ode(A, dm, ds, jm, js)
    E0 = vgpa.kl0(md["m0"], md["s0"])
    F = float(E0 + Esde + Eobs)
    assert abs(F - float(g["F"])) <= 1e-9 * abs(float(g["F"]))
    assert rel_err(lam, g["lamt"]) < 1e-9 and rel_err(psi, g["psit"]) < 1e-9
    assert abs(vgpa.free_energy(x) - F) <= 1e-9 * abs(F)
    vgpa.close()


def test_gradient_arrays_are_owned_by_the_caller():
    """VarGP.gradient returns the page-locked buffer the device copy landed in, without a copy; a
    buffer is reused only when the caller has dropped every reference to it (optim_scg.py keeps the
    last two gradients alive and expects them to stay intact)."""
    from vgpa_b200 import Simulation
    sim = Simulation("t")
    sim.setup(mg.config("L63", "rk2", 2.0))
    vgpa = sim.build()
    x0 = vgpa.initialization()
    rng = np.random.default_rng(0)
    xs = [x0 * (1.0 + 1e-3 * rng.uniform(-1, 1, x0.size)) for _ in range(6)]
    held, copies = [], []
    for x in xs[:4]:                     # hold four gradients alive across later evaluations
        vgpa.free_energy(x)
        g = vgpa.gradient(x)
        held.append(g)
        copies.append(g.copy())
    assert len({id(g) for g in held}) == 4
    for g, c in zip(held, copies):
        assert np.array_equal(g, c)
    # the same x again: an equal but distinct array (the first one is the caller's, even if modified)
    held[3] *= 2.0
    again = vgpa.gradient(xs[3])
    assert again is not held[3] and np.array_equal(again, copies[3])
    # dropping references lets buffers be reused: the pool stays small over many evaluations
    del held, again, g
    import gc
    gc.collect()
    n_buffers = vgpa._pin[1]._allocated
    for k in range(24):
        x = xs[k % 6] * (1.0 + 1e-6 * k)
        vgpa.free_energy(x)
        last = vgpa.gradient(x)
    assert vgpa._pin[1]._allocated == n_buffers and vgpa._pin[1].outstanding <= 2
    # more gradients alive than the pool holds: still correct (copies out of a private buffer)
    many = [vgpa.gradient(x * (1.0 + 1e-5 * k)) for k in range(12)]
    ref = vgpa.gradient(x * (1.0 + 1e-5 * 11))
    assert np.array_equal(many[11], ref) and len({id(m) for m in many}) == 12
    # closing the object must not pull the memory from under arrays the caller still holds
    keep = last.copy()
    pool = vgpa._pin[1]
    vgpa.close()
    assert np.array_equal(last, keep) and np.array_equal(many[0], many[0].copy())
    # ... and what the caller still holds frees itself when dropped (nothing is parked forever)
    held_after_close = pool.outstanding
    assert held_after_close >= 1
    view = last[::2]
    del last
    assert pool.outstanding == held_after_close        # a view keeps its buffer alive
    del view, many
    gc.collect()
    assert pool._closed and pool._free == []


@pytest.mark.parametrize("model", ["DW", "L63"])
def test_simulation_run_with_device_optimizer(model, capsys):
    """Simulation.run(optimizer="device"): the whole optimisation resident on the GPU (BatchedSCG, batch of
    one) must reproduce the reference's recorded trace like the host loop does, and fill the same outputs."""
    from vgpa_b200 import Simulation
    g = np.load(GOLDEN / f"scg_{model}.npz")
    method, tf = JOBS[model]
    sim = Simulation("t")
    sim.setup(mg.config(model, method, tf))
    sim.run(max_it=int(g["max_it"]), display=False, optimizer="device")
    capsys.readouterr()
    st = sim.scg_stats
    n_ref, n_new = int(g["n_it"]), int(st["MaxIt"])
    n = min(n_ref, n_new)
    assert abs(n_ref - n_new) <= max(2, n_ref // 50), (n_ref, n_new)
    ref, new = g["trace_fx"][:n], st["fx"][:n]
    assert np.max(np.abs(new - ref) / np.maximum(np.abs(ref), 1.0)) < 1e-6
    assert abs(sim.output["fx"] - float(g["fx_final"])) <= 1e-6 * max(abs(float(g["fx_final"])), 1.0)
    assert {"at", "bt", "fx", "m0", "s0", "mt", "st", "Efx", "Edf", "lamt", "psit"} <= set(sim.output)


def test_cache_is_not_fooled_by_in_place_mutation():
    """The reference recomputes free_energy(x) on every call; the CUDA VarGP caches the last evaluation and must
    therefore notice an x that was MUTATED IN PLACE between calls (coordinate-wise finite differences, sparse
    updates), including at entries a strided probe would skip."""
    from vgpa_b200 import Simulation
    sim = Simulation("t")
    sim.setup(mg.config("L63", "rk2", 20.0))                         # 24 024 parameters: the probe looks at every 11th
    vgpa = sim.build()
    x = vgpa.initialization()
    assert vgpa._probe_step(x.size) > 1
    f0 = vgpa.free_energy(x)
    n0 = vgpa.n_eval
    assert vgpa.free_energy(x) == f0 and vgpa.n_eval == n0          # same array, same content: cached
    for pos in (1, x.size // 3 + 1, x.size - 1):                     # odd positions: off any power-of-two stride
        old = x[pos]
        x[pos] += 1e-3
        f1 = vgpa.free_energy(x)
        assert f1 != f0 and vgpa.n_eval > n0
        n0 = vgpa.n_eval
        x[pos] = old
        assert vgpa.free_energy(x) == f0                             # and back: recomputed, same value
        n0 = vgpa.n_eval
    g = vgpa.gradient(x)
    h = 1e-6
    x[5] += h
    fp = vgpa.free_energy(x)
    x[5] -= 2 * h
    fm = vgpa.free_energy(x)
    x[5] += h
    assert np.isfinite((fp - fm) / (2 * h)) and g.shape == x.shape
    vgpa.close()
