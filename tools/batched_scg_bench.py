"""Device-resident batched SCG at the BASELINE configs[4] shape: B Lorenz-96 D=40 T=1000
problems (own observation set each, the reference's x0 of each from the on-device initialisation),
optimised to convergence at once.  Prints wall time, iterations and evaluations."""
import sys, time, json
import numpy as np
sys.path.insert(0, "."); sys.path.insert(0, "tests/golden")
import make_golden as mg
from vgpa_b200 import Simulation, BatchEvaluator
from vgpa_b200.batched_scg import BatchedSCG


def main(B=296, model="L96", method="rk2", tf=10.0, max_it=500):
    import torch
    sim = Simulation("cfg"); sim.setup(mg.config(model, method, tf))
    md = sim.m_data
    v = sim.build()
    D, N = v.dim_d, v.dim_n
    rng = np.random.default_rng(7)
    obs = np.asarray(md["obs_y"], dtype=float).reshape(-1, D)
    R = np.diagonal(np.atleast_2d(md["obs_noise"])).copy() if D > 1 else np.atleast_1d(float(md["obs_noise"]))
    obs_y = obs[None] + np.sqrt(R)[None, None, :] * rng.standard_normal((B,) + obs.shape)     # B observation sets
    sig = np.diagonal(md["model"].sigma).copy() if D > 1 else np.atleast_1d(float(md["model"].sigma))
    E0 = float(np.asarray(v.kl0(md["m0"], md["s0"])))
    ev = BatchEvaluator(model, method, N, float(md["time_window"]["dt"]), np.atleast_1d(md["model"].theta), sig, R,
                        np.asarray(md["obs_t"], dtype=np.int64), obs_y, np.atleast_1d(md["m0"]),
                        np.asarray(md["s0"], dtype=float).reshape(D, D), E0, B=B, dt_model=float(v.dt))
    X0 = torch.empty((B, ev.n_x), dtype=torch.float64, device="cuda")
    t0 = time.perf_counter()
    ev.initialization_device(X0.data_ptr(), ev.n_x, float(md["time_window"]["t0"]), torch.cuda.current_stream().cuda_stream)
    ev.sync()
    t_init = time.perf_counter() - t0
    opt = BatchedSCG(ev, {"max_it": max_it, "x_tol": 1.0e-6, "f_tol": 1.0e-8, "display": False})
    t0 = time.perf_counter()
    X, fx = opt(X0, adopt=True)
    torch.cuda.synchronize()
    el = time.perf_counter() - t0
    st = opt.stats
    print(json.dumps({"model": model, "method": method, "N": N, "B": B, "seconds": round(el, 3),
                      "init_seconds": round(t_init, 4), "batch_evaluations": int(st["evaluations"]),
                      "iterations_min_median_max": [int(st["MaxIt"].min()), int(np.median(st["MaxIt"])), int(st["MaxIt"].max())],
                      "fx0_mean": float(st["fx"][0].mean()), "fx_mean": float(np.mean(fx)),
                      "optimisations_per_s": round(B / el, 2), "host_syncs": int(st["host_syncs"]),
                      "device_buffers": 5,
                      "problem_evaluations_per_s": round(B * st["evaluations"] / el, 1)}), flush=True)
    ev.close(); v.close()


if __name__ == "__main__":
    a = sys.argv[1:]
    main(B=int(a[0]) if a else 296, model=a[1] if len(a) > 1 else "L96", method=a[2] if len(a) > 2 else "rk2",
         tf=float(a[3]) if len(a) > 3 else 10.0)
