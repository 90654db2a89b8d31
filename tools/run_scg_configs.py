"""BASELINE configs through the reference-facing interface on the GPU: full SCG runs."""
import sys, time, json
import numpy as np
sys.path.insert(0, "."); sys.path.insert(0, "tests/golden")
import make_golden as mg
from vgpa_b200 import Simulation, SCG

def run(model, method, tf, max_it):
    sim = Simulation("cfg"); sim.setup(mg.config(model, method, tf))
    v = sim.build(); x0 = v.initialization()
    scg = SCG(v.free_energy, v.gradient, {"max_it": max_it, "x_tol": 1e-6, "f_tol": 1e-8, "display": False})
    t0 = time.perf_counter(); x, fx = scg(x0.copy()); el = time.perf_counter() - t0
    st = scg.stats; n = int(st["MaxIt"])
    print(json.dumps({"model": model, "method": method, "N": v.dim_n, "iterations": n, "fx": fx, "fx0": float(st["fx"][0]),
                      "f_eval": st["f_eval"], "cuda_evals": v.n_eval, "seconds": round(el, 2),
                      "ms_per_cuda_eval": round(1e3 * el / max(v.n_eval, 1), 2)}), flush=True)
    v.close()

if __name__ == "__main__":
    run("DW", "euler", 10.0, 500)
    run("OU", "rk4", 10.0, 500)
    run("L63", "heun", 20.0, 500)
    run("L96", "rk2", 10.0, 500)
