"""Latency of ONE free_energy(x) + gradient(x) pair through the reference-facing VarGP
(host numpy in, host numpy out), as the SCG issues it: BASELINE configs[3] shape by default."""
import sys, time, json
import numpy as np
sys.path.insert(0, "."); sys.path.insert(0, "tests/golden")
import make_golden as mg
from vgpa_b200 import Simulation

def run(model="L96", method="rk2", tf=10.0, reps=20):
    sim = Simulation("cfg"); sim.setup(mg.config(model, method, tf))
    v = sim.build(); x0 = v.initialization()
    rng = np.random.default_rng(1)
    xs = [x0 * (1.0 + 1e-3 * rng.uniform(-1, 1, x0.size)) for _ in range(4)]
    for x in xs[:2]:
        v.free_energy(x); v.gradient(x)
    ts = []
    for i in range(reps):
        x = xs[i % 4]
        t0 = time.perf_counter(); f = v.free_energy(x); g = v.gradient(x); ts.append(time.perf_counter() - t0)
    ev = v._ev
    ev.set_timing(True); ev.get_timing()
    v.free_energy(xs[0] * 1.0)
    tm = {k: round(t[0] / max(t[1], 1), 3) for k, t in ev.get_timing().items()}
    ev.set_timing(False)
    print(json.dumps({"model": model, "method": method, "N": v.dim_n, "n_x": int(x0.size),
                      "pair_ms_median": round(1e3 * float(np.median(ts)), 3), "pair_ms_min": round(1e3 * min(ts), 3),
                      "kernel_ms": tm, "F": f}), flush=True)
    v.close()

if __name__ == "__main__":
    run(*(sys.argv[1:3] or ["L96", "rk2"]), tf=float(sys.argv[3]) if len(sys.argv) > 3 else 10.0)
