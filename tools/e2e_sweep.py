"""e2e throughput (host pinned buffers through vgpa_eval) vs chunk size."""
import sys, time, json
import numpy as np
sys.path.insert(0, ".")
import bench
from vgpa_b200.engine import BatchEvaluator
from vgpa_b200._lib import PinnedArray
fam = bench.l96_problem_family(0)
Be = 256
iset, arr = bench.shard_arrays(fam, Be)
xe = PinnedArray((Be, bench.N_X)); ge = PinnedArray((Be, bench.N_X))
rng = np.random.default_rng(0)
for p in range(Be):
    xe.array[p] = fam["sets"][iset[p]]["x0"] * (1 + 0.02 * rng.uniform(-1, 1, bench.N_X))
per = 8 * bench.N_GRID * (2 * 40 + 2 * 1600 + 1)
for chunk in (32, 64, 128, 148, 256):
    ev = BatchEvaluator("L96", "rk2", bench.N_GRID, bench.DT, [8.0], arr["sigma"], np.ones(40), fam["obs_t"], arr["obs_y"],
                        arr["m0"], fam["s0"], arr["E0"], B=Be, dt_model=fam["dt_model"], scratch_bytes=chunk * per + 1024)
    Fe = np.empty(Be)
    for _ in range(2): ev.eval(xe.array, True, Fe, ge.array)
    t0 = time.perf_counter()
    for _ in range(3): ev.eval(xe.array, True, Fe, ge.array)
    el = (time.perf_counter() - t0) / 3
    print(json.dumps({"chunk": ev.chunk_size, "ms": round(el * 1e3, 1), "evals_per_s": round(Be / el, 1),
                      "GBps_each_way": round(Be * bench.N_X * 8 / el / 1e9, 1)}), flush=True)
    ev.close()
