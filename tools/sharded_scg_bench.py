"""A whole Lorenz-96 D=40 T=1000 ensemble OPTIMISED on one GPU (or one rank's share under torchrun) in resident
sub-batches: vgpa_b200.batched_scg.ShardedBatchedSCG with the sub-batch sized from the free HBM.
    python tools/sharded_scg_bench.py [problems [sub_batch (0 = from the free HBM) [concurrent sub-batches]]]"""
import json, sys, time
import numpy as np
sys.path.insert(0, ".")
import bench
import torch
from vgpa_b200.batched_scg import ShardedBatchedSCG
from vgpa_b200.engine import BatchEvaluator


def main(total=2000, sub_batch=None, concurrent=1):
    fam = bench.l96_problem_family(0)
    iset, arr = bench.shard_arrays(fam, total)
    D, N = bench.D, bench.N_GRID

    def make(lo, hi):
        return BatchEvaluator("L96", "rk2", N, bench.DT, [8.0], arr["sigma"][lo:hi], np.ones(D), fam["obs_t"],
                              arr["obs_y"][lo:hi], arr["m0"][lo:hi], fam["s0"], arr["E0"][lo:hi], B=hi - lo,
                              dt_model=fam["dt_model"], device=0)
    ens = ShardedBatchedSCG(total, make, {"max_it": 500, "x_tol": 1.0e-6, "f_tol": 1.0e-8, "display": False},
                            sub_batch=sub_batch)
    free0 = torch.cuda.mem_get_info(0)[0]
    t0 = time.perf_counter()
    res = ens.run(keep=(0, total - 1), concurrent=concurrent)
    el = time.perf_counter() - t0
    out = ens.save("gpurun_out/ensemble_l96", N, D)
    print(json.dumps({"problems": total, "seconds": round(el, 2), "optimisations_per_s": round(total / el, 1),
                      "sub_batch": int(res["sub_batch"]), "concurrent": concurrent, "free_hbm_gb_before": round(free0 / 1e9, 1),
                      "iterations_min_median_max": [int(res["n_it"].min()), int(np.median(res["n_it"])), int(res["n_it"].max())],
                      "fx_mean": float(res["fx"].mean()), "file": str(out), "peak_torch_gb": round(torch.cuda.max_memory_allocated(0) / 1e9, 1)}))


if __name__ == "__main__":
    main(int(sys.argv[1]) if len(sys.argv) > 1 else 2000, int(sys.argv[2]) if len(sys.argv) > 2 and int(sys.argv[2]) > 0 else None,
         int(sys.argv[3]) if len(sys.argv) > 3 else 1)
