#!/usr/bin/env python
"""
bench.py -- VGPA free-energy + gradient evaluations per second on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port P bench.py --gpus N --steps K --warmup W

Workload (BASELINE.json configs[4], the one the metric is quoted on): the Lorenz-96
D=40, T=1000 (N=1001 grid points, RK2) ensemble of 32768 independent inference
problems = 64 observation sets x 32 starts x 16 system-noise values, sharded as
contiguous blocks of 4096 problems per GPU (weak scaling: per-GPU work is fixed;
8 GPUs = the full 32768).  One "step" = free_energy + gradient for every problem of
the shard, x and grad resident in HBM.  NCCL is used only to gather F.

JSON keys follow the driver contract; see DESIGN.md section "Measurement".
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

SEED = 31415926535
D, N_GRID, M_OBS, DT = 40, 1001, 80, 0.01
N_X = N_GRID * D * (D + 1)
PER_GPU = 4096
N_OBS_SETS_PER_GPU, N_STARTS, N_NOISE = 8, 32, 16
# algorithmic work per (problem, time index), SURVEY.md 8(d)
FLOP_STEP = {"fwd": 262400.0, "energy": 429867.0, "bwd": 262400.0 + 265600.0}
BYTE_STEP = {"fwd": 2 * 8.0 * D * (D + 1), "energy": 3 * 8.0 * D * (D + 1), "bwd": 4 * 8.0 * D * (D + 1)}
FLOP_EVAL = 1.2215e9
BYTE_EVAL = 9 * 8.0 * N_GRID * D * (D + 1)
# FP64 peaks measured on this pool's B200 by tools/microbench.cu (profiles/microbench_r01.jsonl)
FP64_DFMA_TFLOPS, FP64_DMMA_TFLOPS = 33.9, 37.1


def l96_problem_family(rank):
    """The shard of the C5 ensemble owned by `rank`: 8 observation sets x 32 starts x 16
    noise values.  Host side: one sample path, per-set observations and the reference's
    cubic-spline initialisation (VarGP.initialization) -- identical code to the single
    problem path."""
    from vgpa_b200.simulation import Simulation
    params = {"Output_Name": "bench", "Model": "L96", "Ode-method": "RK2", "Random-Seed": SEED,
              "Time-window": {"t0": 0.0, "tf": 10.0, "dt": DT}, "Noise": {"sys": [4.0] * D, "obs": 1.0},
              "Observations": {"density": 8, "operator": None}, "Drift": {"theta": 8.0},
              "Prior": {"tau0": 0.5, "mu0": 1.0}}
    sim = Simulation("bench")
    sim.setup(params)
    md = sim.m_data
    path = md["model"].sample_path
    obs_t = np.asarray(md["obs_t"], dtype=np.int64)
    sets = []
    for s in range(N_OBS_SETS_PER_GPU):
        g = rank * N_OBS_SETS_PER_GPU + s
        rng = np.random.default_rng(np.random.SeedSequence([SEED, g]))
        obs_y = path[obs_t] + rng.standard_normal((obs_t.size, D))          # R = 1
        m0 = path[0] + 0.1 * rng.standard_normal(D)
        md["obs_y"], md["m0"] = obs_y, m0
        vg = sim.build()
        sets.append(dict(obs_y=obs_y, m0=m0, x0=vg.initialization(),
                         E0=float(vg.kl0(m0, md["s0"]))))
    noise = np.array([4.0 * 2.0 ** ((j - 8) / 8.0) for j in range(N_NOISE)])
    return dict(obs_t=obs_t, sets=sets, noise=noise, s0=md["s0"], dt_model=float(md["model"].time_step))


def shard_arrays(fam, count):
    """Per-problem parameter arrays for `count` problems of the shard, index
    p = (set * 32 + start) * 16 + noise."""
    idx = np.arange(count)
    iset = idx // (N_STARTS * N_NOISE)
    inoise = idx % N_NOISE
    obs_y = np.stack([fam["sets"][s]["obs_y"] for s in iset])
    m0 = np.stack([fam["sets"][s]["m0"] for s in iset])
    E0 = np.array([fam["sets"][s]["E0"] for s in iset])
    sigma = np.repeat(fam["noise"][inoise][:, None], D, axis=1)
    return iset, dict(obs_y=obs_y, m0=m0, E0=E0, sigma=sigma)


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self.stop_flag = index, [], threading.Event()

    def run(self):
        while not self.stop_flag.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True,
                                     timeout=5).stdout.strip()
                if out:
                    self.rows.append([c.strip() for c in out.split(",")])
            except Exception:
                pass
            self.stop_flag.wait(0.2)

    def summary(self):
        self.stop_flag.set()
        self.join(timeout=6)
        sm = [float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(self.rows)}


def cpu_port_rate(threads, problems, fam=None):
    """The oracle (C port of the reference algorithm) on the host cores: `problems`
    evaluations of the L96 N=1001 problem, one OpenMP thread per problem."""
    from oracle import Oracle, Problem
    fam = fam or l96_problem_family(0)
    s = fam["sets"][0]
    orc = Oracle()
    probs = [Problem(model="L96", method="rk2", D=D, N=N_GRID, dt=DT, theta=np.array([8.0]),
                     sigma=np.full(D, 4.0), R=np.ones(D), obs_t=fam["obs_t"], obs_y=s["obs_y"], m0=s["m0"],
                     s0=fam["s0"], E0=s["E0"], dt_model=fam["dt_model"]) for _ in range(problems)]
    rng = np.random.default_rng(5)
    X = np.stack([s["x0"] * (1.0 + 0.02 * rng.uniform(-1, 1, N_X)) for _ in range(problems)])
    t0 = time.perf_counter()
    F, G = orc.eval_batch(probs, X, want_grad=True, threads=threads)
    el = time.perf_counter() - t0
    assert np.all(np.isfinite(F))
    return problems / el, el


METRIC = "free-energy+grad evals/sec (L96 D=40, T=1000); batched problems/sec"
WORKLOAD = ("L96 D=40 N=1001 (T=1000) RK2 ensemble: 8 obs sets x 32 starts x 16 noise values per GPU "
            "(BASELINE configs[4])")


def run_reference(args):
    """--impl reference: the reference's CPU algorithm (the oracle port; the reference is
    Python and cannot travel to the box) on all host cores, same metric/config."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = len(os.sched_getaffinity(0))
    fam = l96_problem_family(0)
    per_step = 4 * max(threads, 1)     # several problems per thread: the step is not paced by one slow thread
    for _ in range(args.warmup):
        cpu_port_rate(threads, min(per_step, 2 * threads), fam)
    n, el = 0, 0.0
    for _ in range(args.steps):
        el += cpu_port_rate(threads, per_step, fam)[1]
        n += per_step
    val = n / el
    line = {"impl": "reference", "metric": METRIC, "value": val,
            "unit": "evals/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * el / max(args.steps, 1), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOAD, "problems_per_step": per_step},
            "cpu_baseline": {"value": val, "unit": "evals/s", "cores": threads, "kind": "port",
                             "sample": f"{per_step} problems per step on {threads} OpenMP threads (one problem per thread at a time)"},
            "e2e": {"value": val, "unit": "evals/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def run_b200(args):
    import torch
    import torch.distributed as dist
    from vgpa_b200.engine import BatchEvaluator
    from vgpa_b200._lib import PinnedArray

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl b200 needs a CUDA device: there is no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    B = args.per_gpu
    fam = l96_problem_family(rank)
    iset, arr = shard_arrays(fam, B)
    ev = BatchEvaluator("L96", "rk2", N_GRID, DT, [8.0], arr["sigma"], np.ones(D), fam["obs_t"], arr["obs_y"],
                        arr["m0"], fam["s0"], arr["E0"], B=B, dt_model=fam["dt_model"], device=local)
    # x^p = x0^{set(p)} * (1 + 0.02 u_p), u ~ U(-1, 1): generated in HBM (synthetic data)
    X = torch.empty((B, N_X), dtype=torch.float64, device=dev)
    x0s = torch.from_numpy(np.stack([s["x0"] for s in fam["sets"]])).to(dev)
    gen = torch.Generator(device=dev)
    gen.manual_seed(SEED % (2 ** 31) + rank)
    blk = 128
    for p0 in range(0, B, blk):
        p1 = min(B, p0 + blk)
        u = torch.rand((p1 - p0, N_X), dtype=torch.float64, device=dev, generator=gen) * 2.0 - 1.0
        X[p0:p1] = x0s[torch.from_numpy(iset[p0:p1]).to(dev)] * (1.0 + 0.02 * u)
        del u
    G = torch.empty_like(X)
    F = torch.empty(B, dtype=torch.float64, device=dev)
    F_all = torch.empty(B * world, dtype=torch.float64, device=dev) if world > 1 else None
    stream = torch.cuda.current_stream().cuda_stream

    def step():
        ev.eval_device(X.data_ptr(), N_X, F.data_ptr(), G.data_ptr(), N_X, stream)
        if world > 1:
            dist.all_gather_into_tensor(F_all, F)     # the only collective: gather of F

    def fence():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    for _ in range(args.warmup):
        step()
    ev.sync()
    fence()
    ev.set_timing(True)
    ev.get_timing()
    sampler = ClockSampler(local)
    sampler.start()
    launches0 = ev.launch_count
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    fence()
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    fence()
    ev.sync()
    ms = e0.elapsed_time(e1)
    clocks = sampler.summary()
    launches = ev.launch_count - launches0 + (args.steps if world > 1 else 0)
    timing = ev.get_timing()
    ev.set_timing(False)
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    F_host = F.cpu().numpy()
    assert np.all(np.isfinite(F_host)), "non-finite free energy in the bench batch"
    value = B * world * args.steps / (ms * 1e-3)

    line = None
    if rank == 0:
        # ---- roofline of the dominant kernel (largest share of the step) -------------
        kinds = {k: v for k, v in timing.items() if k != "finalize" and v[1] > 0}
        dom = max(kinds, key=lambda k: kinds[k][0])
        tot_ms = sum(v[0] for v in timing.values())
        launches_dom = kinds[dom][1]
        avg_ms = kinds[dom][0] / launches_dom
        units_per_launch = B * N_GRID * args.steps / launches_dom   # (problem, time index) pairs
        tflops = FLOP_STEP[dom] * units_per_launch / (avg_ms * 1e-3) / 1e12
        gbs = BYTE_STEP[dom] * units_per_launch / (avg_ms * 1e-3) / 1e9
        peaks = json.loads((ROOT / "MEASURED_PEAKS.json").read_text()) if (ROOT / "MEASURED_PEAKS.json").exists() else {}
        hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
        traffic = None
        tpath = ROOT / "profiles" / "traffic_r01.json"
        if tpath.exists():
            # the capture is of one 888-problem launch; `achieved` is per AVERAGE launch of the timed region
            # (4096 = 4 x 888 + 544 problems per step), so the measured bytes are scaled to the same units
            traffic = json.loads(tpath.read_text()).get(dom)
            if traffic is not None:
                traffic = float(traffic) * units_per_launch / (888.0 * N_GRID)
        roofline = {"kernel": {"fwd": "l96_fwd_kernel", "energy": "l96_energy_kernel", "bwd": "l96_bwd_kernel"}[dom],
                    "bound": "tensor", "achieved": tflops, "peak": FP64_DMMA_TFLOPS, "unit": "TFLOP/s",
                    "frac": tflops / FP64_DMMA_TFLOPS, "traffic": traffic,
                    "peak_source": "FP64 DMMA peak measured by tools/microbench.cu on this pool "
                                   "(MEASURED_PEAKS.json has no FP64 figure)",
                    "share_of_step": kinds[dom][0] / tot_ms,
                    "hbm": {"achieved": gbs, "peak": hbm_peak, "unit": "GB/s", "frac": gbs / hbm_peak,
                            "peak_source": "MEASURED_PEAKS.json" if peaks else "fallback"},
                    "kernel_ms": {k: round(v[0] / max(v[1], 1), 4) for k, v in timing.items()},
                    "whole_eval": {"tflops": value / world * FLOP_EVAL / 1e12,
                                   "frac_fp64": value / world * FLOP_EVAL / 1e12 / FP64_DMMA_TFLOPS,
                                   "gbs": value / world * BYTE_EVAL / 1e9,
                                   "frac_hbm": value / world * BYTE_EVAL / 1e9 / hbm_peak}}
        line = {"metric": METRIC,
                "value": value, "unit": "evals/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f64", "data": "synthetic",
                "config": {"workload": WORKLOAD,
                           "problems_per_gpu": B, "global_problems": B * world, "parallelism": f"dp{world}",
                           "l2": "inputs larger than L2 (x shard = %.1f GB)" % (B * N_X * 8 / 1e9),
                           "chunk": ev.chunk_size},
                "clocks": clocks, "gpu_launches": int(launches), "roofline": roofline}

    # ---- e2e: the same evaluation through the public host API with HOST buffers ----
    # 1024 problems per step (27 GB of pinned host memory per rank) when the box has the memory for
    # every rank, else 512: the pipeline's fill / drain (first H2D, last kernels + D2H, ~27 ms) is a
    # fixed cost per call, so the longer step measures the steady state of the API more closely
    Be = args.e2e_problems
    if Be <= 0:
        avail_gb = 0.0
        try:
            with open("/proc/meminfo") as fh:
                for ln in fh:
                    if ln.startswith("MemAvailable:"):
                        avail_gb = float(ln.split()[1]) / 1e6
        except OSError:
            pass
        Be = 1024 if avail_gb >= 64.0 * world else 512
        if world > 1:   # every rank must use the same step size: take the smallest choice
            tb = torch.tensor([Be], dtype=torch.int64, device=dev)
            dist.all_reduce(tb, op=dist.ReduceOp.MIN)
            Be = int(tb.item())
    Be = min(Be, B)
    xe = PinnedArray((Be, N_X))
    ge = PinnedArray((Be, N_X))
    xe.array[:] = X[:Be].cpu().numpy()
    ev.close()
    del X, G
    torch.cuda.empty_cache()
    ev_e = BatchEvaluator("L96", "rk2", N_GRID, DT, [8.0], arr["sigma"][:Be], np.ones(D), fam["obs_t"],
                          arr["obs_y"][:Be], arr["m0"][:Be], fam["s0"], arr["E0"][:Be], B=Be,
                          dt_model=fam["dt_model"], device=local,
                          # 64-problem chunks: H2D of chunk c+1, kernels of chunk c and D2H of chunk c-1 overlap
                          scratch_bytes=64 * 8 * N_GRID * (2 * D + 2 * D * D + 1) + 1024)
    Fe = np.empty(Be)
    for _ in range(2):
        ev_e.eval(xe.array, True, Fe, ge.array)
    fence()
    t0 = time.perf_counter()
    for _ in range(args.e2e_steps):
        ev_e.eval(xe.array, True, Fe, ge.array)
    torch.cuda.synchronize()
    el = time.perf_counter() - t0
    te = torch.tensor([el], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_val = Be * world * args.e2e_steps / float(te.item())
    assert np.allclose(Fe, F_host[:Be], rtol=1e-12, atol=0.0), "host-API result differs from device-API result"
    ev_e.close()
    xe.free()
    ge.free()

    if rank == 0:
        line["e2e"] = {"value": e2e_val, "unit": "evals/s", "h2d_bytes_per_step": int(Be * N_X * 8),
                       "d2h_bytes_per_step": int(Be * N_X * 8 + Be * 8),
                       "sample": f"{Be} problems per GPU per step through vgpa_eval with pinned host x/grad, "
                                 f"{args.e2e_steps} steps"}
        if world == 1:
            threads = len(os.sched_getaffinity(0))
            probs = max(threads, 1) * (12 if threads <= 16 else (6 if threads <= 64 else 3))   # ~10-30 s of CPU work
            rate, el_cpu = cpu_port_rate(threads, probs, fam)
            line["cpu_baseline"] = {"value": rate, "unit": "evals/s", "cores": threads, "kind": "port",
                                    "sample": f"{probs} L96 N=1001 problems, one OpenMP thread each, "
                                              f"{el_cpu:.1f} s of wall time (oracle/vgpa_oracle.c)"}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--per-gpu", type=int, default=PER_GPU, help="problems per GPU (default 4096)")
    ap.add_argument("--e2e-problems", type=int, default=0, help="problems per e2e step (0 = 1024 if host memory allows, else 512)")
    ap.add_argument("--e2e-steps", type=int, default=3)
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
