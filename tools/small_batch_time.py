"""Time per batch evaluation of the small models (device-resident x, CUDA events) for a list of
(model, method, problems, N).  Usage: python tools/small_batch_time.py [model method B N] ..."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vgpa_b200.engine import BatchEvaluator  # noqa: E402


def run(model, method, B, N, steps=20):
    D = 3 if model == "L63" else 1
    rng = np.random.default_rng(17)
    M = max(1, N // 50)
    obs_t = np.linspace(0, N, M + 2, dtype=int)[1:-1]
    theta = [10.0, 28.0, 2.6667] if model == "L63" else ([2.0] if model == "OU" else [1.0])
    sig = np.full(D, 10.0 if model == "L63" else 0.8)
    R = np.full(D, 2.0 if model == "L63" else 0.04)
    obs_y = rng.standard_normal((B, M, D)) * (3.0 if model == "L63" else 0.5)
    m0 = rng.standard_normal((B, D))
    s0 = 0.25 * np.eye(D)
    x1 = np.concatenate([np.tile(np.diag(0.5 * sig / 0.25).ravel(), N), np.zeros(N * D)])
    X = torch.from_numpy(x1[None, :] + 0.05 * rng.standard_normal((B, x1.size))).cuda()
    G = torch.empty_like(X)
    F = torch.empty(B, dtype=torch.float64, device="cuda")
    with BatchEvaluator(model, method, N, 0.01, theta, sig, R, obs_t, obs_y, m0, s0, np.zeros(B), B=B) as ev:
        st = torch.cuda.current_stream().cuda_stream
        nx = ev.n_x
        for _ in range(3):
            ev.eval_device(X.data_ptr(), nx, F.data_ptr(), G.data_ptr(), nx, st)
        ev.sync()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            ev.eval_device(X.data_ptr(), nx, F.data_ptr(), G.data_ptr(), nx, st)
        e1.record()
        ev.sync()
    ms = e0.elapsed_time(e1) / steps
    print("%-4s %-5s B=%-6d N=%-5d %8.4f ms/eval  %10.0f evals/s" % (model, method, B, N, ms, B / ms * 1e3), flush=True)


if __name__ == "__main__":
    a = sys.argv[1:]
    cases = [(a[i], a[i + 1], int(a[i + 2]), int(a[i + 3])) for i in range(0, len(a), 4)] or [
        ("OU", "rk4", 1024, 1001), ("OU", "rk4", 1, 1001), ("DW", "euler", 1, 1001), ("OU", "rk2", 1024, 1001),
        ("OU", "heun", 1024, 1001), ("OU", "rk4", 4096, 1001), ("OU", "rk4", 8192, 1001), ("OU", "rk4", 8193, 1001),
        ("OU", "rk4", 65536, 1001)]
    for c in cases:
        run(*c)
