"""Quick device-resident timing of the L96 D=40 batch evaluation (development aid)."""
import sys, time, json
import numpy as np
import torch
sys.path.insert(0, '.')
from vgpa_b200.engine import BatchEvaluator

def main(B=296, N=1001, method="rk2", steps=3, model="L96"):
    method = {"OU": "rk4", "DW": "euler"}.get(model, method)
    D = {"L96": 40, "L63": 3, "OU": 1, "DW": 1}[model]
    rng = np.random.default_rng(0)
    M = 80 if model == "L96" else 20
    obs_t = np.linspace(0, N, M + 2, dtype=int)[1:-1]
    sig = {"L96": 4.0, "L63": 10.0, "OU": 0.8, "DW": 0.8}[model]
    theta = {"L96": [8.0], "L63": [10.0, 28.0, 2.6667], "OU": [2.0], "DW": [1.0]}[model]
    ev = BatchEvaluator(model, method, N, 0.01, theta, np.full(D, sig), np.full(D, 1.0), obs_t,
                        rng.standard_normal((M, D)), rng.standard_normal(D), 0.2 * np.eye(D), 0.0, B=B)
    n = ev.n_x
    x1 = np.concatenate([np.tile((2 * sig * np.eye(D)).ravel(), N) + 0.1 * rng.standard_normal(N * D * D),
                         rng.standard_normal(N * D)])
    X = torch.from_numpy(x1).cuda().repeat(B, 1).contiguous()
    X += 0.01 * torch.randn_like(X)
    F = torch.empty(B, dtype=torch.float64, device="cuda")
    G = torch.empty_like(X)
    st = torch.cuda.current_stream().cuda_stream
    for want_grad in (0, 1):
        for _ in range(2):
            ev.eval_device(X.data_ptr(), n, F.data_ptr(), G.data_ptr() if want_grad else None, n, st)
        ev.sync()
        ev.set_timing(True); ev.get_timing()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            ev.eval_device(X.data_ptr(), n, F.data_ptr(), G.data_ptr() if want_grad else None, n, st)
        e1.record()
        ev.sync()
        ms = e0.elapsed_time(e1) / steps
        tm = {k: round(v[0] / max(v[1], 1), 3) for k, v in ev.get_timing().items()}
        ev.set_timing(False)
        print(json.dumps({"model": model, "B": B, "N": N, "method": method, "want_grad": want_grad,
                          "ms": round(ms, 3), "evals_per_s": round(B / ms * 1e3, 1),
                          "chunk": ev.chunk_size, "kernel_ms": tm, "F0": float(F[0])}))

if __name__ == "__main__":
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 296
    model = sys.argv[2] if len(sys.argv) > 2 else "L96"
    N = int(sys.argv[3]) if len(sys.argv) > 3 else 1001
    main(B=B, model=model, N=N)
