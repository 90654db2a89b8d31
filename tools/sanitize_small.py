"""Small end-to-end evaluations for compute-sanitizer (one tool per gpurun call)."""
import sys
import numpy as np
sys.path.insert(0, ".")
sys.path.insert(0, "tests")
from test_gpu_parity import evaluator_from_golden
for name in ("eval_L96_rk2", "eval_L96_rk4", "eval_L96_euler", "eval_L63_heun", "eval_DW_euler", "eval_OU_rk4"):
    g = np.load(f"tests/golden/{name}.npz")
    with evaluator_from_golden(g, B=3) as ev:
        X = np.stack([g["x"]] * 3)
        F, G = ev.eval(X)
        full = ev.eval_full(g["x"], problem=1)
    print(name, abs(F[0] - float(g["F"])) / abs(float(g["F"])), np.abs(G[0] - g["grad"]).max() / np.abs(g["grad"]).max())
