"""bench.py's reference arm (the CPU port of the reference algorithm on the host cores) runs without a
GPU: check the JSON line it prints against the contract the driver reads, and that under a multi-rank
launch only rank 0 works."""
import json
import os
import subprocess
import sys

from conftest import ROOT

KEYS = {"impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better",
        "scaling", "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e", "gpu_launches"}


def run(env_extra, *args):
    env = dict(os.environ, **env_extra)
    return subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", *args], cwd=ROOT, env=env,
                          capture_output=True, text=True, timeout=600)


def test_reference_arm_line():
    r = run({}, "--steps", "1", "--warmup", "0", "--scg-problems", "0")      # (the SCG sample alone is a minute of CPU)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert KEYS <= set(d)
    assert d["impl"] == "reference" and d["unit"] == "evals/s" and d["higher_is_better"] is True
    assert d["dtype"] == "f64" and d["vs_baseline"] is None and d["gpu_launches"] == 0
    assert d["value"] > 0 and d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"] == {"value": d["value"], "unit": "evals/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and "L96" in d["config"]["workload"]
    # what the arm times is said at the top level, and the unmodified Python reference is timed beside it
    assert d["reference_kind"] == "port"
    cal = d["cpu_baseline_reference"]
    assert cal["kind"] == "reference"
    if "unavailable" not in cal:          # baseline/_ref + numba present: the reference's known answer at x0
        assert abs(cal["F_x0"] - 49769.73517670352) < 1e-6 and 0 < cal["value"] < d["value"]
    # the two arms must agree on what they measure
    src = (ROOT / "bench.py").read_text()
    assert src.count('"metric": METRIC') == 2 and src.count('"workload": WORKLOAD') == 2


def test_reference_arm_other_ranks_do_nothing():
    r = run({"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"}, "--gpus", "2", "--steps", "1", "--warmup", "0")
    assert r.returncode == 0 and r.stdout.strip() == ""
