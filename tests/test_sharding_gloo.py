"""
The N > 1 path on the CPU: two gloo ranks shard a batch of independent problems,
evaluate their blocks (with the CPU oracle standing in for the GPU evaluator) and
gather F; the result must equal the single-process answer BIT FOR BIT (same
per-problem arithmetic, no data-path collective).
"""
import os
import socket
import sys

import numpy as np
import pytest
import torch.multiprocessing as mp

from conftest import GOLDEN, ROOT


def test_shard_bounds_cover_and_balance():
    from vgpa_b200.ensemble import shard_bounds
    for total in (1, 7, 8, 32768, 4097):
        for world in (1, 2, 3, 8):
            blocks = [shard_bounds(total, r, world) for r in range(world)]
            assert blocks[0][0] == 0 and blocks[-1][1] == total
            assert all(blocks[i][1] == blocks[i + 1][0] for i in range(world - 1))
            sizes = [b[1] - b[0] for b in blocks]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_bounds(8, 2, 2)


class _OracleEvaluator:
    def __init__(self, probs):
        from oracle import Oracle
        self.orc, self.probs = Oracle(), probs

    def eval(self, X, want_grad=True):
        return self.orc.eval_batch(self.probs, X, want_grad=want_grad, threads=1)


def _problems(total):
    from oracle import Problem
    g = np.load(GOLDEN / "eval_OU_rk4.npz")
    base = Problem.from_golden(g)
    rng = np.random.default_rng(42)
    probs, X = [], []
    for p in range(total):
        probs.append(Problem(model=base.model, method=base.method, D=1, N=base.N, dt=base.dt, theta=g["theta"],
                             sigma=g["sigma"] * (0.8 + 0.05 * p), R=g["R"], obs_t=g["obs_t"],
                             obs_y=g["obs_y"] + 0.05 * rng.standard_normal(g["obs_y"].shape), m0=g["m0"],
                             s0=g["s0"], E0=base.E0, dt_model=base.dt_model))
        X.append(g["x"] * (1.0 + 0.01 * rng.standard_normal(g["x"].size)))
    return probs, np.stack(X)


def _worker(rank, world, port, total, out_dir):
    sys.path.insert(0, str(ROOT))
    sys.path.insert(0, str(ROOT / "tests"))
    import torch.distributed as dist
    from vgpa_b200.ensemble import ShardedEnsemble
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    probs, X = _problems(total)
    ens = ShardedEnsemble(total, lambda lo, hi: _OracleEvaluator(probs[lo:hi]))
    F_all, G_local = ens.eval(X[ens.lo:ens.hi])
    np.savez(os.path.join(out_dir, f"rank{rank}.npz"), F=F_all, G=G_local, lo=ens.lo, hi=ens.hi)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("total", [6, 7])
def test_two_rank_gloo_matches_single_process(tmp_path, total):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_worker, args=(2, port, total, str(tmp_path)), nprocs=2, join=True)
    probs, X = _problems(total)
    F_ref, G_ref = _OracleEvaluator(probs).eval(X)
    for r in range(2):
        z = np.load(tmp_path / f"rank{r}.npz")
        assert np.array_equal(z["F"], F_ref)                       # bit for bit, on every rank
        assert np.array_equal(z["G"], G_ref[int(z["lo"]):int(z["hi"])])


def test_shard_bounds_partition_property():
    """Blocks are contiguous, ordered, cover [0, total) exactly once and differ in size by at most one."""
    from hypothesis import given, settings, strategies as stg
    from vgpa_b200.ensemble import shard_bounds

    @settings(max_examples=200, deadline=None)
    @given(total=stg.integers(0, 100000), world=stg.integers(1, 64))
    def check(total, world):
        blocks = [shard_bounds(total, r, world) for r in range(world)]
        assert blocks[0][0] == 0 and blocks[-1][1] == total
        assert all(blocks[r][1] == blocks[r + 1][0] for r in range(world - 1))
        sizes = [hi - lo for lo, hi in blocks]
        assert min(sizes) >= 0 and max(sizes) - min(sizes) <= 1
        assert sizes == sorted(sizes, reverse=True)

    check()
    with pytest.raises(ValueError):
        shard_bounds(10, 4, 4)
