"""
Sharding of a batch of independent inference problems over the GPUs of one node.

The path shards naturally: problems are independent, so each rank owns
a contiguous block of problems, evaluates it with its own `BatchEvaluator`, and the
ONLY collective is the gather of the free energies F (B doubles).  Gradients stay on
the rank that produced them (they feed a per-problem optimiser).

Works with any torch.distributed backend: "nccl" on the GPU box, "gloo" in the CPU
tests (tests/test_sharding_gloo.py).
"""
import numpy as np


def shard_bounds(total, rank, world):
    """Contiguous block [lo, hi) of `total` problems owned by `rank` (blocks differ by <= 1)."""
    if not (0 <= rank < world):
        raise ValueError(f"rank {rank} outside world of {world}")
    base, extra = divmod(int(total), int(world))
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def gather_free_energies(F_local, total, group=None):
    """All ranks receive the full F vector (length `total`), ordered by global problem index.
    F_local: 1-D float64 torch tensor (on the backend's device) or numpy array."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return np.asarray(F_local.cpu() if hasattr(F_local, "cpu") else F_local, dtype=np.float64).copy()
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    t = F_local if isinstance(F_local, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(F_local))
    if dist.get_backend(group) == "nccl" and not t.is_cuda:
        t = t.cuda()                      # NCCL moves device memory only (current CUDA device of this rank)
    sizes = [shard_bounds(total, r, world)[1] - shard_bounds(total, r, world)[0] for r in range(world)]
    assert t.numel() == sizes[rank], (t.numel(), sizes[rank])
    if len(set(sizes)) == 1:
        out = torch.empty(total, dtype=t.dtype, device=t.device)
        dist.all_gather_into_tensor(out, t.contiguous(), group=group)
    else:   # uneven blocks: pad to the largest, gather, trim
        width = max(sizes)
        padded = torch.zeros(width, dtype=t.dtype, device=t.device)
        padded[:t.numel()] = t
        buf = torch.empty(width * world, dtype=t.dtype, device=t.device)
        dist.all_gather_into_tensor(buf, padded, group=group)
        out = torch.cat([buf[r * width:r * width + sizes[r]] for r in range(world)])
    return out.cpu().numpy()


class ShardedEnsemble:
    """The local shard of an ensemble of `total` problems.

    `make_evaluator(lo, hi)` must return an object with `.eval(X, want_grad)` for the
    problems [lo, hi) -- a `vgpa_b200.BatchEvaluator` in production."""

    def __init__(self, total, make_evaluator, rank=None, world=None, group=None):
        import torch.distributed as dist
        if rank is None or world is None:
            if dist.is_available() and dist.is_initialized():
                rank, world = dist.get_rank(group), dist.get_world_size(group)
            else:
                rank, world = 0, 1
        self.total, self.rank, self.world, self.group = int(total), rank, world, group
        self.lo, self.hi = shard_bounds(total, rank, world)
        self.evaluator = make_evaluator(self.lo, self.hi)

    def eval(self, X_local, want_grad=True):
        """X_local: the rows [lo, hi) of the global X (host).  Returns (F_all (total,), grad_local)."""
        F_local, G_local = self.evaluator.eval(X_local, want_grad)
        return gather_free_energies(np.asarray(F_local), self.total, self.group), G_local

    def eval_device(self, X, F, G=None, stream=0):
        """Device-resident shard: X (rows, n_x), F (rows,), G (rows, n_x) or None are CUDA float64
        torch tensors of this rank.  Asynchronous on `stream` (0 = the legacy default stream); call
        `gather_device` for the one collective of the path."""
        self.evaluator.eval_device(X.data_ptr(), X.stride(0) if X.dim() == 2 else 0, F.data_ptr(),
                                   None if G is None else G.data_ptr(), None if G is None else G.stride(0), stream)

    def gather_device(self, F):
        """Wait for the shard's evaluation and gather F over the ranks (NCCL over NVLink): returns the
        full (total,) CUDA tensor on every rank, or F itself without a process group."""
        import torch
        import torch.distributed as dist
        self.evaluator.sync()
        if not (dist.is_available() and dist.is_initialized()) or self.world == 1:
            return F
        sizes = [shard_bounds(self.total, r, self.world)[1] - shard_bounds(self.total, r, self.world)[0]
                 for r in range(self.world)]
        if len(set(sizes)) == 1:
            out = torch.empty(self.total, dtype=F.dtype, device=F.device)
            dist.all_gather_into_tensor(out, F.contiguous(), group=self.group)
            return out
        return torch.from_numpy(gather_free_energies(F, self.total, self.group)).to(F.device)
