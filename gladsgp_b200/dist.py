"""Multi-GPU plumbing: one process per GPU, independent units sharded across ranks, one gather.

The path shards by independent unit: MCMC chains (no exchange until the draws are
collected) and test-design blocks for prediction (one all_gather of per-shard moments).  NCCL on
GPUs, gloo in the CPU tests.
"""
import numpy as np


def shard_bounds(n, rank, world):
    """Contiguous [lo, hi) slice of n units owned by `rank` (sizes differ by at most one)."""
    base, rem = divmod(int(n), int(world))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def all_gather_concat(x, dim=0):
    """Concatenate per-rank tensors (equal or ragged along `dim`) on every rank."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return x
    world = dist.get_world_size()
    x = x.movedim(dim, 0).contiguous()
    n = torch.tensor([x.shape[0]], dtype=torch.int64, device=x.device)
    sizes = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(sizes, n)
    sizes = [int(s.item()) for s in sizes]
    mx = max(sizes)
    pad = torch.zeros((mx,) + tuple(x.shape[1:]), dtype=x.dtype, device=x.device)
    pad[:x.shape[0]] = x
    outs = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(outs, pad)
    return torch.cat([o[:s] for o, s in zip(outs, sizes)], dim=0).movedim(0, dim)


def predict_sharded(predictor, xpred):
    """Prediction sharded by test-design block: each rank pushes its slice of designs through its
    copy of the cached factors; means / variances are gathered (B, n)."""
    import torch.distributed as dist
    rank = dist.get_rank() if dist.is_initialized() else 0
    world = dist.get_world_size() if dist.is_initialized() else 1
    lo, hi = shard_bounds(xpred.shape[0], rank, world)
    mean, var = predictor.predict(np.ascontiguousarray(xpred[lo:hi]))
    return all_gather_concat(mean, dim=1), all_gather_concat(var, dim=1)
