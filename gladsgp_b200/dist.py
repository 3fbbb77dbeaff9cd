"""Multi-GPU plumbing: one process per GPU, independent units sharded across ranks, one gather.

The path shards by independent unit: MCMC chains (no exchange until the draws are
collected), the PCs of one chain (one all_gather of per-PC result rows per step: mcmc_by_pc), test-design
blocks for prediction (one all_gather of per-shard moments) and output-column slabs of the ensemble for the
rSVD.  NCCL on GPUs, gloo in the CPU tests.
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


def randomized_svd_sharded(X_slab, p, k=None, q=1, omega_slab=None, products=None):
    """src/svd.py:12-82 with the ensemble sharded by output-column slab: rank g holds
    X[:, c0:c1] (m, n_local) on its GPU and streams only that slab.

      Y = sum_g X_g Omega_g              -> all_reduce of an (m, r) matrix            (svd.py:52)
      Y = sum_g X_g (X_g^T Y)            -> one all_reduce per power iteration        (svd.py:55-56)
      Q = qr(Y)                          -> replicated (same Y on every rank)         (svd.py:59)
      B_g = Q^T X_g                      -> stays sharded: it is Vh sharded           (svd.py:60)
      B B^T = sum_g B_g B_g^T            -> all_reduce of an (r, r) matrix            (svd.py:63)

    Returns (U (m, p), S (p,), Vh_slab (p, n_local)) as tensors on X_slab's device.  `omega_slab` is the block of
    rows of the (n, r) Gaussian test matrix that belongs to this slab; to reproduce a single-GPU run every rank
    draws the full matrix from the same seeded np.random stream and slices it (src/svd.py:51 draws it from the
    global stream); a torch tensor already on the device is used as it is.  `products` = (sketch(X, omegaT) -> X @ omegaT.T, xty(X, Y) -> Y.T @ X); default: the
    tensor-core kernels (ops.rsvd_sketch_tc / ops.rsvd_xty_tc)."""
    import torch
    import torch.distributed as dist
    if k is None:
        k = p
    r = p + k
    if products is None:
        from . import ops
        products = (ops.rsvd_sketch_tc, ops.rsvd_xty_tc)
    sketch, xty = products
    multi = dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1

    def allsum(t):
        if multi:
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return t

    n_local = X_slab.shape[1]
    if omega_slab is None:
        raise ValueError('omega_slab (n_local, r) is required: slice it from the rank-consistent test matrix')
    if torch.is_tensor(omega_slab):                           # already on the device: transpose there
        omT = omega_slab.to(device=X_slab.device, dtype=torch.float32).T.contiguous()
    else:
        omT = torch.as_tensor(np.ascontiguousarray(np.asarray(omega_slab, dtype=np.float32).T), device=X_slab.device)
    if omT.shape != (r, n_local):
        raise ValueError('omega_slab must have shape (n_local, p + k)')
    Y = allsum(sketch(X_slab, omT))
    for _ in range(q):
        Zt = xty(X_slab, Y)
        Y = allsum(sketch(X_slab, Zt.contiguous()))
    Qd, _ = torch.linalg.qr(Y.double(), mode='reduced')       # FP64 orthonormalisation (see gladsgp_b200/svd.py)
    Q = Qd.float()
    B = xty(X_slab, Q.contiguous()).double()                  # (r, n_local)
    G = allsum(B @ B.T)
    lam, E = torch.linalg.eigh(G)
    lam = torch.flip(lam, dims=[0]).clamp_min(0.0)
    E = torch.flip(E, dims=[1])
    S = torch.sqrt(lam)
    U = (Qd @ E).float()
    Vh = ((E.T @ B) / S.clamp_min(1e-300)[:, None]).float()
    return U[:, :p], S[:p].float(), Vh[:p]


# ---------------------------------------------------------------------------------------------
# one chain, PCs spread over the ranks (SURVEY.md 8e ii; BASELINE.json north_star "by independent PC component")
# ---------------------------------------------------------------------------------------------
def pc_shard(pu, rank, world):
    """PCs [begin, begin + count) swept by `rank`: equal blocks of ceil(pu / world) (the last ranks may own fewer or
    none), so that a rank's rows form one contiguous, equally sized slot of the exchange buffer."""
    cp = -(-int(pu) // int(world))
    lo = min(pu, rank * cp)
    return lo, min(pu, lo + cp) - lo, cp


class PcRowGather:
    """The one collective of the PC-sharded sampler: after a step's sweep every rank holds the result rows of its own
    PCs in its slot of xchg (padded_pcs, n_chains, row_len); all_gather_into_tensor fills in the other ranks' slots.
    Per step and rank that is ceil(pu / world) * n_chains * (2 d + 6) doubles (cfg5, 8 GPUs: 3 x 1 x 40 x 8 B = 960 B)."""

    def __init__(self, pu, group=None):
        import torch.distributed as dist
        self.dist, self.group = dist, group
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        self.begin, self.count, self.cp = pc_shard(pu, self.rank, self.world)
        self.padded_pcs = self.cp * self.world
        self.calls = 0

    def __call__(self, xchg):
        mine = xchg[self.rank * self.cp:(self.rank + 1) * self.cp].clone()
        self.dist.all_gather_into_tensor(xchg.view(-1), mine.view(-1), group=self.group)
        self.calls += 1


def mcmc_by_pc(engine, n_steps, step, uniforms=None, replay=None, group=None, **kw):
    """Run engine's chain(s) with the PCs spread over the ranks of `group` (every rank passes the same tables, state
    and random stream / replay tables, and ends with the same state and draws).  One NCCL all_gather per step."""
    g = PcRowGather(engine.pu, group)
    out = engine.run_by_pc(n_steps, step, uniforms=uniforms, replay=replay, shards=[(g.begin, g.count)], gather=g, **kw)
    out['collective_calls'] = g.calls
    return out
