"""CPU, world_size 2 over gloo: sharding by independent unit and the one gather of the path."""
import os
import socket
import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from helpers import ROOT  # noqa: F401
from gladsgp_b200.dist import shard_bounds, all_gather_concat, pc_shard, PcRowGather


def test_shard_bounds_cover_everything():
    for n in (0, 1, 7, 74, 100000):
        for world in (1, 2, 3, 8):
            b = [shard_bounds(n, r, world) for r in range(world)]
            assert b[0][0] == 0 and b[-1][1] == n
            assert all(b[i][1] == b[i + 1][0] for i in range(world - 1))
            sizes = [hi - lo for lo, hi in b]
            assert max(sizes) - min(sizes) <= 1


def _worker(rank, world, port, q):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    # chains sharded across ranks: every rank owns a contiguous slice and produces its per-chain values
    n_chains, n = 7, 11
    lo, hi = shard_bounds(n_chains, rank, world)
    lp = torch.arange(lo, hi, dtype=torch.float64) * 10.0
    full = all_gather_concat(lp)
    # prediction sharded by design block: (B, n_local) moments gathered along the design axis
    dlo, dhi = shard_bounds(n, rank, world)
    mean = torch.arange(3, dtype=torch.float64)[:, None] * 100 + torch.arange(dlo, dhi, dtype=torch.float64)[None, :]
    allm = all_gather_concat(mean, dim=1)
    t = torch.tensor([float(rank + 1)])
    dist.all_reduce(t, op=dist.ReduceOp.MAX)           # max-over-ranks timing reduction used by bench.py
    q.put((rank, full.numpy(), allm.numpy(), float(t)))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_gather():
    s = socket.socket(); s.bind(('127.0.0.1', 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    exp_mean = np.arange(3)[:, None] * 100.0 + np.arange(11)[None, :]
    for rank, full, allm, tmax in res:
        np.testing.assert_array_equal(full, np.arange(7) * 10.0)
        np.testing.assert_array_equal(allm, exp_mean)
        assert tmax == 2.0


def _svd_worker(rank, world, port, q):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from gladsgp_b200.dist import randomized_svd_sharded
    rng = np.random.RandomState(11)
    m, n, p = 40, 501, 6
    X = (rng.standard_normal((m, 8)) @ rng.standard_normal((8, n)) + 0.01 * rng.standard_normal((m, n))).astype(np.float32)
    omega = np.random.RandomState(5).normal(size=(n, p)).astype(np.float32)      # same stream on every rank
    lo, hi = shard_bounds(n, rank, world)
    products = (lambda A, omT: A @ omT.T, lambda A, Y: Y.T @ A)                   # stand-ins for the CUDA passes
    U, S, Vh = randomized_svd_sharded(torch.as_tensor(X[:, lo:hi].copy()), p, k=0, q=1, omega_slab=omega[lo:hi],
                                      products=products)
    Vh_full = all_gather_concat(Vh, dim=1)
    q.put((rank, U.numpy(), S.numpy(), Vh_full.numpy()))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharded_rsvd_matches_single_process():
    """rSVD sharded by output-column slab == the oracle on the whole matrix with the same test matrix."""
    from helpers import svd_oracle
    s = socket.socket(); s.bind(('127.0.0.1', 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    procs = [ctx.Process(target=_svd_worker, args=(r, 2, port, q)) for r in range(2)]
    for pr in procs:
        pr.start()
    res = [q.get(timeout=180) for _ in range(2)]
    for pr in procs:
        pr.join(timeout=60)
        assert pr.exitcode == 0
    rng = np.random.RandomState(11)
    m, n, p = 40, 501, 6
    X = (rng.standard_normal((m, 8)) @ rng.standard_normal((8, n)) + 0.01 * rng.standard_normal((m, n))).astype(np.float32)
    omega = np.random.RandomState(5).normal(size=(n, p)).astype(np.float32)
    U0, S0, Vh0 = svd_oracle.randomized_svd(X, p, k=0, q=1, omega=omega)
    for rank, U, S, Vh in res:
        np.testing.assert_allclose(S, S0, rtol=2e-4)
        np.testing.assert_allclose((U * S) @ Vh, (U0 * S0) @ Vh0, atol=2e-3 * float(S0[0]) / np.sqrt(m))
    np.testing.assert_array_equal(res[0][1], res[1][1])         # U is replicated
    np.testing.assert_array_equal(res[0][3], res[1][3])         # gathered Vh identical on both ranks


def test_pc_shard_partition():
    """PCs of one chain over the ranks: contiguous blocks of ceil(pu / world), every PC owned exactly once."""
    for pu in (1, 5, 10, 20):
        for world in (1, 2, 3, 8):
            owned = []
            for r in range(world):
                lo, cnt, cp = pc_shard(pu, r, world)
                assert cp == -(-pu // world) and 0 <= cnt <= cp and lo == min(pu, r * cp)
                owned += list(range(lo, lo + cnt))
            assert owned == list(range(pu))


def _pc_worker(rank, world, port, q):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    pu, n_chains, d = 5, 2, 3
    g = PcRowGather(pu)
    row = 2 * d + 6
    xchg = torch.zeros((g.padded_pcs, n_chains, row), dtype=torch.float64)
    for step in range(3):                                  # one exchange per mcmc_step
        xchg.zero_()
        for j in range(g.begin, g.begin + g.count):        # what the rank's sweep leaves for its PCs
            xchg[j] = 1000.0 * step + 10.0 * j + torch.arange(row, dtype=torch.float64)[None, :] / 100 + \
                torch.arange(n_chains, dtype=torch.float64)[:, None]
        g(xchg)
        q.put((rank, step, xchg[:pu].numpy().copy()))
    assert g.calls == 3
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_pc_row_exchange():
    """world_size 2 over gloo: every rank ends each step with the rows of all PCs (the sampler's one collective)."""
    s = socket.socket(); s.bind(('127.0.0.1', 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    procs = [ctx.Process(target=_pc_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(6)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    pu, n_chains, row = 5, 2, 12
    for rank, step, got in res:
        exp = 1000.0 * step + 10.0 * np.arange(pu)[:, None, None] + np.arange(row)[None, None, :] / 100 + \
            np.arange(n_chains)[None, :, None]
        np.testing.assert_array_equal(got, exp)
