"""tcgen05 (3xTF32) rSVD passes vs FP64 NumPy and vs the FP32-FMA kernels (src/svd.py:52-60 products)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

CASES = [(128, 64, 8), (256, 4096, 25), (512, 20001, 25), (100, 5003, 25), (700, 9000, 40), (1500, 3001, 25), (33, 257, 3)]


@pytest.mark.parametrize('m,n,r', CASES)
def test_sketch_and_xty_on_tensor_cores(cuda, m, n, r):
    import torch
    from gladsgp_b200 import ops
    rng = np.random.default_rng(m * 7 + r)
    X = (rng.standard_normal((m, n)) * rng.uniform(0.1, 3, size=(1, n))).astype(np.float32)
    Om = rng.standard_normal((r, n)).astype(np.float32)
    Yh = rng.standard_normal((m, r)).astype(np.float32)
    Xd, Od, Yd = (torch.as_tensor(a, device='cuda') for a in (X, Om, Yh))
    X64 = X.astype(np.float64)
    ref = X64 @ Om.astype(np.float64).T
    got = ops.rsvd_sketch_tc(Xd, Od).cpu().numpy()
    # FP32-level: the split keeps ~21 bits per product and chunk sums are added in round-to-nearest FP32
    assert np.abs(got - ref).max() <= 1e-6 * np.abs(ref).max()
    refb = Yh.astype(np.float64).T @ X64
    gotb = ops.rsvd_xty_tc(Xd, Yd).cpu().numpy()
    assert gotb.shape == (r, n)
    assert np.abs(gotb - refb).max() <= 1e-6 * np.abs(refb).max()
    if m <= 1024:
        simt = ops.rsvd_sketch(Xd, Od).cpu().numpy()
        assert np.abs(got - simt).max() <= 2e-6 * np.abs(ref).max()
    # deterministic: same bits on a second call
    assert np.array_equal(got, ops.rsvd_sketch_tc(Xd, Od).cpu().numpy())
    assert np.array_equal(gotb, ops.rsvd_xty_tc(Xd, Yd).cpu().numpy())


def test_exact_on_tf32_representable_inputs(cuda):
    """Inputs that are exactly representable in TF32 with small integer values: every product and sum is exact."""
    import torch
    from gladsgp_b200 import ops
    rng = np.random.default_rng(0)
    m, n, r = 256, 2048, 25
    X = rng.integers(-8, 9, size=(m, n)).astype(np.float32)
    Om = rng.integers(-4, 5, size=(r, n)).astype(np.float32)
    Yh = rng.integers(-4, 5, size=(m, r)).astype(np.float32)
    Xd, Od, Yd = (torch.as_tensor(a, device='cuda') for a in (X, Om, Yh))
    assert np.array_equal(ops.rsvd_sketch_tc(Xd, Od).cpu().numpy(), (X.astype(np.int64) @ Om.astype(np.int64).T).astype(np.float32))
    assert np.array_equal(ops.rsvd_xty_tc(Xd, Yd).cpu().numpy(), (Yh.astype(np.int64).T @ X.astype(np.int64)).astype(np.float32))


def test_randomized_svd_large_m(cuda):
    """m > 1024 (cfg5-sized ensembles) goes through the tensor-core passes; compare with a dense SVD."""
    from gladsgp_b200 import svd
    rng = np.random.default_rng(3)
    m, n, p = 1300, 6000, 12
    U0, _ = np.linalg.qr(rng.standard_normal((m, p)))
    V0, _ = np.linalg.qr(rng.standard_normal((n, p)))
    s0 = np.linspace(50, 5, p)
    X = ((U0 * s0) @ V0.T + 1e-3 * rng.standard_normal((m, n))).astype(np.float32)
    np.random.seed(1)
    U, S, Vh = svd.randomized_svd(X, p, k=8, q=1)
    s_all = np.linalg.svd(X.astype(np.float64), compute_uv=False)
    np.testing.assert_allclose(S, s_all[:p], rtol=2e-4)
    rec = (U * S) @ Vh
    best = np.sqrt(np.sum(s_all[p:] ** 2))                      # error of the optimal rank-p approximation
    assert np.linalg.norm(rec - X) <= 1.02 * best


def test_sharded_rsvd_single_rank_equals_randomized_svd(cuda):
    """dist.randomized_svd_sharded on one rank (no process group) == svd.randomized_svd with the same test matrix."""
    import torch
    from gladsgp_b200 import svd, dist as gdist
    rng = np.random.default_rng(8)
    m, n, p = 96, 4001, 10
    X = (rng.standard_normal((m, 12)) @ rng.standard_normal((12, n)) + 0.01 * rng.standard_normal((m, n))).astype(np.float32)
    omega = rng.standard_normal((n, p)).astype(np.float32)
    U0, S0, Vh0 = svd.randomized_svd(X, p, k=0, q=1, omega=omega)
    U, S, Vh = gdist.randomized_svd_sharded(torch.as_tensor(X, device='cuda'), p, k=0, q=1, omega_slab=omega)
    np.testing.assert_array_equal(S.cpu().numpy(), S0)
    np.testing.assert_array_equal(U.cpu().numpy(), U0)
    np.testing.assert_array_equal(Vh.cpu().numpy(), Vh0)


@pytest.mark.parametrize('m,n,r', [(256, 4096, 25), (512, 20004, 25), (700, 9000, 40), (100, 5004, 25), (33, 260, 3)])
def test_tma_fed_sketch_equals_register_staged_sketch(cuda, monkeypatch, m, n, r):
    """Row pitch a multiple of 16 bytes: the sketch pass takes its operand tiles through TMA (cp.async.bulk.tensor, 128-byte
    swizzle).  Split, MMA order and epilogue are those of the register-staged kernel, so the two give identical bits --
    including ragged last chunks and row blocks (zero fill by the hardware instead of by predicated loads)."""
    import torch
    from gladsgp_b200 import ops
    rng = np.random.default_rng(m + n)
    X = (rng.standard_normal((m, n)) * rng.uniform(0.1, 3, size=(1, n))).astype(np.float32)
    Om = rng.standard_normal((r, n)).astype(np.float32)
    Xd, Od = torch.as_tensor(X, device='cuda'), torch.as_tensor(Om, device='cuda')
    monkeypatch.setenv('GGP_TMA', '0')
    ref = ops.rsvd_sketch_tc(Xd, Od).cpu().numpy()
    monkeypatch.setenv('GGP_TMA', '1')
    got = ops.rsvd_sketch_tc(Xd, Od).cpu().numpy()
    assert np.array_equal(got, ref)
    truth = X.astype(np.float64) @ Om.astype(np.float64).T
    assert np.abs(got - truth).max() <= 1e-6 * np.abs(truth).max()


@pytest.mark.parametrize('mode', ['2', '3'])
@pytest.mark.parametrize('m,n,r', [(512, 20004, 25), (700, 9000, 40), (100, 5004, 25)])
def test_tma_sketch_fused_products_and_raw_hi_tile(cuda, monkeypatch, mode, m, n, r):
    """GGP_TMA=2: x_hi o_hi and x_hi o_lo as one N = 64 MMA, the three partial sums added in the epilogue; GGP_TMA=3 (default):
    additionally the hi tile of X is the raw float32 tile (the tensor core truncates to TF32) and only x - trunc(x) is written.
    Both stay at FP32-level accuracy against the float64 product and are deterministic."""
    import torch
    from gladsgp_b200 import ops
    rng = np.random.default_rng(m + n + 1)
    X = (rng.standard_normal((m, n)) * rng.uniform(0.1, 3, size=(1, n))).astype(np.float32)
    Om = rng.standard_normal((r, n)).astype(np.float32)
    Xd, Od = torch.as_tensor(X, device='cuda'), torch.as_tensor(Om, device='cuda')
    monkeypatch.setenv('GGP_TMA', mode)
    got = ops.rsvd_sketch_tc(Xd, Od).cpu().numpy()
    truth = X.astype(np.float64) @ Om.astype(np.float64).T
    assert np.abs(got - truth).max() <= 1e-6 * np.abs(truth).max()
    assert np.array_equal(got, ops.rsvd_sketch_tc(Xd, Od).cpu().numpy())


@pytest.mark.parametrize('m,n,r', [(512, 20004, 25), (700, 9000, 40), (100, 5004, 25), (33, 260, 3), (256, 4096, 25),
                                   (512, 20032, 25), (300, 8224, 40), (70, 320, 5)])      # last three: n % 32 == 0, one 3-D box per chunk
def test_tma_fed_xty_matches_float64_and_tmem_kernel(cuda, monkeypatch, m, n, r):
    """Y^T X with the X tiles brought in by TMA in the MN-major 128-byte / 32-byte-atom swizzle layout: FP32-level accuracy
    against the float64 product, agreement with the register -> TMEM kernel (GGP_TMA=0) at the same level, deterministic."""
    import torch
    from gladsgp_b200 import ops
    rng = np.random.default_rng(m + n + 2)
    X = (rng.standard_normal((m, n)) * rng.uniform(0.1, 3, size=(1, n))).astype(np.float32)
    Yh = rng.standard_normal((m, r)).astype(np.float32)
    Xd, Yd = torch.as_tensor(X, device='cuda'), torch.as_tensor(Yh, device='cuda')
    truth = Yh.astype(np.float64).T @ X.astype(np.float64)
    monkeypatch.setenv('GGP_TMA_XTY', '0')
    ref = ops.rsvd_xty_tc(Xd, Yd).cpu().numpy()
    monkeypatch.setenv('GGP_TMA_XTY', '1')             # (off by default: slower than the register -> TMEM kernel, see csrc)
    got = ops.rsvd_xty_tc(Xd, Yd).cpu().numpy()
    assert got.shape == (r, n)
    assert np.abs(got - truth).max() <= 1e-6 * np.abs(truth).max()
    assert np.abs(got - ref).max() <= 2e-6 * np.abs(truth).max()
    assert np.array_equal(got, ops.rsvd_xty_tc(Xd, Yd).cpu().numpy())


@pytest.mark.parametrize('m,n', [(512, 1460000), (1100, 1200004)])
def test_passes_at_full_ensemble_size_against_float64(cuda, m, n):
    """BASELINE.json cfg3 at full size (512 x 1.46 M float32, 3 GB) and an ensemble beyond 4 GB (byte offsets above 2^32): both
    passes against float64 products formed on the device in column chunks; inputs generated on the device."""
    import torch
    from gladsgp_b200 import ops
    r = 25
    g = torch.Generator(device='cuda'); g.manual_seed(m)
    X = torch.randn((m, n), dtype=torch.float32, device='cuda', generator=g)
    X *= torch.rand((1, n), dtype=torch.float32, device='cuda', generator=g) * 2.9 + 0.1
    Om = torch.randn((r, n), dtype=torch.float32, device='cuda', generator=g)
    Yh = torch.randn((m, r), dtype=torch.float32, device='cuda', generator=g)
    got = ops.rsvd_sketch_tc(X, Om).double()
    gotb = ops.rsvd_xty_tc(X, Yh)
    ref = torch.zeros((m, r), dtype=torch.float64, device='cuda')
    errb, scaleb = 0.0, 0.0
    step = 100000
    for c0 in range(0, n, step):
        Xc = X[:, c0:c0 + step].double()
        ref += Xc @ Om[:, c0:c0 + step].double().T
        rb = Yh.double().T @ Xc
        errb = max(errb, float((gotb[:, c0:c0 + step].double() - rb).abs().max()))
        scaleb = max(scaleb, float(rb.abs().max()))
    assert float((got - ref).abs().max()) <= 2e-6 * float(ref.abs().max())
    assert errb <= 1e-6 * scaleb
