"""GPU ensemble ingest passes vs the NumPy restatement of src/model.py:60-73, 218-224."""
import numpy as np
import pytest

from helpers import synthetic, make_problem
from oracle import init_oracle as io

pytestmark = pytest.mark.gpu


def _ensemble(m, n, seed=0, const_cols=3):
    rng = np.random.default_rng(seed)
    y = (1.0 + 0.3 * rng.standard_normal((m, n)) * rng.uniform(0.1, 2.0, size=n) + rng.uniform(0, 2, size=n)).astype(np.float32)
    y[:, :const_cols] = 0.75                      # nodes without variation (the sd clamp of src/model.py:64)
    return y


@pytest.mark.parametrize('m,n', [(64, 1000), (100, 4099), (512, 20000), (37, 129)])
def test_colstats_and_standardize_both_layouts(cuda, m, n):
    import torch
    from gladsgp_b200 import ops
    y = _ensemble(m, n, seed=m + n)
    mu_ref, sd_ref = io.column_stats(y.astype(np.float64))
    for transposed in (False, True):
        yd = torch.as_tensor(np.ascontiguousarray(y.T) if transposed else y, device='cuda')
        mu, sd = ops.colstats(yd, transposed=transposed, ddof=1, sd_floor=1e-6)
        mu, sd = mu.cpu().numpy(), sd.cpu().numpy()
        np.testing.assert_allclose(mu, mu_ref, rtol=2e-7)
        np.testing.assert_allclose(sd, sd_ref, rtol=3e-7, atol=1e-9)
        assert np.all(sd[:3] == np.float32(1e-6))
        # standardise with NumPy's own float32 statistics: bit-identical to the reference expression
        mu32, sd32 = io.column_stats(y)
        ys = ops.standardize(yd, torch.as_tensor(mu32, device='cuda'), torch.as_tensor(sd32, device='cuda'),
                             transposed=transposed).cpu().numpy()
        assert ys.shape == (m, n)
        assert np.array_equal(ys, io.standardize(y, mu32, sd32))


def test_row_sliced_views_are_used_in_place(cuda):
    import torch
    from gladsgp_b200 import ops
    M, m, n = 96, 40, 3001
    y = _ensemble(M, n, seed=9)
    mu_ref, sd_ref = io.column_stats(y[:m].astype(np.float64))
    yd = torch.as_tensor(y, device='cuda')
    ytd = torch.as_tensor(np.ascontiguousarray(y.T), device='cuda')
    for view, tr in ((yd[:m], False), (ytd[:, :m], True)):
        mu, sd = ops.colstats(view, transposed=tr, sd_floor=1e-6)
        np.testing.assert_allclose(mu.cpu().numpy(), mu_ref, rtol=2e-7)
        np.testing.assert_allclose(sd.cpu().numpy(), sd_ref, rtol=3e-7, atol=1e-9)
        out = ops.standardize(view, mu, sd, transposed=tr).cpu().numpy()
        assert np.array_equal(out, (y[:m] - mu.cpu().numpy()) / sd.cpu().numpy())


def test_scalar_mean_and_sd(cuda):
    import torch
    from gladsgp_b200 import ops
    y = _ensemble(50, 777, seed=2)
    out = ops.standardize(torch.as_tensor(y, device='cuda'), torch.tensor([0.5], device='cuda'),
                          torch.tensor([2.5], device='cuda')).cpu().numpy()
    assert np.array_equal(out, (y - np.float32(0.5)) / np.float32(2.5))


@pytest.mark.parametrize('m,n,pu', [(64, 5000, 3), (512, 20011, 10), (300, 9000, 20), (700, 4000, 25), (10, 40000, 10)])
def test_project_fp64(cuda, m, n, pu):
    import torch
    from gladsgp_b200 import ops
    rng = np.random.default_rng(pu)
    X = rng.standard_normal((m, n)).astype(np.float32)
    K = rng.standard_normal((pu, n)).astype(np.float32)
    P = ops.project(torch.as_tensor(X, device='cuda'), torch.as_tensor(K, device='cuda')).cpu().numpy()
    X64, K64 = X.astype(np.float64), K.astype(np.float64)
    scale = np.sqrt(n)
    np.testing.assert_allclose(P[:, :pu], X64 @ K64.T, rtol=0, atol=1e-11 * scale)
    np.testing.assert_allclose(P[:, pu], X64.sum(1), rtol=0, atol=1e-11 * scale)
    np.testing.assert_allclose(P[:, pu + 1], (X64 * X64).sum(1), rtol=1e-13)
    # fixed reduction order: repeated calls agree bit for bit
    P2 = ops.project(torch.as_tensor(X, device='cuda'), torch.as_tensor(K, device='cuda')).cpu().numpy()
    assert np.array_equal(P, P2)


def test_pc_weights_and_precision_match_reference_lines(cuda):
    from gladsgp_b200 import ingest
    pr = make_problem(m=64, q=3, pu=4, n_x=60, n_t=25)
    ys = pr['y_std'].astype(np.float32)
    proj = ingest.project_basis(ingest.upload(ys), ingest.upload(pr['K']))
    w_ref, prec_ref, ss_ref = io.pc_weights_and_precision(ys, pr['K'])
    np.testing.assert_allclose(proj['w'], w_ref, rtol=1e-9, atol=1e-11)
    np.testing.assert_allclose(proj['resid_ss'], ss_ref, rtol=1e-9)
    np.testing.assert_allclose(ingest.pc_precision_from(proj), prec_ref, rtol=1e-9)


def test_device_path_of_the_api_equals_host_path(cuda, monkeypatch):
    """SepiaData / SepiaModel / model.pc_precision give the same numbers whichever side does the passes; the
    transposed (n_y, m) file view is handled without a host transpose."""
    from gladsgp_b200 import ingest, model as gmodel
    from sepia.SepiaData import SepiaData
    from sepia.SepiaModel import SepiaModel
    pr = make_problem(m=64, q=3, pu=3, n_x=40, n_t=12)
    y = pr['y'].astype(np.float32)
    file_layout = np.ascontiguousarray(y.T)                      # (n_y, m), as aggregate_outputs.py writes it
    mu, sd = io.column_stats(y)

    def build(y_in, min_elems):
        monkeypatch.setattr(ingest, 'DEVICE_MIN_ELEMS', min_elems)
        d = SepiaData(t_sim=pr['t'], y_sim=y_in, y_ind_sim=np.linspace(0, 1, y.shape[1]))
        d.transform_xt(t_notrans=np.arange(3))
        d.standardize_y(y_mean=mu, y_sd=sd)
        d.create_K_basis(K=pr['K'])
        mdl = SepiaModel(d)
        return d, mdl, gmodel.pc_precision(d.sim_data)

    d_h, m_h, p_h = build(y, 1 << 60)
    assert d_h.sim_data._y_std_dev is None
    for y_in in (y, file_layout.T):
        d_d, m_d, p_d = build(y_in, 0)
        assert d_d.sim_data._y_std_dev is not None and d_d.sim_data._y_std is None
        assert np.array_equal(d_d.sim_data.y_std, d_h.sim_data.y_std)          # lazily downloaded, same bits
        np.testing.assert_allclose(m_d.num.w, m_h.num.w, rtol=1e-5, atol=1e-6)  # host path multiplies in float32
        np.testing.assert_allclose(m_d.num.LamSim, m_h.num.LamSim, rtol=1e-12)
        np.testing.assert_allclose(p_d, p_h, rtol=1e-6)
        w_ref, prec_ref, _ = io.pc_weights_and_precision(d_h.sim_data.y_std, pr['K'])
        np.testing.assert_allclose(m_d.num.w.reshape(-1, 64).T, w_ref, rtol=1e-9, atol=1e-11)
        np.testing.assert_allclose(p_d, prec_ref, rtol=1e-9)
        for a, b in zip(m_d.params.lamWOs.prior.params, m_h.params.lamWOs.prior.params):
            np.testing.assert_allclose(a, b, rtol=1e-6)
    # default standardisation (column mean, scalar sd) on the device vs NumPy
    monkeypatch.setattr(ingest, 'DEVICE_MIN_ELEMS', 0)
    d2 = SepiaData(t_sim=pr['t'], y_sim=y, y_ind_sim=np.linspace(0, 1, y.shape[1]))
    d2.standardize_y()
    yc = y - np.mean(y, axis=0)
    np.testing.assert_allclose(d2.sim_data.orig_y_sd, np.std(yc.astype(np.float64), ddof=1), rtol=1e-6)
    # (NumPy's float32 column mean of a constant column is off by an ulp or two; the device mean is exact)
    np.testing.assert_allclose(d2.sim_data.y_std, yc / np.std(yc, ddof=1), rtol=1e-5, atol=1e-5)
    d3 = SepiaData(t_sim=pr['t'], y_sim=y, y_ind_sim=np.linspace(0, 1, y.shape[1]))
    d3.standardize_y(scale='columnwise')
    np.testing.assert_allclose(d3.sim_data.orig_y_sd, np.std(y.astype(np.float64), ddof=1, axis=0), rtol=1e-6)
    # user-supplied vector mean that is NOT the column mean, scalar sd left to the library: np.std(y - y_mean, ddof=1)
    shifted = (np.mean(y, axis=0) + np.linspace(-0.3, 0.4, y.shape[1])).astype(np.float32)
    d4 = SepiaData(t_sim=pr['t'], y_sim=y, y_ind_sim=np.linspace(0, 1, y.shape[1]))
    d4.standardize_y(y_mean=shifted)
    np.testing.assert_allclose(d4.sim_data.orig_y_sd, np.std(y.astype(np.float64) - shifted.astype(np.float64), ddof=1), rtol=1e-6)


def test_init_model_on_device(cuda, tmp_path, monkeypatch):
    from gladsgp_b200 import ingest, model as gmodel
    monkeypatch.setattr(ingest, 'DEVICE_MIN_ELEMS', 0)
    t = synthetic.design(64, 3, seed=1)
    y = synthetic.ensemble(t, n_x=50, n_t=20, seed=1).astype(np.float32)
    np.random.seed(0)
    data, mdl = gmodel.init_model(t, np.ascontiguousarray(y.T).T, 'dev', 4, data_dir=str(tmp_path))
    mu, sd = io.column_stats(y.astype(np.float64))
    np.testing.assert_allclose(data.sim_data.orig_y_mean, mu, rtol=2e-7)
    np.testing.assert_allclose(data.sim_data.orig_y_sd, sd, rtol=3e-7, atol=1e-9)
    assert data.sim_data.K.shape == (4, y.shape[1]) and mdl.num.w.shape == (4 * 64, 1)
    ys = data.sim_data.y_std
    w_ref, prec_ref, _ = io.pc_weights_and_precision(ys, data.sim_data.K)
    np.testing.assert_allclose(mdl.num.w.reshape(4, 64).T, w_ref, rtol=1e-8, atol=1e-10)
    np.testing.assert_allclose(gmodel.pc_precision(data.sim_data), prec_ref, rtol=1e-8)
    assert np.isfinite(mdl.logLik())
