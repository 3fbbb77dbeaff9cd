"""GPU parity of the Sobol' / Saltelli sensitivity path: gladsgp_b200.sensitivity against outputs of the
reference's own src/utils.py (tests/golden/sobol_reference.npz) and against the oracle on seeded inputs."""
import os
import sys
import warnings

import numpy as np
import pytest

from helpers import so, synthetic, make_problem, ROOT

sys.path.insert(0, os.path.join(ROOT, 'tests', 'golden'))
from make_golden_sobol_function import analytic_function  # noqa: E402
from oracle import sobol_oracle as sob  # noqa: E402

pytestmark = pytest.mark.gpu
G = np.load(os.path.join(ROOT, 'tests', 'golden', 'sobol_reference.npz'))
IDX_RTOL = 1e-10      # indices and confidence limits: FP64 sums in a different order than NumPy's pairwise ones


def _ci(r):
    return np.array((r.confidence_interval.low, r.confidence_interval.high))


def test_pca_indices_and_bca_limits_match_reference_outputs(cuda):
    from gladsgp_b200 import sensitivity
    np.random.seed(int(G['np_seed']))           # the reference's bootstrap draws from the global state; so does the mirror
    first, total, gf, gt, res = sensitivity.PCA_saltelli_sensitivity_indices(
        analytic_function, int(G['n_dim']), int(G['m']), G['pcvar'], bootstrap=True, AB=G['pca_AB'])
    np.testing.assert_allclose(first, G['pca_first'], rtol=IDX_RTOL, atol=1e-14)
    np.testing.assert_allclose(total, G['pca_total'], rtol=IDX_RTOL, atol=1e-14)
    np.testing.assert_allclose(gf, G['pca_gen_first'], rtol=IDX_RTOL)
    np.testing.assert_allclose(gt, G['pca_gen_total'], rtol=IDX_RTOL)
    for k in ('first_order', 'total_index', 'general_first_order', 'general_total_index'):
        np.testing.assert_allclose(_ci(res[k]), G['pca_ci_' + k], rtol=1e-9, atol=1e-12, err_msg=k)
        np.testing.assert_allclose(res[k].standard_error, G['pca_se_' + k], rtol=1e-9, err_msg=k)
        assert res[k].bootstrap_distribution.shape[-1] == 9999


def test_scalar_indices_match_reference_outputs(cuda):
    from gladsgp_b200 import sensitivity
    np.random.seed(int(G['np_seed']))
    first, total, res = sensitivity.saltelli_sensitivity_indices(analytic_function, int(G['n_dim']), int(G['m']),
                                                                 bootstrap=True, AB=G['scalar_AB'])
    np.testing.assert_allclose(first, G['scalar_first'], rtol=IDX_RTOL, atol=1e-14)
    np.testing.assert_allclose(total, G['scalar_total'], rtol=IDX_RTOL, atol=1e-14)
    for k in ('first_order', 'total_index'):
        np.testing.assert_allclose(_ci(res[k]), G['scalar_ci_' + k], rtol=1e-9, atol=1e-12, err_msg=k)


@pytest.mark.parametrize('n_dim,m,p', [(8, 6, 10), (2, 4, 1), (5, 7, 3)])
def test_statistics_of_index_sets_match_oracle(cuda, n_dim, m, p):
    """The kernel's per-index-set statistics against the closures the reference hands to scipy.stats.bootstrap."""
    from gladsgp_b200 import ops
    rng = np.random.default_rng(100 + n_dim)
    N = 2 ** m
    f_A = rng.standard_normal((N, p)) + 0.3
    f_B = 0.5 * f_A + rng.standard_normal((N, p))
    f_AB = 0.7 * f_B[None] + 0.3 * rng.standard_normal((n_dim, N, p))
    st = sob.statistics(f_A, f_B, f_AB)
    dev = ops.sobol_upload(f_A, f_B, f_AB)
    idx = rng.integers(0, N, (37, N))
    first, total = ops.sobol_stats(dev, N, p, n_dim, idx=idx, clamp=True)
    for r in range(37):
        np.testing.assert_allclose(first[r], st['first_order'](idx[r]), rtol=IDX_RTOL, atol=1e-14)
        np.testing.assert_allclose(total[r], st['total_index'](idx[r]), rtol=IDX_RTOL, atol=1e-14)
    # ragged use: jackknife sets (N-1 entries) and the unclamped point estimates
    jk = np.array([np.delete(np.arange(N), i) for i in range(0, N, max(N // 8, 1))])
    fj, tj = ops.sobol_stats(dev, N, p, n_dim, idx=jk, clamp=True)
    for r in range(jk.shape[0]):
        np.testing.assert_allclose(fj[r], st['first_order'](jk[r]), rtol=IDX_RTOL, atol=1e-14)
    f0, t0 = ops.sobol_stats(dev, N, p, n_dim, idx=None, clamp=False)
    fo, to = sob.point_estimates(f_A, f_B, f_AB)
    np.testing.assert_allclose(f0[0], fo, rtol=IDX_RTOL, atol=1e-14)
    np.testing.assert_allclose(t0[0], to, rtol=IDX_RTOL, atol=1e-14)
    with pytest.raises(ValueError):
        ops.sobol_stats(dev, N, p, n_dim, idx=np.full((2, N), N))


def test_seeded_generator_matches_oracle_bootstrap(cuda):
    """Same seeded Generator on both sides, fewer resamples: the BCa limits agree with scipy.stats.bootstrap."""
    from gladsgp_b200 import sensitivity
    AB = sob.sobol_matrix(4, 6, seed=3)

    def func(x):
        return np.stack([np.sin(3 * x[:, 0]) + x[:, 1] * x[:, 2] + 0.5 * x[:, 3], x[:, 0] ** 2 + 0.3 * x[:, 1] + x[:, 2] + 0.2 * x[:, 3]], axis=1)
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        ref = sob.PCA_saltelli_sensitivity_indices(func, 4, 6, np.array([0.7, 0.3]), AB=AB, n_resamples=499, rng=np.random.default_rng(9))
    got = sensitivity.PCA_saltelli_sensitivity_indices(func, 4, 6, np.array([0.7, 0.3]), AB=AB, n_resamples=499, rng=np.random.default_rng(9))
    for a, b in zip(ref[:4], got[:4]):
        np.testing.assert_allclose(b, a, rtol=IDX_RTOL, atol=1e-14)
    for k in ref[4]:
        np.testing.assert_allclose(_ci(got[4][k]), _ci(ref[4][k]), rtol=1e-9, atol=1e-12, err_msg=k)


def test_emulator_mean_function_and_indices(cuda):
    """func of sensitivity_indices.py:73-91 on the cached-factor predictor: equals the posterior-sample mean of the predictive
    means of SepiaEmulatorPrediction, and drives the Saltelli scheme end to end."""
    from gladsgp_b200 import sensitivity
    from sepia.SepiaData import SepiaData
    from sepia.SepiaModel import SepiaModel
    from sepia.SepiaPredict import SepiaEmulatorPrediction
    pr = make_problem(m=64, q=3, pu=2)
    data = SepiaData(t_sim=pr['t'], y_sim=pr['y'], y_ind_sim=np.arange(pr['y'].shape[1], dtype=float))
    data.transform_xt(t_notrans=np.arange(3)); data.standardize_y(y_mean=pr['mu'], y_sd=pr['sd']); data.create_K_basis(K=pr['K'])
    model = SepiaModel(data)
    samples = synthetic.posterior_samples(5, model.num.p + model.num.q, 2, seed=2)
    func = sensitivity.emulator_mean_function(model, samples)
    x = synthetic.test_design(16, 3)
    pe = SepiaEmulatorPrediction(t_pred=x, samples=samples, model=model, storeRlz=False, storeMuSigma=True)
    mu = pe.get_mu_sigma()[0].reshape(5, 2, 16).mean(axis=0).T
    np.testing.assert_allclose(func(x), mu, rtol=1e-10, atol=1e-12)
    first, total, gf, gt, res = sensitivity.PCA_saltelli_sensitivity_indices(func, 3, 6, np.array([0.8, 0.2]), n_resamples=199,
                                                                             AB=sob.sobol_matrix(3, 6, seed=1), rng=np.random.default_rng(0))
    assert first.shape == (2, 3) and gf.shape == (3,) and np.all(np.isfinite(total)) and np.all(total > -1e-12)
    assert np.all(res['general_total_index'].confidence_interval.low <= res['general_total_index'].confidence_interval.high)
