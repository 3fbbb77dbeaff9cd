"""CPU: the oracle against the committed golden vectors and the facts the reference records."""
import os
import numpy as np
import pytest

from helpers import so, svd_oracle, make_problem, tables_from_oracle, replay_from_trace

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')


def _signfix(A, B, axis):
    """Align the sign of singular vectors along `axis`."""
    s = np.sign(np.sum(A * B, axis=axis, keepdims=True))
    s[s == 0] = 1
    return A * s


@pytest.mark.parametrize('tag', ['a', 'b', 'c'])
def test_svd_oracle_matches_reference_svd_py(tag):
    """oracle/svd_oracle.py vs the outputs of the reference's own src/svd.py (pinned parity)."""
    g = np.load(os.path.join(GOLD, 'rsvd_reference.npz'))
    X = g['X']
    p, k, q = [int(v) for v in g['pkq_' + tag]]
    k = None if k < 0 else k
    np.random.seed(1000 + p)
    U, S, Vh = svd_oracle.randomized_svd(X, p, k=k, q=q)             # same global stream as the fixture
    assert U.shape == g['U_' + tag].shape and Vh.shape == g['Vh_' + tag].shape
    np.testing.assert_allclose(S, g['S_' + tag], rtol=1e-4)
    nz = S > 1e-3 * S[0]
    np.testing.assert_allclose(_signfix(Vh, g['Vh_' + tag], 1)[nz], g['Vh_' + tag][nz], atol=2e-3)
    np.testing.assert_allclose(_signfix(U, g['U_' + tag], 0)[:, nz], g['U_' + tag][:, nz], atol=2e-3)
    assert float(g['err_bound']) == 0.0        # the reference's error bound is always 0 (svd.py:67,73-76)


def test_oracle_regression_against_golden():
    g = np.load(os.path.join(GOLD, 'sepia_oracle.npz'))
    num = so.OracleNum(g['t'], (g['y'] - g['mu']) / g['sd'], g['K'])
    np.testing.assert_allclose(num.w, g['w'], rtol=1e-6, atol=1e-7)
    for b in range(4):
        C = so.block_cov(num, g['beta'][b], g['lamz'][b], g['lamws'][b], g['lamwos'][b], int(g['js'][b]))
        np.testing.assert_allclose(C, g['C'][b], rtol=1e-12, atol=1e-14)
        ll = so.do_loglik(g['C'][b], g['w'][:, int(g['js'][b])])
        np.testing.assert_allclose(ll, g['loglik'][b], rtol=1e-10)
    mod = so.OracleModel(num)
    mod.override_lamWOs(40.0)
    mod.do_mcmc(6, rng=np.random.RandomState(21))
    s = mod.get_samples()
    draws = np.concatenate([s['betaU'], s['lamUz'], s['lamWs'], s['lamWOs']], axis=1)
    np.testing.assert_allclose(draws, g['chain_draws'], rtol=1e-9)
    samples = {k[3:]: g[k] for k in g.files if k.startswith('ps_')}
    _, mu, Sig = so.w_pred(num, g['t_pred'], samples, draw=False)
    np.testing.assert_allclose(mu, g['pred_mu'], rtol=1e-8, atol=1e-10)
    np.testing.assert_allclose(Sig, g['pred_Sigma'], rtol=1e-8, atol=1e-12)


def test_defaults_recorded_in_the_reference_notebooks():
    """examples/04_*.ipynb:297-316 (default step sizes), :400 (flat sample shape (n, d*pu))."""
    pr = make_problem(m=32, q=8, pu=6, n_x=10, n_t=8)
    mod = so.OracleModel(pr['num'])
    assert mod.betaU.step.shape == (9, 6) and np.all(mod.betaU.step == 0.1)
    assert np.all(mod.lamUz.step == 5.0) and mod.lamUz.step.shape == (1, 6)
    assert np.all(mod.lamWs.step == 100.0) and np.all(mod.lamWOs.step == 100.0)
    mod.do_mcmc(3, rng=np.random.RandomState(0))
    s = mod.get_samples()
    assert s['betaU'].shape == (3, 54) and s['lamUz'].shape == (3, 6) and s['lamWOs'].shape == (3, 1)
    # Fortran flattening: first d entries are the d inputs of PC 0
    np.testing.assert_array_equal(s['betaU'][-1, :9], mod.betaU.val[:, 0])


def test_covariance_properties_and_loglik_failure():
    pr = make_problem(m=40, q=3, pu=2, n_x=8, n_t=5)
    num = pr['num']
    beta = np.array([0.3, 1.0, 2.0, 0.1])
    C = so.cov_self(num, beta, 2.0)
    assert np.allclose(C, C.T) and np.allclose(np.diag(C), 0.5)
    X = so.cov_cross(num.zt, num.zt, beta, 2.0)
    np.testing.assert_allclose(X, C, rtol=1e-12)
    assert so.do_loglik(-np.eye(5), np.ones(5)) == -np.inf
    # log-lik equals the Gaussian log-density up to the dropped 2 pi constant
    Cn = so.block_cov(num, beta, 2.0, 500.0, 80.0, 0)
    w = num.wv[:40, 0]
    sign, ld = np.linalg.slogdet(Cn)
    ref = -0.5 * ld - 0.5 * w @ np.linalg.solve(Cn, w)
    np.testing.assert_allclose(so.do_loglik(Cn, w), ref, rtol=1e-10)


def test_rng_consumption_rule():
    """One uniform per visited site plus one iff the candidate is valid."""
    pr = make_problem(m=24, q=2, pu=2, n_x=6, n_t=5)
    mod = so.OracleModel(pr['num'])
    mod.trace = []
    rs = np.random.RandomState(3)
    mod.do_mcmc(4, rng=rs)
    used = sum(1 + int(t['valid']) for t in mod.trace)
    rs2 = np.random.RandomState(3)
    rs2.random_sample(used)
    assert rs.random_sample() == rs2.random_sample()
    P = tables_from_oracle(mod)['theta'].size
    replay, acc = replay_from_trace(mod.trace, 4, P)
    assert replay['cand'].shape == (4, 1, P) and acc.sum() > 0


def test_accept_margins_far_from_ties_at_config_sizes():
    """The device forms the accept test from per-PC differences, (ll_new - ll_old) + (prior_new - prior_old), where
    SepiaModel.mcmc_step subtracts two full sums.  The two differ by rounding (~1e-12 absolute); a decision
    could only flip inside that band.  Over the 13 576 evaluated decisions of the golden chains (cfg1, cfg2, cfg3 sizes)
    the smallest |margin| = |(clp - lp + log aCorr) - log u| is five orders of magnitude above 1e-9."""
    import os
    gold = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')
    n = 0
    for cfg in ('cfg1', 'cfg2', 'cfg3'):
        g = np.load(os.path.join(gold, 'chain_%s.npz' % cfg))
        mg = g['margin']
        ok = np.isfinite(mg)
        assert np.array_equal(ok[:, None, :], g['rp_valid'].astype(bool))
        assert np.array_equal((mg > 0) & ok, g['chain_acc'][:, 0, :].astype(bool))
        assert np.min(np.abs(mg[ok])) > 1e-6
        n += int(ok.sum())
    assert n >= 10000
