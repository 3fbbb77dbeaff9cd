"""GPU edge cases: degenerate sizes, ragged shapes, error behaviour of the C ABI."""
import ctypes as C

import numpy as np
import pytest
import scipy.linalg

pytestmark = pytest.mark.gpu


def _ref_ll(X, w, beta, lamz, dadd):
    D = ((X[:, None, :] - X[None, :, :]) ** 2) @ beta
    Cm = np.exp(-D) / lamz
    np.fill_diagonal(Cm, 1.0 / lamz + dadd)
    L = scipy.linalg.cholesky(Cm, lower=True)
    u = scipy.linalg.solve_triangular(L, w, lower=True)
    return -np.sum(np.log(np.diag(L))) - 0.5 * u @ u


@pytest.mark.parametrize('m,d', [(1, 1), (2, 3), (7, 1), (31, 2), (32, 9), (65, 17)])
def test_tiny_and_ragged_matrices(cuda, m, d):
    """Matrices smaller than one 32-column panel, a single point, one input dimension, m = 32k + 1."""
    from gladsgp_b200 import ops
    rng = np.random.default_rng(m * 31 + d)
    X = rng.uniform(0, 1, size=(m, d))
    B = 3
    beta = rng.uniform(0.05, 3.0, size=(B, d)); lamz = rng.uniform(0.5, 2.0, size=B); dadd = rng.uniform(1e-3, 1e-2, size=B)
    W = rng.standard_normal((B, m))
    out = ops.loglik_batched(X, W, beta, lamz, dadd)
    ll = out['loglik'].cpu().numpy()
    assert np.all(out['info'].cpu().numpy() == 0)
    for b in range(B):
        ref = _ref_ll(X, W[b], beta[b], lamz[b], dadd[b])
        assert abs(ll[b] - ref) <= 1e-8 * max(abs(ref), 1e-12), (ll[b], ref)
    # covariance / cross-covariance at the same sizes
    Cg = ops.cov_build(X, beta, lamz, dadd).cpu().numpy()
    Xp = rng.uniform(0, 1, size=(1, d))                      # a single test design
    S21 = ops.cross_cov(X, Xp, beta, lamz).cpu().numpy()
    for b in range(B):
        D = ((X[:, None, :] - X[None, :, :]) ** 2) @ beta[b]
        Cm = np.exp(-D) / lamz[b]
        np.fill_diagonal(Cm, 1.0 / lamz[b] + dadd[b])
        np.testing.assert_allclose(Cg[b], Cm, rtol=1e-13)
        Dp = ((X[:, None, :] - Xp[None, :, :]) ** 2) @ beta[b]
        np.testing.assert_allclose(S21[b], np.exp(-Dp) / lamz[b], rtol=1e-13)


def test_single_design_prediction(cuda):
    """npred = 1 (the reference predicts one or four designs per call)."""
    from gladsgp_b200 import ops
    rng = np.random.default_rng(1)
    m, d, B = 50, 4, 3
    X = rng.uniform(0, 1, size=(m, d))
    beta = rng.uniform(0.1, 2.0, size=(B, d)); lamz = rng.uniform(0.5, 2.0, size=B)
    lamws = rng.uniform(300, 3000, size=B); dadd = 1.0 / lamws + 1e-3
    W = rng.standard_normal((B, m))
    P = ops.Predictor(X, W, beta, lamz, dadd, 1.0 / lamz + 1.0 / lamws)
    xp = rng.uniform(0, 1, size=(1, d))
    mean, var = P.predict(xp)
    for b in range(B):
        D = ((X[:, None, :] - X[None, :, :]) ** 2) @ beta[b]
        S22 = np.exp(-D) / lamz[b]
        np.fill_diagonal(S22, 1.0 / lamz[b] + dadd[b])
        s21 = np.exp(-(((X - xp) ** 2) @ beta[b])) / lamz[b]
        a = np.linalg.solve(S22, s21)
        np.testing.assert_allclose(mean[b, 0].item(), a @ W[b], rtol=1e-6, atol=1e-9)
        np.testing.assert_allclose(var[b, 0].item(), 1.0 / lamz[b] + 1.0 / lamws[b] - s21 @ a, rtol=1e-6)


def test_all_proposals_out_of_bounds_consume_one_uniform_per_site(cuda):
    """Bounds that no candidate can satisfy: nothing is evaluated, the state does not move and exactly one uniform per
    visited site is consumed (SEPIA draws the candidate before it tests the bounds)."""
    from helpers import make_problem, tables_from_oracle, so
    from gladsgp_b200 import ops
    pr = make_problem(m=40, q=2, pu=2)
    om = so.OracleModel(pr['num'])
    tb = tables_from_oracle(om)
    P = tb['theta'].size
    tb['lo'] = tb['theta'] + 1e6          # every candidate is below its lower bound
    tb['hi'] = tb['theta'] + 2e6
    eng = ops.McmcEngine(pr['num'].zt, pr['num'].wv.reshape(2, 40), pr['num'].LamSim, tb, n_chains=1)
    eng.set_state(tb['theta'])
    us = np.random.default_rng(0).random((1, 2 * P * 3))
    out = eng.run(3, tb['step'], uniforms=us, record=True, record_accept=True)
    assert int(out['consumed'].cpu()[0]) == 3 * P
    assert not out['accepted'].cpu().numpy().any()
    np.testing.assert_array_equal(out['draws'].cpu().numpy()[:, 0, :], np.tile(tb['theta'], (3, 1)))


@pytest.mark.parametrize('chains', [1, 3, 160])
def test_candidate_with_a_failing_cholesky_is_rejected_inside_the_sampler(cuda, chains):
    """SEPIA: a covariance that is not positive definite gives log-lik -inf, i.e. the proposal is rejected and the sweep goes
    on.  Replayed candidates that break the factorisation (negative lamUz; a negative nugget through lamWs) must leave
    exactly the chain of a run in which those proposals were marked invalid -- for one chain (cluster / speculative
    kernels), a few chains (cluster step kernel) and many (look-ahead step kernel)."""
    from helpers import make_problem, tables_from_oracle, replay_from_trace, so
    from gladsgp_b200 import ops
    pr = make_problem(m=100, q=3, pu=2)
    num = pr['num']
    om = so.OracleModel(num)
    tb = tables_from_oracle(om)
    P = tb['theta'].size
    d, pu = num.d, 2
    om.trace = []
    om.do_mcmc(4, rng=np.random.RandomState(3))
    replay, _ = replay_from_trace(om.trace, 4, P)
    bad = {k: v.copy() for k, v in replay.items()}
    off = {k: v.copy() for k, v in replay.items()}
    sites = [(1, d * pu + 1, -3.0),            # step 1: lamUz of PC 1 negative -> negative definite covariance
             (2, d * pu + pu + 0, -1.0e-3)]    # step 2: lamWs of PC 0 -> nugget 1/lamWs = -1000
    for t, s_, val in sites:
        bad['cand'][t, 0, s_] = val; bad['valid'][t, 0, s_] = 1; bad['logu'][t, 0, s_] = -1e300; bad['logacorr'][t, 0, s_] = 0.0
        off['valid'][t, 0, s_] = 0
    res = []
    for rp in (bad, off):
        rpc = {k: np.repeat(v, chains, axis=1) for k, v in rp.items()}
        eng = ops.McmcEngine(num.zt, num.w.T.copy(), num.LamSim, tb, n_chains=chains)
        eng.set_state(tb['theta'])
        out = eng.run(4, tb['step'], replay=rpc, record_accept=True)
        res.append((out['draws'].cpu().numpy(), out['accepted'].cpu().numpy(), out['lp'].cpu().numpy()))
    assert np.isfinite(res[0][2]).all()
    for a, b in zip(res[0], res[1]):
        assert np.array_equal(a, b)
    for t, s_, _ in sites:
        assert not res[0][1][t, :, s_].any()
    for c in range(1, chains):
        assert np.array_equal(res[0][0][:, c], res[0][0][:, 0])


def test_c_abi_reports_errors(cuda):
    """Bad arguments and short workspaces come back as error codes with a message; nothing is thrown across the ABI."""
    import torch
    from gladsgp_b200 import _lib
    lib = _lib.load()
    x = torch.zeros(16, dtype=torch.float64, device='cuda')
    rc = lib.ggp_loglik_batched_f64(x.data_ptr(), 0, 1, x.data_ptr(), 0, x.data_ptr(), x.data_ptr(), x.data_ptr(), 1,
                                    x.data_ptr(), None, x.data_ptr(), x.data_ptr(), None)
    assert rc == -1 and b'positive' in lib.ggp_last_error_string()
    rc = lib.ggp_loglik_batched_f64(None, 4, 1, x.data_ptr(), 4, x.data_ptr(), x.data_ptr(), x.data_ptr(), 1,
                                    x.data_ptr(), None, x.data_ptr(), x.data_ptr(), None)
    assert rc == -1 and b'null' in lib.ggp_last_error_string()
    a = _lib.McmcArgs()
    a.m, a.d, a.pu, a.n_chains, a.n_steps = 8, 1, 1, 1, 1
    rc = lib.ggp_mcmc_run_f64(C.byref(a), None)
    assert rc == -1
    f = torch.zeros(64, dtype=torch.float32, device='cuda')
    rc = lib.ggp_rsvd_sketch_tc_f32(f.data_ptr(), 8, 8, f.data_ptr(), 4, f.data_ptr(), f.data_ptr(), 16, None)
    assert rc == -4 and b'workspace' in lib.ggp_last_error_string()
    assert lib.ggp_project_workspace_bytes(8, 33) == -1
    with pytest.raises(_lib.GgpError):
        _lib.check(rc, 'ggp_rsvd_sketch_tc_f32')
