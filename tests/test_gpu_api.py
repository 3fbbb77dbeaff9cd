"""GPU end-to-end through the reference-facing Python API (sepia.* mirror) vs the CPU oracle."""
import os
import numpy as np
import pytest

from helpers import so, svd_oracle, synthetic, make_problem, make_scalar_problem

pytestmark = pytest.mark.gpu


def _build(pr, q):
    from sepia.SepiaData import SepiaData
    from sepia.SepiaModel import SepiaModel
    d = SepiaData(t_sim=pr['t'], y_sim=pr['y'], y_ind_sim=np.linspace(0, 1, pr['y'].shape[1]))
    d.transform_xt(t_notrans=np.arange(q))
    d.standardize_y(y_mean=pr['mu'], y_sd=pr['sd'])
    d.create_K_basis(K=pr['K'])
    return d, SepiaModel(d)


def test_do_mcmc_follows_numpy_stream_like_oracle(cuda):
    """Same seed -> same accept sequence and chain as the oracle driven by the same global stream,
    and the global stream is left where SEPIA would leave it."""
    pr = make_problem(m=64, q=3, pu=2)
    data, model = _build(pr, 3)
    om = so.OracleModel(so.OracleNum(pr['t'], data.sim_data.y_std, pr['K']))
    # identical lamWOs prior and PC weights (the oracle's pinv / residual sum differ at float32 round-off)
    om.lamWOs.params = [p.copy() for p in model.params.lamWOs.prior.params]
    om.num.w = np.ascontiguousarray(model._w_pcs.T)
    om.num.wv = om.num.w.reshape((-1, 1), order='F')
    om.num.LamSim = np.asarray(model.num.LamSim, dtype=np.float64).copy()      # (float64 vs float32 product of K K^T)
    np.random.seed(7)
    om.do_mcmc(15, rng=np.random)
    after_oracle = np.random.random_sample()
    np.random.seed(7)
    model.do_mcmc(15, prog=False)
    after_gpu = np.random.random_sample()
    assert after_gpu == after_oracle
    ref = om.get_samples()
    got = model.get_samples()
    # same accept decisions at every site (a draw changes exactly where the oracle's does); values to the last few ulp:
    # the candidates come from the same uniforms through exp / log, where CUDA's libm and glibc may differ in the last
    # bit (under replayed candidates the chains are bit-identical: tests/test_gpu_chains.py)
    for k in ('betaU', 'lamUz', 'lamWs', 'lamWOs'):
        assert got[k].shape == ref[k].shape
        assert np.array_equal(np.diff(got[k], axis=0) != 0, np.diff(ref[k], axis=0) != 0)
        np.testing.assert_allclose(got[k], ref[k], rtol=1e-12)
    np.testing.assert_allclose(got['logPost'], ref['logPost'], rtol=1e-10)
    assert model.get_samples(5, nburn=3)['betaU'].shape == (5, 8)
    np.testing.assert_allclose(model.logPost(), ref['logPost'][-1, 0], rtol=1e-7)


def test_reference_fit_recipe_and_prediction(cuda, tmp_path):
    """src/model.py:218-238 recipe (lamWOs override, tune, mcmc, save/restore) then the
    assess_all_models.py:468-492 prediction pattern; moments vs oracle w_pred."""
    from sepia.SepiaPredict import SepiaEmulatorPrediction
    from gladsgp_b200 import model as gmodel
    pr = make_problem(m=64, q=3, pu=2)
    data, model = _build(pr, 3)
    pc_prec = gmodel.pc_precision(data.sim_data)
    w = np.dot(np.linalg.pinv(data.sim_data.K).T, data.sim_data.y_std.T).T
    ref_prec = 1.0 / np.var(data.sim_data.y_std - np.dot(w, data.sim_data.K))
    assert abs(pc_prec - ref_prec) < 1e-4 * ref_prec
    gmodel.override_lamWOs(model, pc_prec)
    np.random.seed(3)
    model.tune_step_sizes(20, 5, prog=False)
    assert np.all(model.params.betaU.mcmc.stepParam > 0)
    model.do_mcmc(40, prog=False)
    path = os.path.join(tmp_path, 'mod')
    model.save_model_info(path)
    import pickle
    raw = pickle.load(open(path + '.pkl', 'rb'))
    assert np.array(raw['samples']['betaU']).shape == (40, 4, 2)
    data2, model2 = _build(pr, 3)
    model2.restore_model_info(path)
    samples = model2.get_samples(8, nburn=10)
    for key in samples.keys():
        samples[key] = samples[key].astype(np.float32)
    tp = synthetic.test_design(4, 3)
    np.random.seed(1)
    preds = SepiaEmulatorPrediction(t_pred=tp, samples=samples, model=model2, storeMuSigma=True)
    assert preds.w.shape == (8, 4, 2)
    num = so.OracleNum(pr['t'], data.sim_data.y_std, pr['K'])
    _, mu, Sig = so.w_pred(num, tp, samples, draw=False)
    np.testing.assert_allclose(preds.mu, mu, rtol=1e-5, atol=1e-5 * np.abs(mu).max())
    np.testing.assert_allclose(preds.sigma, Sig, rtol=1e-5, atol=1e-5 * np.abs(Sig).max())
    preds.w = preds.w.astype(np.float32)
    y = preds.get_y()
    assert y.shape == (8, 4, pr['y'].shape[1]) and y.dtype == np.float32
    ref = so.get_y(preds.w, pr['K'], pr['sd'].astype(np.float32), pr['mu'].astype(np.float32))
    np.testing.assert_allclose(y, ref, rtol=1e-4, atol=1e-4)


def test_realisations_have_the_right_distribution(cuda):
    from sepia.SepiaPredict import SepiaEmulatorPrediction
    pr = make_problem(m=64, q=3, pu=2)
    data, model = _build(pr, 3)
    s1 = synthetic.posterior_samples(1, 4, 2, seed=2)
    samples = {k: np.repeat(v, 4000, axis=0) for k, v in s1.items()}
    tp = synthetic.test_design(3, 3)
    np.random.seed(5)
    preds = SepiaEmulatorPrediction(t_pred=tp, samples=samples, model=model, storeMuSigma=True)
    mu, Sig = preds.mu[0], preds.sigma[0]
    w = preds.w.transpose(0, 2, 1).reshape(4000, -1)          # (pu, npred) PC-major like mu
    emp_mu = w.mean(axis=0); emp_cov = np.cov(w.T)
    sd = np.sqrt(np.diag(Sig))
    assert np.all(np.abs(emp_mu - mu) < 5 * sd / np.sqrt(4000))
    np.testing.assert_allclose(emp_cov, Sig, atol=0.15 * np.outer(sd, sd).max())


def test_joint_realisation_is_mean_plus_cholesky_factor_times_the_numpy_draws(cuda):
    """pred.w with the default joint draw: with the global stream seeded, the realisation is mu + L z for the SAME
    np.random.normal values SEPIA consumes (one vector of npred*pu values per sample, in sample order), L the Cholesky
    factor of the per-PC block of Sigma (ggp_chol_draw_f64)."""
    from sepia.SepiaPredict import SepiaEmulatorPrediction
    pr = make_problem(m=64, q=3, pu=2)
    data, model = _build(pr, 3)
    samples = synthetic.posterior_samples(5, 4, 2, seed=3)
    tp = synthetic.test_design(6, 3)
    np.random.seed(99)
    preds = SepiaEmulatorPrediction(t_pred=tp, samples=samples, model=model, storeMuSigma=True)
    mu, sigma = preds.get_mu_sigma()
    ns, npred, pu = preds.w.shape
    np.random.seed(99)
    z = np.random.normal(size=ns * npred * pu).reshape(ns, pu, npred)
    for s in range(ns):
        for j in range(pu):
            sl = slice(j * npred, (j + 1) * npred)
            L = np.linalg.cholesky(sigma[s, sl, sl])
            np.testing.assert_allclose(preds.w[s, :, j], mu[s, sl] + L @ z[s, j], rtol=1e-9, atol=1e-11)


def test_scalar_model_api(cuda):
    """fit_scalar_models.py:45-47,471-483 pattern (pu = 1, default priors, scalar standardisation)."""
    from sepia.SepiaData import SepiaData
    from sepia.SepiaModel import SepiaModel
    from sepia.SepiaPredict import SepiaEmulatorPrediction
    pr = make_scalar_problem(m=100)
    data = SepiaData(t_sim=pr['t'], y_sim=pr['y'])
    data.transform_xt()
    data.standardize_y()
    model = SepiaModel(data)
    assert str(data).splitlines()[-1] == 'pu =     1 (univariate response dimension)'
    np.random.seed(2)
    model.do_mcmc(30, prog=False)
    samples = model.get_samples(nburn=10, numsamples=8)
    tp = np.linspace(0.05, 0.95, 50)[:, None]
    preds = SepiaEmulatorPrediction(model=model, t_pred=tp, samples=samples, storeMuSigma=True)
    y = preds.get_y()
    assert y.shape == (8, 50, 1)
    om_num = so.OracleNum(data.sim_data.t_trans, data.sim_data.y_std, None)
    tt = (tp - data.sim_data.orig_t_min) / (data.sim_data.orig_t_max - data.sim_data.orig_t_min)
    _, mu, Sig = so.w_pred(om_num, tt, samples, draw=False)
    np.testing.assert_allclose(preds.mu, mu, rtol=1e-6, atol=1e-6 * np.abs(mu).max())
    np.testing.assert_allclose(preds.sigma, Sig, rtol=1e-6, atol=1e-6 * np.abs(Sig).max())


def test_fixed_parameter_and_uniform_prior(cuda):
    """include_trunc_error.py:76-83 pattern."""
    from sepia.SepiaPrior import SepiaPrior
    pr = make_problem(m=64, q=3, pu=2)
    data, model = _build(pr, 3)
    var = 77.0
    model.params.lamWOs.fixed = np.array([[True]])
    model.params.lamWOs.val = np.array([[var]])
    model.params.lamWOs.prior = SepiaPrior(model.params.lamWOs, dist='Uniform', params=[0, 2 * var], bounds=[0, 2 * var])
    np.random.seed(0)
    model.tune_step_sizes(10, 4, prog=False)
    model.do_mcmc(12, prog=False)
    s = model.get_samples()
    assert np.all(s['lamWOs'] == var)
    assert len(np.unique(s['betaU'][:, 0])) > 1


def test_randomized_svd_matches_reference_algorithm(cuda):
    from gladsgp_b200 import svd
    pr = make_problem(m=100, q=4, pu=3, n_x=60, n_t=50)
    X = pr['y_std'].astype(np.float32)
    rng = np.random.RandomState(4)
    omega = rng.normal(size=(X.shape[1], 25)).astype(np.float32)
    Ur, Sr, Vr = svd_oracle.randomized_svd(X, 25, k=0, q=1, omega=omega)
    U, S, Vh = svd.randomized_svd(X, 25, k=0, q=1, omega=omega)
    assert U.shape == Ur.shape and S.shape == Sr.shape and Vh.shape == Vr.shape
    np.testing.assert_allclose(S[:10], Sr[:10], rtol=2e-4)
    for i in range(5):           # leading vectors up to sign
        sgn = np.sign(np.dot(Vh[i], Vr[i]))
        np.testing.assert_allclose(sgn * Vh[i], Vr[i], atol=5e-3 * np.abs(Vr[i]).max() + 1e-4)
        np.testing.assert_allclose(sgn * U[:, i], Ur[:, i], atol=5e-3)
    rec = (U * S) @ Vh
    recr = (Ur * Sr) @ Vr
    assert np.linalg.norm(rec - recr) < 2e-3 * np.linalg.norm(recr)
    # global-stream consumption is the reference's: one normal(size=(n, p+k)) draw
    np.random.seed(9); svd.randomized_svd(X, 5, k=0, q=1); a = np.random.random_sample()
    np.random.seed(9); np.random.normal(size=(X.shape[1], 5)); b = np.random.random_sample()
    assert a == b


def test_tune_step_sizes_matches_oracle(cuda):
    """Same global stream -> same acceptance counts -> same selected step sizes as the oracle's tuner
    (10 warm-up steps, ladder default*2^e, do_propMH=False, logit fit, target 1/e)."""
    import contextlib, io
    pr = make_problem(m=64, q=3, pu=2)
    data, model = _build(pr, 3)
    om = so.OracleModel(so.OracleNum(pr['t'], data.sim_data.y_std, pr['K']))
    om.lamWOs.params = [p.copy() for p in model.params.lamWOs.prior.params]
    np.random.seed(11)
    om.tune_step_sizes(12, 4, rng=np.random)
    after_oracle = np.random.random_sample()
    np.random.seed(11)
    with contextlib.redirect_stdout(io.StringIO()) as out:
        model.tune_step_sizes(12, 4, prog=False)
    after_gpu = np.random.random_sample()
    assert after_gpu == after_oracle
    txt = out.getvalue()
    assert txt.startswith('Starting tune_step_sizes...\nDefault step sizes:\nbetaU\n')      # examples/04_*.ipynb:297-300
    assert 'Done with tune_step_size.\nSelected step sizes:\n' in txt                        # :327-328
    for name in ('betaU', 'lamUz', 'lamWs', 'lamWOs'):
        np.testing.assert_allclose(getattr(model.params, name).mcmc.stepParam, getattr(om, name).step, rtol=1e-6)
        np.testing.assert_allclose(getattr(model.params, name).val, getattr(om, name).val, rtol=1e-7)


def test_parallel_step_size_tuning_equals_one_chain_per_level(cuda):
    """tune_step_sizes(parallel=True): the n_levels trial sizes run as simultaneous chains; level l's
    acceptance record is, bit for bit, that of a single chain run with trial size l on its slice of the np.random stream,
    so the selected step sizes are those of n_levels separate runs."""
    import contextlib, io, copy
    pr = make_problem(m=64, q=3, pu=2)
    data, model = _build(pr, 3)
    n_burn, n_levels = 12, 4
    _, tb = model._get_engine(1)
    P = tb['theta'].size
    ladder = tb['step'][None, :] * np.power(2.0, np.linspace(-(n_levels - 1) / 2.0, (n_levels - 1) / 2.0, n_levels))[:, None]
    # reference: warm-up, then one run per level from the warmed-up state with the matching slice of the stream
    np.random.seed(21)
    draws, lps, _ = model._run(10, tb['step'], do_propMH=False)
    model._store_state(draws[-1, 0, :])
    warmed = [b.val.copy() for b in model._blocks()[0]]
    state = np.random.get_state()
    acc_ref = []
    for lvl in range(n_levels):
        for b, v in zip(model._blocks()[0], warmed):
            b.val = v.copy()
        np.random.set_state(state)
        np.random.random_sample(lvl * 2 * P * n_burn)
        _, _, acc = model._run(n_burn, ladder[lvl], do_propMH=False, record_accept=True)
        acc_ref.append(acc[:, 0, :].astype(np.int64).sum(axis=0))
    acc_ref = np.array(acc_ref)
    data2, model2 = _build(pr, 3)
    np.random.seed(21)
    with contextlib.redirect_stdout(io.StringIO()):
        lad, acc = model2.tune_step_sizes(n_burn, n_levels, prog=False, diagnostics=True, parallel=True)
    np.testing.assert_array_equal(lad, ladder)
    np.testing.assert_array_equal(acc, acc_ref)
    for name in ('betaU', 'lamUz', 'lamWs', 'lamWOs'):
        stp = getattr(model2.params, name).mcmc.stepParam
        assert np.all(np.isfinite(stp)) and np.all(stp > 0)


def test_batched_chains_equal_independent_single_chains(cuda):
    """Chains are independent units: a batched run of 3 chains gives, bit for bit, the 3 chains
    obtained one at a time from the same per-chain uniform streams."""
    from gladsgp_b200 import ops
    pr = make_problem(m=64, q=3, pu=2)
    data, model = _build(pr, 3)
    tb = model._tables()
    P = tb['theta'].size
    n_steps, C = 8, 3
    streams = np.random.RandomState(5).random_sample((C, 2 * P * n_steps))
    eng = ops.McmcEngine(model.num.zt, model._w_pcs, model.num.LamSim, tb, n_chains=C)
    eng.set_state(tb['theta'])
    out = eng.run(n_steps, tb['step'], uniforms=streams)
    draws = out['draws'].cpu().numpy(); lp = out['lp'].cpu().numpy(); used = out['consumed'].cpu().numpy()
    one = ops.McmcEngine(model.num.zt, model._w_pcs, model.num.LamSim, tb, n_chains=1)
    for c in range(C):
        one.set_state(tb['theta'])
        o = one.run(n_steps, tb['step'], uniforms=streams[c:c + 1])
        assert np.array_equal(o['draws'].cpu().numpy()[:, 0], draws[:, c])
        assert np.array_equal(o['lp'].cpu().numpy()[:, 0], lp[:, c])
        assert int(o['consumed'].cpu().numpy()[0]) == int(used[c])
    assert not np.array_equal(draws[:, 0], draws[:, 1])
    # public multi-chain entry point
    np.random.seed(1)
    d2, l2 = model.do_mcmc_chains(4, 2)
    assert d2.shape == (4, 2, P) and l2.shape == (4, 2) and np.all(np.isfinite(l2))


def test_reference_workflow_end_to_end(cuda, tmp_path):
    """fit_models -> load_model -> batched prediction, reduced cfg 1 shape (examples/synthetic_fit_predict.py)."""
    import subprocess, sys
    from helpers import ROOT
    out = subprocess.run([sys.executable, os.path.join(ROOT, 'examples', 'synthetic_fit_predict.py'), str(tmp_path),
                          '--m', '64', '--nx', '60', '--nt', '12', '--mcmc', '96', '--tune', '20', '--ntest', '8'],
                         capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    files = os.listdir(os.path.join(tmp_path, 'data', 'models'))
    for f in ('pca_synthetic_n064_U.npy', 'pca_synthetic_n064_S.npy', 'pca_synthetic_n064_Vh.npy',
              'synthetic_n064_p05.pkl', 'timing.csv'):
        assert f in files, files
    line = [l for l in out.stdout.splitlines() if l.startswith('test RMSE')][0]
    rmse = float(line.split()[2]); cover = float(line.split()[-1])
    assert rmse < 0.2 and 0.6 < cover <= 1.0, line


def test_model_batch_equals_models_fitted_one_by_one(cuda):
    """gladsgp_b200.batch.ModelBatch (the per-threshold scalar models of fit_scalar_models.py
    fitted together): every model reaches exactly the state it reaches alone with its own seeded stream."""
    import copy
    import io
    from contextlib import redirect_stdout
    from sepia.SepiaData import SepiaData
    from sepia.SepiaModel import SepiaModel
    from gladsgp_b200.batch import ModelBatch
    rng = np.random.default_rng(4)
    m, q = 48, 3
    t = synthetic.design(m, q, seed=9)
    ys = [np.sin(3 * t[:, 0]) + 0.1 * rng.standard_normal(m), t[:, 1] ** 2 - t[:, 2] + 0.05 * rng.standard_normal(m),
          np.cos(2 * t[:, 2]) * t[:, 0] + 0.2 * rng.standard_normal(m)]

    def make(y):
        d = SepiaData(t_sim=t, y_sim=y.astype(np.float32))
        d.transform_xt()
        d.standardize_y()
        return SepiaModel(d)

    seeds = [11, 22, 33]
    solo = [make(y) for y in ys]
    with redirect_stdout(io.StringIO()):
        for mm, s in zip(solo, seeds):
            np.random.seed(s)
            mm.tune_step_sizes(4, 3, prog=False)
            mm.do_mcmc(6, prog=False)
    together = [make(y) for y in ys]
    batch = ModelBatch(together, seeds=seeds)
    batch.tune_step_sizes(4, 3)
    batch.do_mcmc(6)
    for a, b in zip(solo, together):
        sa, sb = a.get_samples(), b.get_samples()
        for k in ('betaU', 'lamUz', 'lamWs', 'lamWOs', 'logPost'):
            np.testing.assert_allclose(sb[k], sa[k], rtol=1e-12, atol=0)
        for pa, pb in zip(a.params.mcmcList, b.params.mcmcList):
            np.testing.assert_allclose(pb.mcmc.stepParam, pa.mcmc.stepParam, rtol=1e-12)
            np.testing.assert_allclose(pb.val, pa.val, rtol=1e-12)
    # a multivariate batch: two ensembles on one design (own K, own lamWOs prior each)
    pr1 = make_problem(m=40, q=3, pu=2, seed=5)
    pr2 = make_problem(m=40, q=3, pu=2, seed=5)
    pr2['y'] = (pr2['y'] * 1.3 + 0.05 * np.random.default_rng(1).standard_normal(pr2['y'].shape)).astype(pr2['y'].dtype)
    solo, together = [], []
    for pr in (pr1, pr2):
        for lst in (solo, together):
            d = SepiaData(t_sim=pr['t'], y_sim=pr['y'], y_ind_sim=np.linspace(0, 1, pr['y'].shape[1]))
            d.transform_xt(); d.standardize_y(); d.create_K_basis(n_pc=2)
            lst.append(SepiaModel(d))
    for mm, s in zip(solo, (5, 6)):
        np.random.seed(s)
        mm.do_mcmc(5, prog=False)
    ModelBatch(together, seeds=(5, 6)).do_mcmc(5)
    for a, b in zip(solo, together):
        sa, sb = a.get_samples(), b.get_samples()
        for k in ('betaU', 'lamUz', 'lamWs', 'lamWOs', 'logPost'):
            np.testing.assert_allclose(sb[k], sa[k], rtol=1e-12, atol=0)
