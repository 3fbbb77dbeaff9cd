"""The reference's own driver code run, unmodified, against this repository's `sepia` mirror.

/root/reference/src/model.py is imported as it is (its `from sepia...` lines resolve against the top-level `sepia/` shim;
its PCA goes through its own src/svd.py).  The tests need the reference tree, which exists in the build container but not on
the GPU box, so:
  * CPU (here): `init_model` (src/model.py:20-107) runs end to end on the host path of the mirror, and `fit_models`
    (src/model.py:152-245) gets as far as its first compute call, where the product fails LOUDLY -- there is no CPU fallback;
  * GPU (skipped where the reference tree is absent): `fit_models`, `load_model` (src/model.py:109-150) and the prediction
    loop of experiments/synthetic/analysis/assess_all_models.py:468-500, statement for statement.
"""
import os
import sys
import types

import numpy as np
import pytest

from helpers import ROOT, synthetic

REF = '/root/reference'
needs_ref = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, 'src')), reason='reference tree not present')


def _ref_model():
    for p in (REF, ROOT):
        if p in sys.path:
            sys.path.remove(p)
    sys.path.insert(0, ROOT)           # `sepia` -> this repository's mirror
    sys.path.insert(1, REF)            # `src`   -> the reference's drivers
    import importlib
    return importlib.import_module('src.model')


def _train_config(tmp_path, m=40, q=3, n_x=12, n_t=6):
    """A train_config module as experiments/synthetic/train_config.py defines it: X_standard (csv with a header line),
    Y_physical ((n_y, m) .npy, src/aggregate_outputs.py:61-68), data_dir, exp."""
    t = synthetic.design(m, q, seed=7)
    y = synthetic.ensemble(t, n_x=n_x, n_t=n_t, seed=7)
    xs = os.path.join(tmp_path, 'X_standard.csv')
    np.savetxt(xs, t, delimiter=',', header=','.join('p%d' % i for i in range(q)), comments='')
    yp = os.path.join(tmp_path, 'Y_physical.npy')
    np.save(yp, np.ascontiguousarray(y.T))
    cfg = types.SimpleNamespace(X_standard=xs, Y_physical=yp, data_dir=str(tmp_path), exp='synth')
    return cfg, t, y


@needs_ref
def test_init_model_runs_unmodified_on_the_host_path(tmp_path):
    rm = _ref_model()
    cfg, t, y = _train_config(tmp_path)
    np.random.seed(0)
    data, model = rm.init_model(t_std=t.astype(np.float32), y_sim=y.astype(np.float32), exp='synth_n040', p=3,
                                data_dir=os.path.join(tmp_path, 'models'))
    assert type(data).__module__.startswith('gladsgp_b200.sepia') and type(model).__module__.startswith('gladsgp_b200.sepia')
    assert data.sim_data.K.shape == (3, y.shape[1]) and data.sim_data.K.dtype == np.float32
    assert model.num.w.shape == (40 * 3, 1) and model.num.LamSim.shape == (3,)
    for arr in ('U', 'S', 'Vh'):
        assert os.path.exists(os.path.join(tmp_path, 'models', 'pca_synth_n040_%s.npy' % arr))     # src/model.py:87-94
    w = np.dot(np.linalg.pinv(data.sim_data.K).T, data.sim_data.y_std.T).T                        # src/model.py:219
    np.testing.assert_allclose(model.num.w.reshape(3, 40).T, w, rtol=1e-4, atol=1e-5)
    assert str(data).startswith('This SepiaData instance implies the following:')


@needs_ref
def test_fit_models_fails_loudly_without_a_gpu(tmp_path):
    import torch
    if torch.cuda.is_available():
        pytest.skip('a GPU is present: see test_fit_load_predict_verbatim')
    from gladsgp_b200._lib import GgpError
    rm = _ref_model()
    cfg, t, y = _train_config(tmp_path)
    np.random.seed(0)
    with pytest.raises(GgpError, match='no CPU fallback'):
        rm.fit_models(cfg, [40], [3])


@pytest.mark.gpu
@needs_ref
def test_fit_load_predict_verbatim(cuda, tmp_path):
    from sepia.SepiaPredict import SepiaEmulatorPrediction
    rm = _ref_model()
    cfg, t, y = _train_config(tmp_path)
    np.random.seed(0)
    models = rm.fit_models(cfg, [40], [2, 3])                                  # src/model.py:152-245
    assert len(models) == 2
    timing = np.loadtxt(os.path.join(tmp_path, 'models', 'timing.csv'), delimiter=',', skiprows=1)
    assert timing.shape == (2, 4)
    sepia_data, model = rm.load_model(cfg, 40, 3)                               # src/model.py:109-150
    # ---- assess_all_models.py:463-500, statement for statement (dtype = float32, quantile = 0.025) ----
    dtype, quantile = np.float32, 0.025
    mu_y = np.mean(model.data.sim_data.y, axis=0)
    sd_y = np.std(model.data.sim_data.y, ddof=1, axis=0)
    sd_y[sd_y < 1e-6] = 1e-6
    samples = model.get_samples(6, nburn=0)
    for key in samples.keys():
        samples[key] = samples[key].astype(dtype)
    x_pred = synthetic.test_design(6, 3)
    n_per_batch = 4
    n_batches = int(np.ceil(len(x_pred) / n_per_batch))
    batch_indices = np.array_split(np.arange(len(x_pred)), n_batches)
    ypred_mean = np.zeros((len(x_pred), y.shape[1]), dtype=dtype)
    ypred_lq = np.zeros_like(ypred_mean); ypred_uq = np.zeros_like(ypred_mean)
    for j in range(n_batches):
        tj_pred = x_pred[batch_indices[j], :]
        preds = SepiaEmulatorPrediction(t_pred=tj_pred, samples=samples, model=model)
        preds.w = preds.w.astype(np.float32)
        ypreds = preds.get_y()
        error_preds = np.zeros(ypreds.shape, dtype=np.float32)
        for l_pred in range(len(batch_indices[j])):
            for l_sample in range(error_preds.shape[0]):
                err_sd = 1 / np.sqrt(samples['lamWOs'][l_sample])
                error_preds[l_sample][l_pred] = sd_y * np.random.normal(scale=err_sd)
        ypred_mean[batch_indices[j]] = np.mean(ypreds, axis=0)
        ypred_lq[batch_indices[j]] = np.quantile(ypreds + error_preds, quantile, axis=0)
        ypred_uq[batch_indices[j]] = np.quantile(ypreds + error_preds, 1 - quantile, axis=0)
    assert ypreds.shape == (6, len(batch_indices[-1]), y.shape[1]) and ypreds.dtype == np.float32
    assert np.all(np.isfinite(ypred_mean)) and np.all(ypred_lq <= ypred_uq)
    # the emulator interpolates its own training design: mean prediction at training points close to the data
    preds = SepiaEmulatorPrediction(t_pred=t[:4].astype(np.float32), samples=samples, model=model)
    preds.w = preds.w.astype(np.float32)
    fit = np.mean(preds.get_y(), axis=0)
    assert np.sqrt(np.mean((fit - y[:4]) ** 2)) < 0.5 * np.std(y)
