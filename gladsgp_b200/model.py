"""Fit / load drivers mirroring /root/reference/src/model.py (init_model :20-107, load_model :109-150,
fit_models :152-245) on the B200 path.  Same arguments, same files written (pca_{exp}_{U,S,Vh}.npy,
{exp}_n{m}_p{p}.pkl, timing.csv); the reference's own src/model.py also runs unchanged against the
`sepia` package of this repository."""
import os
import time

import numpy as np

from sepia.SepiaModel import SepiaModel
from sepia.SepiaData import SepiaData
from sepia import SepiaParam

from . import svd, ingest

PMAX = 25      # src/model.py:81


def init_model(t_std, y_sim, exp, p, data_dir='data/', sd_threshold=1e-6, recompute=False):
    """src/model.py:20-107.  y_sim may be the transposed view of the (n_y, m) ensemble file (np.load(path).T[:m],
    no host copy): large float32 ensembles are uploaded once as stored and every pass over them (column statistics,
    standardisation, rSVD, projection) runs on the device."""
    y_ind_sim = np.linspace(0, 1, y_sim.shape[1])
    data = SepiaData(t_sim=t_std, y_sim=y_sim, y_ind_sim=y_ind_sim)
    on_device = ingest.use_device(y_sim.size) and y_sim.dtype == np.float32
    if on_device:
        ydev, tr = data.sim_data.y_device()
        mu_d, sd_d = ingest.column_stats(ydev, transposed=tr, sd_threshold=sd_threshold)      # src/model.py:60-64
        mu_y, sd_y = mu_d.cpu().numpy(), sd_d.cpu().numpy()
    else:
        mu_y = np.mean(y_sim, axis=0)
        sd_y = np.std(y_sim, ddof=1, axis=0)
        sd_y[sd_y < sd_threshold] = sd_threshold
    data.transform_xt(t_notrans=np.arange(t_std.shape[1]))
    data.standardize_y(y_mean=mu_y, y_sd=sd_y)
    os.makedirs(data_dir, exist_ok=True)
    pat = os.path.join(data_dir, 'pca_{}_{}.npy')
    have = all(os.path.exists(pat.format(exp, a)) for a in ('U', 'S', 'Vh'))
    if recompute or not have:
        U, S, Vh = svd.randomized_svd(data.sim_data.y_std_device() if on_device else data.sim_data.y_std,
                                      PMAX, k=0, q=1)
        np.save(pat.format(exp, 'U'), U[:, :PMAX])
        np.save(pat.format(exp, 'S'), S)
        np.save(pat.format(exp, 'Vh'), Vh[:PMAX, :])
    S = np.load(pat.format(exp, 'S'))
    Vh = np.load(pat.format(exp, 'Vh'))
    S2 = S ** 2
    print('SVD proportion of variance:', (S2 / np.sum(S2))[:10])
    K = (S[:p, None] * Vh[:p]) / np.sqrt(y_sim.shape[0])         # diag(S[:p]) @ Vh[:p] / sqrt(m)
    data.create_K_basis(K=K.astype(np.float32))
    print('K.shape', K.shape)
    return data, SepiaModel(data)


def pc_precision(sim_data):
    """src/model.py:219-223: 1 / var(y_std - w K).  Large ensembles: from the device projection pass (shared with
    SepiaModel.__init__); small ones: streamed over column blocks on the host."""
    if getattr(sim_data, '_proj', None) is not None:
        return ingest.pc_precision_from(sim_data._proj)
    if (ingest.use_device(sim_data.y.size) and np.asarray(sim_data.K).dtype == np.float32 and
            (sim_data._y_std_dev is not None or sim_data._y_std.dtype == np.float32)):
        sim_data._proj = ingest.project_basis(sim_data.y_std_device(), sim_data.K_device())
        return ingest.pc_precision_from(sim_data._proj)
    K = np.asarray(sim_data.K, dtype=np.float64)
    G = K @ K.T
    n = 0
    s1 = s2 = 0.0
    ys = sim_data.y_std
    blk = 1 << 16
    YK = np.zeros((ys.shape[0], K.shape[0]))
    for c0 in range(0, ys.shape[1], blk):
        YK += ys[:, c0:c0 + blk].astype(np.float64) @ K[:, c0:c0 + blk].T
    w = np.linalg.solve(G, YK.T).T
    for c0 in range(0, ys.shape[1], blk):
        r = ys[:, c0:c0 + blk].astype(np.float64) - w @ K[:, c0:c0 + blk]
        s1 += r.sum(); s2 += (r * r).sum(); n += r.size
    return 1.0 / (s2 / n - (s1 / n) ** 2)


def override_lamWOs(model, pc_prec, gamma_a=50):
    """src/model.py:225-231."""
    model.params.lamWOs = SepiaParam(val=pc_prec, name='lamWOs', val_shape=(1, 1), dist='Gamma',
                                     params=[gamma_a, gamma_a / pc_prec], bounds=[1., np.inf],
                                     mcmcStepParam=10, mcmcStepType='Uniform')
    model.params.mcmcList = [model.params.betaU, model.params.lamUz, model.params.lamWs, model.params.lamWOs]


def _load_ensemble(path, dtype):
    """np.load(path).T.astype(dtype) (src/model.py:133,185) without the host transpose when the file already has the
    requested dtype: the (m, n_y) result is then a view of the (n_y, m) file array and is transposed on the device."""
    a = np.load(path)
    return a.T if a.dtype == dtype else a.T.astype(dtype)


def load_model(train_config, m, p, dtype=np.float32):
    t_std = np.loadtxt(train_config.X_standard, delimiter=',', skiprows=1, comments=None).astype(dtype)[:m]
    y_sim = _load_ensemble(train_config.Y_physical, dtype)[:m]
    data_dir = os.path.join(train_config.data_dir, 'models')
    os.makedirs(data_dir, exist_ok=True)
    model_path = os.path.join(data_dir, '{}_n{:03d}_p{:02d}'.format(train_config.exp, m, p))
    data, model = init_model(t_std=t_std, y_sim=y_sim, exp='{}_n{:03d}'.format(train_config.exp, m), p=p,
                             data_dir=data_dir, recompute=False)
    print('Restoring from:', model_path)
    model.restore_model_info(model_path)
    return data, model


def fit_models(train_config, n_sims, n_pcs, dtype=np.float32, recompute=False, n_tune=(100, 5), n_mcmc=512):
    t_std = np.loadtxt(train_config.X_standard, delimiter=',', skiprows=1, comments=None).astype(dtype)
    y_sim = _load_ensemble(train_config.Y_physical, dtype)
    data_dir = os.path.join(train_config.data_dir, 'models')
    os.makedirs(data_dir, exist_ok=True)
    models, rows = [], []
    for m in n_sims:
        for p in n_pcs:
            t0 = time.perf_counter()
            data, model = init_model(t_std=t_std[:m], y_sim=y_sim[:m], exp='{}_n{:03d}'.format(train_config.exp, m),
                                     p=p, data_dir=data_dir, recompute=recompute)
            t1 = time.perf_counter()
            pc_prec = pc_precision(data.sim_data)
            print('PC PRECISION:', pc_prec)
            override_lamWOs(model, pc_prec)
            t2 = time.perf_counter()
            model.tune_step_sizes(*n_tune)
            model.do_mcmc(n_mcmc)
            t3 = time.perf_counter()
            model.save_model_info(os.path.join(data_dir, '{}_n{:03d}_p{:02d}'.format(train_config.exp, m, p)))
            models.append(model)
            rows.append([m, p, t1 - t0, t3 - t2])
    np.savetxt(os.path.join(data_dir, 'timing.csv'), np.array(rows), delimiter=',', fmt='%.3f',
               header='sims,PCs,PCA (seconds),MCMC (seconds)')
    return models
