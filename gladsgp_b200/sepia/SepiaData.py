"""SepiaData mirror, sim-only.

Call sites: /root/reference/src/model.py:57,68,71,102;
experiments/synthetic/analysis/fit_scalar_models.py:45-47; sensitivity_indices.py:183.
Pure host-side container logic (no heavy arithmetic): input scaling, y standardisation, K basis.
"""
import numpy as np


class DataContainer:
    def __init__(self, x, y, t=None, y_ind=None):
        self.x = x
        self.y = y
        self.t = t
        self.y_ind = y_ind
        self.x_trans = None
        self.t_trans = None
        self.y_std = None
        self.K = None
        self.orig_y_mean = None
        self.orig_y_sd = None
        self.orig_x_min = self.orig_x_max = None
        self.orig_t_min = self.orig_t_max = None


class SepiaData:
    def __init__(self, x_sim=None, t_sim=None, y_sim=None, y_ind_sim=None, x_obs=None, y_obs=None,
                 Sigy=None, y_ind_obs=None, theta_dim=None, x_cat_ind=None, t_cat_ind=None, xt_sim_sep=None):
        if y_obs is not None or x_obs is not None or Sigy is not None or y_ind_obs is not None:
            raise NotImplementedError('observation data / calibration is outside the GladsGP path '
                                      '(the reference only builds simulator-only models)')
        if xt_sim_sep is not None or x_cat_ind is not None or t_cat_ind is not None:
            raise NotImplementedError('Kronecker-separable designs and categorical inputs are outside the GladsGP path')
        if y_sim is None:
            raise TypeError('y_sim is required')
        if x_sim is None and t_sim is None:
            raise TypeError('at least one of x_sim, t_sim is required')
        y_sim = np.asarray(y_sim)
        if y_sim.ndim == 1:
            y_sim = y_sim[:, None]
        m = y_sim.shape[0]
        if t_sim is not None:
            t_sim = np.asarray(t_sim)
            if t_sim.ndim == 1:
                t_sim = t_sim[:, None]
            if t_sim.shape[0] != m:
                raise ValueError('Number of observations in t_sim and y_sim must be the same size')
        self.dummy_x = x_sim is None
        if x_sim is None:
            x_sim = 0.5 * np.ones((m, 1))
        else:
            x_sim = np.asarray(x_sim)
            if x_sim.ndim == 1:
                x_sim = x_sim[:, None]
            if x_sim.shape[0] != m:
                raise ValueError('Number of observations in x_sim and y_sim must be the same size')
        if y_sim.shape[1] > 1 and y_ind_sim is None:
            raise TypeError('y_ind_sim is required for multivariate y_sim')
        if y_ind_sim is not None and y_sim.shape[1] > 1 and np.asarray(y_ind_sim).shape[0] != y_sim.shape[1]:
            raise ValueError('y_ind_sim must have one entry per column of y_sim')
        self.sim_data = DataContainer(x=x_sim, y=y_sim, t=t_sim, y_ind=y_ind_sim)
        self.obs_data = None
        self.sim_only = True
        self.ragged_obs = False
        self.scalar_out = (y_sim.shape[1] == 1)
        self.x_cat_ind = np.zeros(x_sim.shape[1])
        self.t_cat_ind = np.zeros(0 if t_sim is None else t_sim.shape[1])

    # text recorded at examples/04_GP_emulation_multivariate_ensemble.ipynb:255-260 and 03_*:161-166
    def __str__(self):
        sd = self.sim_data
        res = 'This SepiaData instance implies the following:\n'
        res += 'This is a simulator (eta)-only model, y dimension %d\n' % sd.y.shape[1]
        res += 'm  = %5d (number of simulated data)\n' % sd.x.shape[0]
        res += 'p  = %5d (number of inputs)\n' % sd.x.shape[1]
        if sd.t is not None:
            res += 'q  = %5d (number of additional simulation inputs)\n' % sd.t.shape[1]
        if self.scalar_out:
            res += 'pu =     1 (univariate response dimension)\n'
        elif sd.K is not None:
            res += 'pu = %5d (transformed response dimension)\n' % sd.K.shape[0]
        else:
            res += 'pu NOT SET (transformed response dimension); call method create_K_basis \n'
        return res

    def transform_xt(self, x_notrans=None, t_notrans=None, x=None, t=None):
        """Column-wise [0,1] scaling; columns listed in *_notrans (or constant) are left untouched.
        With x / t given, returns those arrays transformed with the stored ranges (used at prediction)."""
        sd = self.sim_data

        def ranges(a, notrans):
            a = np.asarray(a)
            lo = np.min(a, axis=0, keepdims=True).astype(np.float64)
            hi = np.max(a, axis=0, keepdims=True).astype(np.float64)
            same = (hi - lo) == 0
            lo[same] = 0.0
            hi[same] = 1.0
            if notrans is not None:
                idx = np.asarray(notrans, dtype=int).reshape(-1)
                lo[:, idx] = 0.0
                hi[:, idx] = 1.0
            return lo, hi

        if x is None and t is None:
            sd.orig_x_min, sd.orig_x_max = ranges(sd.x, x_notrans)
            sd.x_trans = (sd.x - sd.orig_x_min) / (sd.orig_x_max - sd.orig_x_min)
            if sd.t is not None:
                sd.orig_t_min, sd.orig_t_max = ranges(sd.t, t_notrans)
                sd.t_trans = (sd.t - sd.orig_t_min) / (sd.orig_t_max - sd.orig_t_min)
            return None
        if sd.orig_x_min is None:
            self.transform_xt()
        xt = None if x is None else (np.asarray(x) - sd.orig_x_min) / (sd.orig_x_max - sd.orig_x_min)
        tt = None if t is None else (np.asarray(t) - sd.orig_t_min) / (sd.orig_t_max - sd.orig_t_min)
        return xt, tt

    def standardize_y(self, center=True, scale='scalar', y_mean=None, y_sd=None):
        sd = self.sim_data
        y = sd.y
        if y_mean is None:
            y_mean = np.mean(y, axis=0) if center else 0.0
        if y_sd is None:
            yc = y - y_mean
            if scale == 'scalar':
                y_sd = np.std(yc, ddof=1)
            elif scale == 'columnwise':
                y_sd = np.std(yc, ddof=1, axis=0)
            elif scale is False:
                y_sd = 1.0
            else:
                raise ValueError('scale must be "scalar", "columnwise" or False')
        sd.orig_y_mean = y_mean
        sd.orig_y_sd = y_sd
        sd.y_std = (y - y_mean) / y_sd

    def create_K_basis(self, n_pc=0.995, K=None):
        if self.scalar_out:
            if K is not None:
                raise ValueError('K basis is not used for univariate output')
            return
        sd = self.sim_data
        if K is not None:
            K = np.asarray(K)
            if K.ndim != 2 or K.shape[1] != sd.y.shape[1]:
                raise ValueError('K must have shape (pu, %d), got %s' % (sd.y.shape[1], K.shape))
            sd.K = K
            return
        if sd.y_std is None:
            self.standardize_y()
        # SEPIA default: SVD of y_std^T scaled U*s/sqrt(m) (the scaling src/model.py:100-101 refers to)
        m = sd.y_std.shape[0]
        U, s, _ = np.linalg.svd(np.asarray(sd.y_std, dtype=np.float64).T, full_matrices=False)
        if n_pc < 1:
            cv = np.cumsum(s ** 2) / np.sum(s ** 2)
            n_pc = int(np.searchsorted(cv, n_pc) + 1)
        n_pc = int(n_pc)
        sd.K = (U[:, :n_pc] * s[:n_pc] / np.sqrt(m)).T
