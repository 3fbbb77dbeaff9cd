"""SepiaData mirror, sim-only.

Call sites: /root/reference/src/model.py:57,68,71,102;
experiments/synthetic/analysis/fit_scalar_models.py:45-47; sensitivity_indices.py:183.
Host-side container logic (input scaling, K basis).  Large ensembles (>= ingest.DEVICE_MIN_ELEMS elements) are
standardised on the GPU (csrc/ggp_ingest.cu) and kept device-resident for the rSVD and projection passes; the host
copy of `y_std` is then only materialised when a caller reads it.
"""
import numpy as np

from .. import ingest


class DataContainer:
    """sim_data: x, t, y, y_ind, x_trans, t_trans, y_std, K, orig_y_mean / orig_y_sd (SEPIA's DataContainer).
    y_std and K carry an optional device copy (torch CUDA tensor) next to the host array."""

    def __init__(self, x, y, t=None, y_ind=None):
        self.x = x
        self.y = y
        self.t = t
        self.y_ind = y_ind
        self.x_trans = None
        self.t_trans = None
        self._y_std = None
        self._y_std_dev = None
        self._y_dev = None            # (tensor, transposed) raw ensemble, dropped once standardised
        self._K = None
        self._K_dev = None
        self._stat_dev = None         # (sd, mean) device copies for get_y
        self._proj = None             # cached projection on K (w, K K^T, residual sums)
        self.orig_y_mean = None
        self.orig_y_sd = None
        self.orig_x_min = self.orig_x_max = None
        self.orig_t_min = self.orig_t_max = None

    # -- y_std: host array, downloaded from the device copy on first read
    @property
    def y_std(self):
        if self._y_std is None and self._y_std_dev is not None:
            self._y_std = self._y_std_dev.cpu().numpy()
        return self._y_std

    @y_std.setter
    def y_std(self, v):
        self._y_std = v
        self._y_std_dev = None
        self._proj = None

    def has_y_std(self):
        return self._y_std is not None or self._y_std_dev is not None

    def y_std_device(self):
        if self._y_std_dev is None:
            self._y_std_dev = ingest.upload(self._y_std)
        return self._y_std_dev

    def y_device(self):
        """Raw ensemble on the device as (tensor, transposed).  An (n_y, m)-major array viewed through `.T` (the
        reference's np.load(Y_physical).T, src/model.py:133) is uploaded as stored and transposed on the device."""
        if self._y_dev is None:
            y = self.y
            if y.dtype == np.float32 and not y.flags['C_CONTIGUOUS'] and y.T.strides[1] == y.itemsize:
                self._y_dev = (ingest.upload(np.ascontiguousarray(y.T)), True)
            else:
                self._y_dev = (ingest.upload(y), False)
        return self._y_dev

    @property
    def K(self):
        return self._K

    @K.setter
    def K(self, v):
        self._K = v
        self._K_dev = None
        self._proj = None

    def K_device(self):
        if self._K_dev is None:
            self._K_dev = ingest.upload(self._K)
        return self._K_dev

    def stats_device(self):
        """(orig_y_sd, orig_y_mean) as flat float32 device vectors (length 1 or n_y)."""
        key = (id(self.orig_y_sd), id(self.orig_y_mean))
        if self._stat_dev is None or self._stat_dev[0] != key:
            self._stat_dev = (key, ingest.upload(np.asarray(self.orig_y_sd, dtype=np.float32).reshape(-1)),
                              ingest.upload(np.asarray(self.orig_y_mean, dtype=np.float32).reshape(-1)))
        return self._stat_dev[1], self._stat_dev[2]

    def __getstate__(self):           # device tensors do not travel in pickles
        st = dict(self.__dict__)
        if st.get('_y_std') is None and st.get('_y_std_dev') is not None:
            st['_y_std'] = st['_y_std_dev'].cpu().numpy()
        for k in ('_y_std_dev', '_y_dev', '_K_dev', '_stat_dev'):
            st[k] = None
        return st


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
        if ingest.use_device(y.size) and y.dtype == np.float32:
            return self._standardize_y_device(center, scale, y_mean, y_sd)
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
        sd._stat_dev = None
        sd.y_std = (y - y_mean) / y_sd

    def _standardize_y_device(self, center, scale, y_mean, y_sd):
        """Same rules on the GPU (ggp_colstats_f32 / ggp_standardize_f32); y_std stays on the device."""
        from .. import ops
        sd = self.sim_data
        m, n = sd.y.shape
        ydev, tr = sd.y_device()
        if scale not in ('scalar', 'columnwise', False):
            raise ValueError('scale must be "scalar", "columnwise" or False')
        if y_mean is None or y_sd is None:
            cm, cs = ops.colstats(ydev, transposed=tr, ddof=1, sd_floor=0.0)
            if y_mean is None:
                y_mean = cm.cpu().numpy() if center else 0.0
            if y_sd is None:
                if scale == 'columnwise':
                    y_sd = cs.cpu().numpy()
                elif scale is False:
                    y_sd = 1.0
                else:
                    # np.std(y - y_mean, ddof=1) over all elements, from the column sums
                    cmh = cm.double().cpu().numpy()
                    ss = float(np.sum(cs.double().cpu().numpy() ** 2) * (m - 1))
                    # between-column part: column means of the residual y - y_mean about their grand mean (zero when
                    # y_mean is the column mean; non-zero for a scalar, a user-supplied vector, or center=False)
                    resid_mean = cmh - np.asarray(y_mean, dtype=np.float64).reshape(-1)
                    ss += m * float(np.sum((resid_mean - resid_mean.mean()) ** 2))
                    y_sd = np.float32(np.sqrt(ss / (m * n - 1)))
        mean_h = np.asarray(y_mean, dtype=np.float32).reshape(-1)
        sd_h = np.asarray(y_sd, dtype=np.float32).reshape(-1)
        if mean_h.size not in (1, n) or sd_h.size not in (1, n):
            raise ValueError('y_mean / y_sd must be scalars or have one entry per column of y_sim')
        sd.orig_y_mean = y_mean
        sd.orig_y_sd = y_sd
        sd._stat_dev = ((id(y_sd), id(y_mean)), ingest.upload(sd_h), ingest.upload(mean_h))
        sd._y_std = None
        sd._proj = None
        sd._y_std_dev = ops.standardize(ydev, sd._stat_dev[2], sd._stat_dev[1], transposed=tr)
        sd._y_dev = None              # the raw device copy is not needed again

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
        if not sd.has_y_std():
            self.standardize_y()
        # SEPIA default: SVD of y_std^T scaled U*s/sqrt(m) (the scaling src/model.py:100-101 refers to)
        m = sd.y_std.shape[0]
        U, s, _ = np.linalg.svd(np.asarray(sd.y_std, dtype=np.float64).T, full_matrices=False)
        if n_pc < 1:
            cv = np.cumsum(s ** 2) / np.sum(s ** 2)
            n_pc = int(np.searchsorted(cv, n_pc) + 1)
        n_pc = int(n_pc)
        sd.K = (U[:, :n_pc] * s[:n_pc] / np.sqrt(m)).T
