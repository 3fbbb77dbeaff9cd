"""SepiaParam mirror (constructed by the reference at /root/reference/src/model.py:227-229)."""
import numpy as np

from .SepiaPrior import SepiaPrior
from .SepiaMCMC import SepiaMCMC


class SepiaParam:
    def __init__(self, val, name, val_shape=(1, 1), dist='Normal', params=[], bounds=[],
                 mcmcStepParam=0.1, mcmcStepType='Normal', orig_range=None, fixed=None):
        self.name = name
        self.val_shape = tuple(int(s) for s in val_shape)
        v = np.asarray(val, dtype=np.float64)
        if v.shape == self.val_shape:
            self.val = v.copy()
        elif v.size == 1:
            self.val = np.ones(self.val_shape) * float(v.reshape(-1)[0])
        else:
            raise ValueError('val shape %s does not match val_shape %s' % (v.shape, self.val_shape))
        self.fixed = np.zeros(self.val_shape, dtype=bool) if fixed is None else np.asarray(fixed, dtype=bool).reshape(self.val_shape)
        self.prior = SepiaPrior(self, dist=dist, params=params, bounds=bounds)
        self.mcmc = SepiaMCMC(self, stepType=mcmcStepType, stepParam=mcmcStepParam)
        self.orig_range = orig_range
        self.refVal = None

    def get_num_samples(self):
        return len(self.mcmc.draws)

    def set_val(self, val):
        v = np.asarray(val, dtype=np.float64)
        self.val = v.copy() if v.shape == self.val_shape else np.ones(self.val_shape) * float(v.reshape(-1)[0])

    def mcmc_to_array(self, sampleset=None, flat=True):
        """Recorded draws as (n, prod(shape)) in Fortran order (flat) or (n,) + shape."""
        dr = np.array(self.mcmc.draws, dtype=np.float64).reshape((-1,) + self.val_shape)
        if sampleset is not None:
            dr = dr[np.asarray(sampleset, dtype=int)]
        if flat:
            return dr.transpose((0, 2, 1)).reshape(dr.shape[0], -1)      # per draw: order='F'
        return dr

    # sampling-order (Fortran) views used to build the device tables
    def flat_val(self):
        return self.val.reshape(-1, order='F')
