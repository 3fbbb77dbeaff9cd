"""SepiaMCMC mirror: proposal bookkeeping + recorded draws of one parameter."""
import numpy as np


class SepiaMCMC:
    def __init__(self, parent, stepType='Normal', stepParam=0.1):
        if stepType not in ('Normal', 'Uniform', 'BetaRho', 'PropMH', 'Recorder'):
            raise ValueError('Unknown mcmcStepType %r' % (stepType,))
        self.parent = parent
        self.stepType = stepType
        sp = np.asarray(stepParam, dtype=np.float64)
        self.stepParam = sp.copy() if sp.shape == tuple(parent.val_shape) else np.ones(parent.val_shape) * float(sp.reshape(-1)[0])
        self.draws = []
        self.aCorr = 1

    def record(self):
        self.draws.append(self.parent.val.copy())

    def draw_candidate(self, arr_ind, do_propMH):
        """Host restatement of the proposal (one np.random draw); the device sampler consumes the same
        stream the same way (csrc/ggp_mcmc.cu plan_kernel)."""
        self.aCorr = 1
        x = self.parent.val[arr_ind]
        st = self.stepParam[arr_ind]
        if self.stepType == 'Uniform' or (self.stepType == 'PropMH' and not do_propMH):
            return x + st * np.random.uniform(-0.5, 0.5)
        if self.stepType == 'BetaRho':
            cand = np.exp(-x / 4.0) + st * np.random.uniform(-0.5, 0.5)
            return np.inf if cand <= 0 else -4.0 * np.log(cand)
        if self.stepType == 'PropMH':
            w = max(1.0, x / 3.0)
            dval = x + w * np.random.uniform(-1.0, 1.0)
            w1 = max(1.0, dval / 3.0)
            self.aCorr = False if x > dval + w1 else w / w1
            return dval
        if self.stepType == 'Normal':
            return x + st * np.random.normal()
        raise ValueError(self.stepType)
