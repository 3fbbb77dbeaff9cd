"""SepiaPrior mirror (constructed by callers at
/root/reference/experiments/synthetic/analysis/include_trunc_error.py:78-79)."""
import numpy as np

_DEFAULT_BOUNDS = {'Normal': [-np.inf, np.inf], 'Gamma': [0.0, np.inf], 'Beta': [0.0, 1.0],
                   'Uniform': [-np.inf, np.inf]}
_DEFAULT_PARAMS = {'Normal': [0.0, 1.0], 'Gamma': [1.0, 1.0], 'Beta': [1.0, 1.0], 'Uniform': [0.0, 1.0]}


class SepiaPrior:
    def __init__(self, parent, dist='Normal', params=False, bounds=False):
        if dist not in _DEFAULT_BOUNDS:
            raise ValueError('Unknown prior distribution %r' % (dist,))
        self.parent = parent
        self.dist = dist
        if params is False or params is None or len(params) == 0:
            params = _DEFAULT_PARAMS[dist]
        if bounds is False or bounds is None or len(bounds) == 0:
            bounds = _DEFAULT_BOUNDS[dist]
        shape = parent.val_shape
        self.params = []
        for prm in params:
            prm = np.asarray(prm, dtype=np.float64)
            if prm.shape != tuple(shape):
                if prm.size != 1:
                    raise ValueError('prior parameter shape %s does not match val_shape %s' % (prm.shape, shape))
                prm = np.ones(shape) * float(prm.reshape(-1)[0])
            self.params.append(prm)
        self.bounds = [float(bounds[0]), float(bounds[1])]

    def is_in_bounds(self, x=None):
        x = self.parent.val if x is None else x
        return bool(np.all(x >= self.bounds[0]) and np.all(x <= self.bounds[1]))

    def compute_log_prior(self):
        """Host evaluation (used for printing / logPost bookkeeping; scalar work only)."""
        x = self.parent.val
        if not self.is_in_bounds(x):
            return -np.inf
        if self.dist == 'Gamma':
            a, b = self.params[0], self.params[1]
            return float(np.sum((a - 1.0) * np.log(x) - b * x))
        if self.dist == 'Beta':
            a, b = self.params[0], self.params[1]
            rho = np.exp(-x / 4.0)
            rho[rho > 0.999] = 0.999
            return float(np.sum((a - 1.0) * np.log(rho) + (b - 1.0) * np.log(1.0 - rho)))
        if self.dist == 'Normal':
            mu, sd = self.params[0], self.params[1]
            return float(-0.5 * np.sum(np.square((x - mu) / sd)))
        return 0.0
