"""Golden vectors for the Sobol' / Saltelli sensitivity path, produced by the REFERENCE's own code.

Run in the build container (needs /root/reference):  python tests/golden/make_golden_sobol.py
Imports /root/reference/src/utils.py unmodified; the only interventions are on its sources of randomness: the
`scipy.stats.qmc.Sobol` constructor it calls without a seed is wrapped to pass one, and `np.random.seed` pins the global
state that scipy.stats.bootstrap falls back to.  Output: tests/golden/sobol_reference.npz (inputs captured from the calls of
the test function, point estimates, bootstrap confidence limits).
"""
import contextlib
import io
import os
import sys

import numpy as np
from scipy import stats

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, '/root/reference')
from src import utils as ref_utils  # noqa: E402
sys.path.insert(0, HERE)
from make_golden_sobol_function import analytic_function  # noqa: E402

SOBOL_SEED = 20240318
NP_SEED = 4242


def main():
    n_dim, m = 3, 5
    pcvar = np.array([0.6, 0.25, 0.1, 0.05])
    real_sobol = stats.qmc.Sobol

    def seeded_sobol(d, **kw):
        kw.setdefault('seed', SOBOL_SEED)
        return real_sobol(d=d, **kw)

    out = {'n_dim': n_dim, 'm': m, 'pcvar': pcvar, 'sobol_seed': SOBOL_SEED, 'np_seed': NP_SEED}
    for name in ('pca', 'scalar'):
        calls = []

        def func(x):
            calls.append(np.array(x))
            return analytic_function(x)
        stats.qmc.Sobol = seeded_sobol
        np.random.seed(NP_SEED)
        try:
            with contextlib.redirect_stdout(io.StringIO()):
                if name == 'pca':
                    r = ref_utils.PCA_saltelli_sensitivity_indices(func, n_dim, m, pcvar, bootstrap=True)
                    first, total, gfirst, gtotal, res = r
                    out['pca_gen_first'] = gfirst
                    out['pca_gen_total'] = gtotal
                else:
                    first, total, res = ref_utils.saltelli_sensitivity_indices(func, n_dim, m, bootstrap=True)
        finally:
            stats.qmc.Sobol = real_sobol
        A, B = calls[0], calls[1]
        out[name + '_AB'] = np.concatenate([B, A], axis=1)          # AB[:, :n_dim] = B, AB[:, n_dim:] = A (src/utils.py:68-69)
        out[name + '_first'] = first
        out[name + '_total'] = total
        for k, v in res.items():
            out['%s_ci_%s' % (name, k)] = np.array((v.confidence_interval.low, v.confidence_interval.high))
            out['%s_se_%s' % (name, k)] = np.asarray(v.standard_error)
    np.savez_compressed(os.path.join(HERE, 'sobol_reference.npz'), **out)
    print('wrote sobol_reference.npz:', {k: np.shape(v) for k, v in out.items()})


if __name__ == '__main__':
    main()
