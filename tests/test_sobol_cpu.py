"""CPU: the Sobol' / Saltelli oracle (oracle/sobol_oracle.py) against outputs of the reference's own src/utils.py
(tests/golden/sobol_reference.npz, produced by tests/golden/make_golden_sobol.py)."""
import os
import sys
import warnings

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'tests', 'golden'))
from make_golden_sobol_function import analytic_function  # noqa: E402
from oracle import sobol_oracle as sob  # noqa: E402

G = np.load(os.path.join(ROOT, 'tests', 'golden', 'sobol_reference.npz'))


def test_oracle_pca_matches_reference():
    np.random.seed(int(G['np_seed']))
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        first, total, gf, gt, res = sob.PCA_saltelli_sensitivity_indices(analytic_function, int(G['n_dim']), int(G['m']), G['pcvar'],
                                                                         bootstrap=True, AB=G['pca_AB'])
    np.testing.assert_allclose(first, G['pca_first'], rtol=1e-13, atol=1e-15)
    np.testing.assert_allclose(total, G['pca_total'], rtol=1e-13, atol=1e-15)
    np.testing.assert_allclose(gf, G['pca_gen_first'], rtol=1e-13)
    np.testing.assert_allclose(gt, G['pca_gen_total'], rtol=1e-13)
    for k in ('first_order', 'total_index', 'general_first_order', 'general_total_index'):
        ci = np.array((res[k].confidence_interval.low, res[k].confidence_interval.high))
        np.testing.assert_allclose(ci, G['pca_ci_' + k], rtol=1e-12, atol=1e-14)
        np.testing.assert_allclose(res[k].standard_error, G['pca_se_' + k], rtol=1e-12)


def test_oracle_scalar_matches_reference():
    np.random.seed(int(G['np_seed']))
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        first, total, res = sob.saltelli_sensitivity_indices(analytic_function, int(G['n_dim']), int(G['m']), bootstrap=True,
                                                             AB=G['scalar_AB'])
    np.testing.assert_allclose(first, G['scalar_first'], rtol=1e-13, atol=1e-15)
    np.testing.assert_allclose(total, G['scalar_total'], rtol=1e-13, atol=1e-15)
    for k in ('first_order', 'total_index'):
        ci = np.array((res[k].confidence_interval.low, res[k].confidence_interval.high))
        np.testing.assert_allclose(ci, G['scalar_ci_' + k], rtol=1e-12, atol=1e-14)


def test_sobol_matrix_is_the_references_split():
    AB = sob.sobol_matrix(3, 4, seed=1)
    assert AB.shape == (16, 6) and AB.min() >= 0.0 and AB.max() < 1.0
