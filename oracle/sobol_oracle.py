"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference's Sobol' / Saltelli sensitivity path.

Follows /root/reference/src/utils.py:27-125 (`saltelli_sensitivity_indices`, scalar outputs) and :128-256
(`PCA_saltelli_sensitivity_indices`, PC weights + explained-variance weighted "general" indices), callers
/root/reference/experiments/synthetic/analysis/sensitivity_indices.py:96,214.  Same arithmetic, same library calls
(`scipy.stats.bootstrap`, default BCa method, one call per statistic in the reference's order); the only change is that the
three sources of randomness can be injected: the 2*n_dim-column Sobol' matrix `AB` (the reference draws an unseeded
scrambled `scipy.stats.qmc.Sobol`), the number of resamples (the reference hard-codes 9999) and the random state of the
bootstrap (the reference uses the global `np.random` state, which is also the default here).

Parity: PINNED.  tests/golden/make_golden_sobol.py imports the reference's own `src/utils.py` (it needs only numpy/scipy),
runs both functions on an analytic test function with the Sobol' sampler seeded and `np.random.seed` set, and commits inputs
and outputs (tests/golden/sobol_reference.npz); tests/test_oracle_cpu.py checks this file against them.

Only tests/, __graft_entry__.smoke() and bench.py's CPU baseline may import this module.
"""
import numpy as np
from scipy import stats


def sobol_matrix(n_dim, m, seed=None):
    """AB (2**m, 2*n_dim) as src/utils.py:66-67; A = AB[:, n_dim:], B = AB[:, :n_dim] (:68-69)."""
    sampler = stats.qmc.Sobol(d=2 * n_dim, seed=seed)
    return sampler.random_base2(m=int(m))


def evaluate_blocks(func, AB, n_dim):
    """f_A, f_B, f_AB[ix] in the reference's call order (src/utils.py:70-80)."""
    A = AB[:, n_dim:]
    B = AB[:, :n_dim]
    f_A = func(A)
    f_B = func(B)
    f_AB = np.zeros((n_dim, f_A.shape[0], f_A.shape[1]))
    for ix in range(n_dim):
        C = B.copy()
        C[:, ix] = A[:, ix]
        f_AB[ix] = func(C)
    return f_A, f_B, f_AB


def point_estimates(f_A, f_B, f_AB):
    """first_order, total_index (p, n_dim): src/utils.py:71-92."""
    n_dim = f_AB.shape[0]
    first_order = np.zeros((f_A.shape[1], n_dim))
    total_index = np.zeros((f_A.shape[1], n_dim))
    var = np.var([f_A, f_B], axis=(0, 1))
    for ix in range(n_dim):
        f_C = f_AB[ix]
        V_i = np.mean(f_A * (f_C - f_B), axis=0)
        E_i = 0.5 * np.mean((f_B - f_C) ** 2, axis=0)
        first_order[:, ix] = V_i / var
        total_index[:, ix] = E_i / var
    return first_order, total_index


def statistics(f_A, f_B, f_AB, pcvar=None):
    """The closures handed to scipy.stats.bootstrap (src/utils.py:97-118, :213-243)."""
    def first_order_statistic(arg):
        f_A_ = f_A[arg, :]
        f_B_ = f_B[arg, :]
        f_AB_ = f_AB[:, arg, :]
        V_ix = np.mean(f_A_ * (f_AB_ - f_B_), axis=(1,))
        V_ix[V_ix < 0] = 0
        var = np.var([f_A_, f_B_], axis=(0, 1))
        return (V_ix / var).T

    def total_index_statistic(arg):
        f_A_ = f_A[arg, :]
        f_B_ = f_B[arg, :]
        f_AB_ = f_AB[:, arg, :]
        E_ix = 0.5 * np.mean((f_B_ - f_AB_) ** 2, axis=(1,))
        E_ix[E_ix < 0] = 0
        var = np.var([f_A_, f_B_], axis=(0, 1))
        return (E_ix / var).T

    out = {'first_order': first_order_statistic, 'total_index': total_index_statistic}
    if pcvar is not None:
        out['general_first_order'] = lambda arg: np.sum(first_order_statistic(arg) * np.vstack(pcvar), axis=0)
        out['general_total_index'] = lambda arg: np.sum(total_index_statistic(arg) * np.vstack(pcvar), axis=0)
    return out


def saltelli_sensitivity_indices(func, n_dim, m, bootstrap=True, AB=None, n_resamples=9999, rng=None):
    """src/utils.py:27-125."""
    AB = sobol_matrix(n_dim, m) if AB is None else np.asarray(AB)
    f_A, f_B, f_AB = evaluate_blocks(func, AB, n_dim)
    first_order, total_index = point_estimates(f_A, f_B, f_AB)
    res = None
    if bootstrap:
        st = statistics(f_A, f_B, f_AB)
        N = f_A.shape[0]
        kw = {} if rng is None else {'rng': rng}      # no keyword at all = the reference's call: global np.random state
        res = {k: stats.bootstrap([np.arange(N)], st[k], n_resamples=n_resamples, **kw)
               for k in ('first_order', 'total_index')}
    return first_order, total_index, res


def PCA_saltelli_sensitivity_indices(func, n_dim, m, pcvar, bootstrap=True, AB=None, n_resamples=9999, rng=None):
    """src/utils.py:128-256."""
    AB = sobol_matrix(n_dim, m) if AB is None else np.asarray(AB)
    f_A, f_B, f_AB = evaluate_blocks(func, AB, n_dim)
    first_order, total_index = point_estimates(f_A, f_B, f_AB)
    res = None
    if bootstrap:
        st = statistics(f_A, f_B, f_AB, pcvar=pcvar)
        N = f_A.shape[0]
        kw = {} if rng is None else {'rng': rng}
        res = {k: stats.bootstrap([np.arange(N)], st[k], n_resamples=n_resamples, **kw)
               for k in ('first_order', 'total_index', 'general_first_order', 'general_total_index')}
    gen_first = np.sum(first_order.T * pcvar, axis=1)
    gen_total = np.sum(total_index.T * pcvar, axis=1)
    return first_order, total_index, gen_first, gen_total, res
