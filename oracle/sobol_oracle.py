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
    """Function values on the two base designs and on the n_dim "radial" designs (B with column i taken from A), in the
    reference's call order A, B, AB_0, ... (src/utils.py:68-80).  Returns f_A, f_B (N, p) and f_AB (n_dim, N, p)."""
    base_a, base_b = AB[:, n_dim:], AB[:, :n_dim]
    f_A = func(base_a)
    f_B = func(base_b)
    mixed = []
    for i in range(n_dim):
        swap = np.arange(n_dim) == i
        mixed.append(func(np.where(swap[None, :], base_a, base_b)))
    return f_A, f_B, np.stack(mixed).astype(np.float64)


def _saltelli_ratios(fa, fb, fab, clamp):
    """Saltelli (2010) estimators on already gathered rows: fa, fb (n, p), fab (n_dim, n, p) -> first, total (p, n_dim).
    Numerators: mean f_A (f_AB - f_B) and half the mean of (f_B - f_AB)^2; denominator: population variance of the 2n pooled
    values of f_A and f_B (src/utils.py:81-92; clamped at zero inside the bootstrap statistics, :101, :113)."""
    n = fa.shape[0]
    num_first = (fa[None, :, :] * (fab - fb[None, :, :])).sum(axis=1) / n            # (n_dim, p)
    num_total = ((fb[None, :, :] - fab) ** 2).sum(axis=1) / n * 0.5
    if clamp:
        num_first = np.maximum(num_first, 0.0)
        num_total = np.maximum(num_total, 0.0)
    pooled_var = np.var(np.stack([fa, fb]), axis=(0, 1))                               # (p,)
    return (num_first / pooled_var).T, (num_total / pooled_var).T


def point_estimates(f_A, f_B, f_AB):
    """first_order, total_index (p, n_dim) on the full sample, not clamped (src/utils.py:71-92)."""
    return _saltelli_ratios(f_A, f_B, f_AB, clamp=False)


def statistics(f_A, f_B, f_AB, pcvar=None):
    """The index-set statistics scipy.stats.bootstrap is given (src/utils.py:97-118; with the explained-variance weighted sums of
    :237-243 when pcvar is given): callables of one argument, the array of row indices of a resample."""
    def first(rows):
        return _saltelli_ratios(f_A[rows], f_B[rows], f_AB[:, rows], clamp=True)[0]

    def total(rows):
        return _saltelli_ratios(f_A[rows], f_B[rows], f_AB[:, rows], clamp=True)[1]

    out = {'first_order': first, 'total_index': total}
    if pcvar is not None:
        wts = np.asarray(pcvar, dtype=np.float64)[:, None]
        out['general_first_order'] = lambda rows: (first(rows) * wts).sum(axis=0)
        out['general_total_index'] = lambda rows: (total(rows) * wts).sum(axis=0)
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
