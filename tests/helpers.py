"""Shared problem builders for the tests (oracle side only)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from oracle import sepia_oracle as so          # noqa: E402
from oracle import svd_oracle                  # noqa: E402
from gladsgp_b200 import synthetic             # noqa: E402


def make_problem(m=64, q=3, pu=2, n_x=20, n_t=10, seed=5):
    """Returns dict(t, y, y_std, mu, sd, K, num) for a small multivariate sim-only problem."""
    t = synthetic.design(m, q, seed=seed)
    y = synthetic.ensemble(t, n_x=n_x, n_t=n_t, seed=seed)
    y_std, mu, sd = synthetic.standardize(y)
    rng = np.random.RandomState(seed)
    U, S, Vh = svd_oracle.randomized_svd(y_std, min(25, m), k=0, q=1, rng=rng)
    K = svd_oracle.k_basis(S, Vh, pu, m).astype(np.float32)
    num = so.OracleNum(t, y_std, K)
    return dict(t=t, y=y, y_std=y_std, mu=mu, sd=sd, K=K, num=num)


def make_scalar_problem(m=200, seed=3, noise=0.01):
    """cfg 2: 1-D x, f(x) = x sin(2 pi x) (examples/02_univariate_GP_regression.ipynb cell 3)."""
    rng = np.random.default_rng(seed)
    x = np.sort(rng.uniform(0, 1, m)).astype(np.float32)[:, None]
    f = (x[:, 0] * np.sin(2 * np.pi * x[:, 0]) + noise * rng.standard_normal(m)).astype(np.float32)
    y_std = ((f - f.mean()) / f.std(ddof=1))[:, None]
    num = so.OracleNum(x, y_std, None)
    return dict(t=x, y=f[:, None], y_std=y_std, num=num, mu=f.mean(), sd=f.std(ddof=1))


def random_hypers(num, B, seed=0):
    """B random hyper-parameter sets (beta (B,d), lamz (B,), lamws (B,), lamwos (B,))."""
    rng = np.random.default_rng(seed)
    beta = np.exp(rng.uniform(np.log(0.02), np.log(4.0), size=(B, num.d)))
    lamz = rng.gamma(5.0, 0.2, size=B) + 0.3
    lamws = rng.uniform(200.0, 5000.0, size=B)
    lamwos = rng.uniform(20.0, 300.0, size=B)
    return beta, lamz, lamws, lamwos


def tables_from_oracle(model):
    """Length-P per-element tables (SEPIA sampling order) from an OracleModel."""
    pk = {'Uniform': 0, 'Gamma': 1, 'Beta': 2, 'Normal': 3}
    qk = {'Uniform': 0, 'BetaRho': 1, 'PropMH': 2}
    tb = {k: [] for k in ('prior_kind', 'prior_a', 'prior_b', 'lo', 'hi', 'prop_kind', 'fixed', 'step', 'theta')}
    for p in model.mcmcList:
        n = p.val.size
        tb['prior_kind'] += [pk[p.dist]] * n
        tb['prior_a'] += list(p.params[0].reshape(-1, order='F'))
        tb['prior_b'] += list(p.params[1].reshape(-1, order='F'))
        tb['lo'] += [p.bounds[0]] * n
        tb['hi'] += [p.bounds[1]] * n
        tb['prop_kind'] += [qk[p.step_type]] * n
        tb['fixed'] += list(p.fixed.reshape(-1, order='F').astype(np.uint8))
        tb['step'] += list(p.step.reshape(-1, order='F'))
        tb['theta'] += list(p.val.reshape(-1, order='F'))
    return {k: np.asarray(v) for k, v in tb.items()}


def replay_from_trace(trace, n_steps, P):
    """Oracle per-site trace -> replay tensors (n_steps, 1, P)."""
    cand = np.zeros((n_steps, 1, P)); lac = np.zeros((n_steps, 1, P)); lu = np.zeros((n_steps, 1, P))
    valid = np.zeros((n_steps, 1, P), dtype=np.uint8); acc = np.zeros((n_steps, 1, P), dtype=np.uint8)
    assert len(trace) == n_steps * P
    for i, tr in enumerate(trace):
        t, s = divmod(i, P)
        cand[t, 0, s] = tr['cand']
        valid[t, 0, s] = tr['valid']
        acc[t, 0, s] = tr['accept']
        if tr['valid']:
            lac[t, 0, s] = np.log(tr['aCorr'])
            lu[t, 0, s] = np.log(tr['u2'])
    return dict(cand=cand, logacorr=lac, logu=lu, valid=valid), acc
