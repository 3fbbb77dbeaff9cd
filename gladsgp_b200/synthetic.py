"""Seeded synthetic "GlaDS-shaped" ensembles.

The real 2.8 GB ensemble is not in the reference repository, so benchmarks and tests use
smooth space-time fields with the reference's PCA spectrum
(experiments/synthetic/analysis/data/architecture/pca_cvar_n512.csv:1-10).
"""
import numpy as np

BASE_SEED = 20240318   # as experiments/synthetic/train_config.py:41
_CVAR = np.array([0.809, 0.867, 0.901, 0.918, 0.932, 0.940, 0.948, 0.954, 0.957, 0.960])


def design(m, q, seed=BASE_SEED):
    """Scrambled Sobol' design in [0,1]^q cast to float32 (train_config.py:41)."""
    from scipy.stats import qmc
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        t = qmc.Sobol(q, seed=seed).random(m)
    return t.astype(np.float32)


def test_design(n, q, seed=42186):
    """LHS test designs (assess_all_models.py:446-447)."""
    from scipy.stats import qmc
    return qmc.LatinHypercube(q, seed=seed).random(n).astype(np.float32)


def mode_weights(t, r=32, seed=BASE_SEED):
    """Smooth functions a_k(t) of the design: random-feature GP draws, lengthscale 0.3-1."""
    rng = np.random.default_rng(seed + 7)
    q = t.shape[1]
    nf = 64
    out = np.zeros((t.shape[0], r))
    for k in range(r):
        ell = rng.uniform(0.3, 1.0)
        Wf = rng.normal(size=(q, nf)) / ell
        ph = rng.uniform(0, 2 * np.pi, nf)
        cf = rng.normal(size=nf) * np.sqrt(2.0 / nf)
        out[:, k] = np.cos(t.astype(np.float64) @ Wf + ph) @ cf
    return out


def spectrum(r=32):
    var = np.diff(np.concatenate([[0.0], _CVAR]))
    tail = (1.0 - _CVAR[-1]) * 0.5 ** np.arange(1, r - len(var) + 1)
    v = np.concatenate([var, tail])[:r]
    return np.sqrt(v / v.sum())


def ensemble(t, n_x=64, n_t=16, r=32, noise=0.05, seed=BASE_SEED, n_const=3, dtype=np.float32):
    """(m, n_x*n_t) float32 flotation-fraction-like fields: mean + sum_k a_k(t) phi_k sigma_k + eps."""
    rng = np.random.default_rng(seed + 11)
    m = t.shape[0]
    a = mode_weights(t, r, seed)
    sig = spectrum(r)
    s = (np.arange(n_x) + 0.5) / n_x
    tau = (np.arange(n_t) + 0.5) / n_t
    n_y = n_x * n_t
    y = np.empty((m, n_y), dtype=dtype)
    mu = (1.0 + 0.5 * np.cos(np.pi * s)[:, None] * (0.6 + 0.4 * np.sin(2 * np.pi * tau)[None, :])).reshape(-1)
    phi = np.empty((r, n_y))
    for k in range(r):
        fs, ft = rng.integers(0, 4), rng.integers(0, 3)
        ps, pt = rng.uniform(0, np.pi), rng.uniform(0, np.pi)
        phi[k] = (np.cos(np.pi * fs * s + ps)[:, None] * np.cos(2 * np.pi * ft * tau + pt)[None, :]).reshape(-1)
        phi[k] /= max(np.sqrt(np.mean(phi[k] ** 2)), 1e-12)
    coef = a * sig[None, :] * 0.35
    blk = 1 << 16
    for c0 in range(0, n_y, blk):
        c1 = min(n_y, c0 + blk)
        y[:, c0:c1] = (mu[c0:c1][None, :] + coef @ phi[:, c0:c1] +
                       noise * rng.standard_normal((m, c1 - c0))).astype(dtype)
    if n_const:
        y[:, :n_const] = y[0, :n_const]          # constant columns -> sd clamp (src/model.py:64)
    return y


def posterior_samples(nsamp, d, pu, seed=BASE_SEED, f32=True):
    """Synthetic posterior draws for prediction-only benches, rounded to f32 as the
    callers do (assess_all_models.py:473-474)."""
    rng = np.random.default_rng(seed + 23)
    s = dict(
        betaU=np.exp(rng.uniform(np.log(1e-2), np.log(5.0), size=(nsamp, d * pu))),
        lamUz=rng.gamma(5.0, 1.0 / 5.0, size=(nsamp, pu)) + 0.3,
        lamWs=rng.uniform(500.0, 5000.0, size=(nsamp, pu)),
        lamWOs=rng.uniform(10.0, 200.0, size=(nsamp, 1)),
    )
    if f32:
        s = {k: v.astype(np.float32) for k, v in s.items()}
    return s


def standardize(y, sd_threshold=1e-6):
    """src/model.py:60-64,72."""
    mu = np.mean(y, axis=0)
    sd = np.std(y, ddof=1, axis=0)
    sd[sd < sd_threshold] = sd_threshold
    return (y - mu) / sd, mu, sd
