"""CPU oracle for the GladsGP / SEPIA emulator hot path.  TEST INFRASTRUCTURE ONLY.

This module is a NumPy/SciPy FP64 restatement of the arithmetic that the reference
(timghill/GladsGP) drives through its un-vendored dependency
``sepia @ git+https://github.com/timghill/SEPIA.git@ffe3b60c4b864c32fb15c26003d3a77e95910093``
(pinned at /root/reference/requirements-cc.txt:42).  The package source is absent from
/root/reference and cannot be installed here, so this file restates the published SEPIA
algorithm and anchors on the reference's own call sites:

* src/model.py:56-106   SepiaData / transform_xt / standardize_y / create_K_basis / SepiaModel
* src/model.py:218-235  w = pinv(K)^T y_std^T, lamWOs override, tune_step_sizes, do_mcmc
* experiments/synthetic/analysis/assess_all_models.py:468-500  get_samples, prediction, get_y
* examples/04_GP_emulation_multivariate_ensemble.ipynb:255-316  recorded shapes / step sizes

PARITY STATUS: **parity unpinned** for the SEPIA part -- the reference tree holds no golden
vector, known-answer test or fixture for any log-likelihood, chain or prediction (SURVEY.md
section 8c).  What *is* pinned: default step sizes / shapes recorded in the notebooks
(tests/test_oracle_cpu.py) and, for the PCA, the reference's own src/svd.py (see
oracle/svd_oracle.py, validated against an import of /root/reference/src/svd.py).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module.  The product path (gladsgp_b200/, sepia/) never does.

The structure deliberately mirrors SEPIA's: pair-list covariance build + scatter, one
scipy.linalg.cholesky + solve_triangular per block evaluation, single-site sequential sweep,
and, for prediction, a fresh S22 solve per (sample, PC, call).
"""
from __future__ import annotations

import copy
import numpy as np
import scipy.linalg


# --------------------------------------------------------------------------------------------
# set-up (call sites src/model.py:56-106, :219)
# --------------------------------------------------------------------------------------------
class OracleNum:
    """Numeric set-up of a sim-only SepiaModel (SepiaModel.__init__ upstream)."""

    def __init__(self, t_trans, y_std, K=None, resid_ss=None):
        t_trans = np.asarray(t_trans)
        y_std = np.asarray(y_std)
        if y_std.ndim == 1:
            y_std = y_std[:, None]
        m = t_trans.shape[0]
        self.m = m
        self.p = 1
        self.q = t_trans.shape[1]
        self.d = self.p + self.q
        self.scalar_out = (y_std.shape[1] == 1) or K is None
        x = 0.5 * np.ones((m, 1))                      # dummy x (float64 -> zt float64)
        self.zt = np.concatenate([x, t_trans], axis=1)
        self.iu = np.triu_indices(m, 1)
        self.sqdist = np.square(self.zt[self.iu[0]] - self.zt[self.iu[1]])
        if self.scalar_out:
            self.pu = 1
            self.w = y_std.astype(np.float64).reshape(m, 1)
            self.LamSim = np.ones(1)
            self.K = None
        else:
            self.K = np.asarray(K)
            self.pu = self.K.shape[0]
            # src/model.py:219  w = (pinv(K).T @ y_std.T).T
            self.w = np.dot(np.linalg.pinv(self.K).T, y_std.T).T.astype(np.float64)
            self.LamSim = np.diag(np.dot(self.K, self.K.T)).astype(np.float64)
        self.wv = self.w.reshape((-1, 1), order='F')   # PC-major stack
        self.n_y = y_std.shape[1]
        if self.scalar_out:
            self.resid_ss = 0.0
        elif resid_ss is not None:
            self.resid_ss = float(resid_ss)     # caller-supplied (skips the m x n_y residual pass)
        else:
            r = y_std - np.dot(self.w, self.K)
            self.resid_ss = float(np.sum(np.square(r, dtype=np.float64)))


# --------------------------------------------------------------------------------------------
# covariance (SepiaDistCov upstream)
# --------------------------------------------------------------------------------------------
def cov_self(num, beta, lamz):
    """SepiaDistCov type 1: pair list -> exp -> scatter -> symmetrise -> diagonal."""
    m = num.m
    C = np.zeros((m, m))
    C[num.iu] = np.exp(-(num.sqdist @ beta)) / lamz
    C = C + C.T
    np.fill_diagonal(C, 1.0 / lamz)
    return C


def cov_cross(A, B, beta, lamz):
    """SepiaDistCov type 2: (len(A), len(B)) cross covariance."""
    D = np.square(A[:, None, :] - B[None, :, :])
    return np.exp(-(D @ beta)) / lamz


def do_loglik(C, wj):
    """doLogLik upstream: -sum log diag chol(C) - 1/2 ||L^-1 w||^2 ; -inf on failure."""
    try:
        L = scipy.linalg.cholesky(C, lower=True)
    except (np.linalg.LinAlgError, scipy.linalg.LinAlgError, ValueError):
        return -np.inf
    u = scipy.linalg.solve_triangular(L, wj, lower=True)
    return -np.sum(np.log(np.diag(L))) - 0.5 * np.sum(u * u)


def block_cov(num, beta_j, lamUz_j, lamWs_j, lamWOs, j):
    """C_j of compute_log_lik: cov_self + nuggets on the diagonal."""
    C = cov_self(num, beta_j, lamUz_j)
    np.fill_diagonal(C, C.diagonal() + 1.0 / (num.LamSim[j] * lamWOs) + 1.0 / lamWs_j)
    return C


# --------------------------------------------------------------------------------------------
# parameters, priors, proposals (SepiaParam / SepiaPrior / SepiaMCMC upstream)
# --------------------------------------------------------------------------------------------
class OracleParam:
    def __init__(self, val, name, val_shape, dist, params, bounds, step, step_type):
        self.name = name
        self.val_shape = tuple(val_shape)
        self.val = np.ones(self.val_shape) * np.asarray(val, dtype=np.float64)
        self.dist = dist
        self.params = [np.ones(self.val_shape) * float(p) for p in params]
        self.bounds = [float(bounds[0]), float(bounds[1])]
        self.step = np.ones(self.val_shape) * float(step)
        self.step_type = step_type
        self.fixed = np.zeros(self.val_shape, dtype=bool)
        self.draws = []
        self.aCorr = 1.0

    def in_bounds(self, x=None):
        x = self.val if x is None else x
        return bool(np.all(x >= self.bounds[0]) and np.all(x <= self.bounds[1]))

    def log_prior(self):
        x = self.val
        if not self.in_bounds():
            return -np.inf
        if self.dist == 'Gamma':
            a, b = self.params
            return float(np.sum((a - 1.0) * np.log(x) - b * x))
        if self.dist == 'Beta':
            a, b = self.params
            rho = np.exp(-x / 4.0)
            rho[rho > 0.999] = 0.999
            return float(np.sum((a - 1.0) * np.log(rho) + (b - 1.0) * np.log(1.0 - rho)))
        if self.dist == 'Normal':
            mu, sd = self.params
            return float(-0.5 * np.sum(np.square((x - mu) / sd)))
        if self.dist == 'Uniform':
            return 0.0
        raise ValueError(self.dist)

    def draw_candidate(self, ai, do_propMH, rng):
        """One RNG draw, always.  Returns (cand, aCorr, raw_uniform)."""
        x = self.val[ai]
        st = self.step[ai]
        u = rng.random_sample()
        aCorr = 1.0
        if self.step_type == 'Uniform' or (self.step_type == 'PropMH' and not do_propMH):
            cand = x + st * (-0.5 + 1.0 * u)          # np.random.uniform(-0.5, 0.5)
        elif self.step_type == 'BetaRho':
            rho = np.exp(-x / 4.0) + st * (-0.5 + 1.0 * u)
            cand = np.inf if rho <= 0 else -4.0 * np.log(rho)
        elif self.step_type == 'PropMH':
            w = max(1.0, x / 3.0)
            cand = x + w * (-1.0 + 2.0 * u)           # np.random.uniform(-1, 1)
            w1 = max(1.0, cand / 3.0)
            aCorr = 0.0 if x > cand + w1 else w / w1
        elif self.step_type == 'Normal':
            raise NotImplementedError('Normal proposals are never used on the reference path')
        else:
            raise ValueError(self.step_type)
        return cand, aCorr, u


class OracleModel:
    """Sim-only SepiaModel restatement."""

    def __init__(self, num: OracleNum):
        self.num = num
        d, pu, m = num.d, num.pu, num.m
        self.betaU = OracleParam(0.1, 'betaU', (d, pu), 'Beta', [1.0, 0.1], [0.0, np.inf], 0.1, 'BetaRho')
        self.lamUz = OracleParam(1.0, 'lamUz', (1, pu), 'Gamma', [5.0, 5.0], [0.3, np.inf], 5.0, 'PropMH')
        self.lamWs = OracleParam(1000.0, 'lamWs', (1, pu), 'Gamma', [3.0, 3e-3], [60.0, 1e5], 100.0, 'PropMH')
        if num.scalar_out:
            a, b = 5.0, 5e-3
        else:
            a = 5.0 + 0.5 * m * (num.n_y - pu)
            b = 5e-3 + 0.5 * num.resid_ss
        self.lamWOs = OracleParam(max(100.0, a / b), 'lamWOs', (1, 1), 'Gamma', [a, b], [60.0, 1e5], 100.0, 'PropMH')
        self.mcmcList = [self.betaU, self.lamUz, self.lamWs, self.lamWOs]
        self.SigWl = np.zeros(pu)
        self.lp_draws = []
        self.trace = None          # optional per-site recording (replay tensors)

    # reference override, src/model.py:225-231
    def override_lamWOs(self, pc_prec, gamma_a=50.0):
        self.lamWOs = OracleParam(pc_prec, 'lamWOs', (1, 1), 'Gamma', [gamma_a, gamma_a / pc_prec],
                                  [1.0, np.inf], 10.0, 'Uniform')
        self.mcmcList = [self.betaU, self.lamUz, self.lamWs, self.lamWOs]

    # -- likelihood ---------------------------------------------------------------------------
    def log_lik(self, cvar='all', cindex=None):
        num = self.num
        d, pu, m = num.d, num.pu, num.m
        if cvar == 'betaU':
            J = [cindex // d]
        elif cvar in ('lamUz', 'lamWs'):
            J = [cindex]
        else:
            J = range(pu)
        for j in J:
            C = block_cov(num, self.betaU.val[:, j], self.lamUz.val[0, j], self.lamWs.val[0, j],
                          self.lamWOs.val[0, 0], j)
            self.SigWl[j] = do_loglik(C, num.wv[j * m:(j + 1) * m, 0])
        return float(np.sum(self.SigWl))

    def log_post(self, cvar='all', cindex=None):
        ll = self.log_lik(cvar, cindex)
        return ll + sum(p.log_prior() for p in self.mcmcList)

    # -- one Metropolis-within-Gibbs step -------------------------------------------------------
    def mcmc_step(self, do_propMH=True, rng=np.random):
        lp = self.log_post()
        for prm in self.mcmcList:
            for ind in range(prm.val.size):
                ai = np.unravel_index(ind, prm.val.shape, order='F')
                ref_val, ref_SigWl = prm.val.copy(), self.SigWl.copy()
                cand, aCorr, u1 = prm.draw_candidate(ai, do_propMH, rng)     # RNG draw #1 (always)
                prm.val[ai] = cand
                accept = False
                valid = False
                u2 = np.nan
                clp = np.nan
                if not prm.fixed[ai]:
                    if aCorr and prm.in_bounds(prm.val[ai]):
                        valid = True
                        clp = self.log_post(prm.name, ind)
                        u2 = rng.random_sample()                             # RNG draw #2 (conditional)
                        if np.log(u2) < clp - lp + np.log(aCorr):
                            accept = True
                if self.trace is not None:
                    self.trace.append(dict(name=prm.name, ind=ind, cand=cand, aCorr=aCorr, u1=u1, u2=u2,
                                           valid=valid, accept=accept, clp=clp, lp=lp))
                if accept:
                    lp = clp
                else:
                    prm.val = ref_val
                    self.SigWl[:] = ref_SigWl
            prm.draws.append(prm.val.copy())
        self.lp_draws.append(lp)
        return lp

    def do_mcmc(self, nsamp, do_propMH=True, rng=np.random):
        for _ in range(nsamp):
            self.mcmc_step(do_propMH, rng)

    # -- samples (see DESIGN.md for the frozen linspace convention) ----------------
    def get_samples(self, numsamples=None, nburn=0):
        total = len(self.lp_draws)
        ss = np.arange(nburn, total)
        if numsamples is not None and numsamples < total:
            ss = np.array([int(i) for i in np.linspace(nburn, total - 1, numsamples)])
        out = {}
        for prm in self.mcmcList:
            dr = np.array(prm.draws)[ss]
            out[prm.name] = dr.reshape((dr.shape[0], -1), order='F') if dr.ndim == 2 else \
                np.stack([x.reshape(-1, order='F') for x in dr])
        out['logPost'] = np.array(self.lp_draws)[ss].reshape((-1, 1))
        return out

    # -- step size tuning (YADAS-style) ----------------------------------------------
    def tune_step_sizes(self, n_burn, n_levels, rng=np.random, warmup=10):
        mod = copy.deepcopy(self)
        for prm in mod.mcmcList:
            prm.draws = []
        mod.lp_draws = []
        ex = np.linspace(-(n_levels - 1) / 2.0, (n_levels - 1) / 2.0, n_levels)
        steps = {p.name: [p.step * np.power(2.0, e) for e in ex] for p in mod.mcmcList}
        acc = {p.name: np.zeros((n_levels,) + p.val_shape) for p in mod.mcmcList}
        for _ in range(warmup):
            mod.mcmc_step(do_propMH=False, rng=rng)
        for _ in range(n_burn):
            for lev in range(n_levels):
                for p in mod.mcmcList:
                    p.step = steps[p.name][lev].copy()
                before = {p.name: p.val.copy() for p in mod.mcmcList}
                mod.mcmc_step(do_propMH=False, rng=rng)
                for p in mod.mcmcList:
                    acc[p.name][lev] += (p.val != before[p.name])
        target = np.log(1.0 / (np.exp(1.0) - 1.0))
        for p, q in zip(self.mcmcList, mod.mcmcList):
            new = p.step.copy()
            for ind in range(p.val.size):
                ai = np.unravel_index(ind, p.val_shape, order='F')
                if p.fixed[ai]:
                    continue
                x = np.log(np.array([s[ai] for s in steps[p.name]]))
                k = np.array([acc[p.name][lev][ai] for lev in range(n_levels)])
                b = logit_glm(x, k, n_burn)
                if b is not None and b[1] < 0 and np.all(np.isfinite(b)):
                    new[ai] = np.exp((target - b[0]) / b[1])
            p.step = new
            p.val = q.val.copy()
        self.SigWl = mod.SigWl.copy()
        return acc


def logit_glm(x, k, n, iters=100, tol=1e-10):
    """Binomial-logit IRLS of k successes out of n on [1, x] (stand-in for statsmodels GLM)."""
    X = np.stack([np.ones_like(x), x], axis=1)
    y = k / float(n)
    b = np.zeros(2)
    for _ in range(iters):
        eta = np.clip(X @ b, -30, 30)
        mu = 1.0 / (1.0 + np.exp(-eta))
        W = n * mu * (1.0 - mu)
        z = eta + (y - mu) / np.maximum(mu * (1.0 - mu), 1e-12)
        A = X.T @ (W[:, None] * X)
        try:
            bn = np.linalg.solve(A, X.T @ (W * z))
        except np.linalg.LinAlgError:
            return None
        if np.max(np.abs(bn - b)) < tol:
            b = bn
            break
        b = bn
    return b


# --------------------------------------------------------------------------------------------
# prediction (SepiaPredict.wPred / get_y upstream)
# --------------------------------------------------------------------------------------------
def w_pred(num: OracleNum, t_pred, samples, rng=np.random, draw=True):
    """Returns (w_draws (nsamp,npred,pu) or None, mu (nsamp, npred*pu), Sigma (nsamp, npred*pu, npred*pu))."""
    t_pred = np.asarray(t_pred, dtype=np.float64)
    npred = t_pred.shape[0]
    d, pu, m = num.d, num.pu, num.m
    xp = np.concatenate([0.5 * np.ones((npred, 1)), t_pred], axis=1)
    nsamp = samples['lamWs'].shape[0]
    mus = np.zeros((nsamp, npred * pu))
    Sigs = np.zeros((nsamp, npred * pu, npred * pu))
    w_out = np.zeros((nsamp, npred, pu)) if draw else None
    for s in range(nsamp):
        bU = np.asarray(samples['betaU'][s], dtype=np.float64).reshape((d, pu), order='F')
        lUz = np.asarray(samples['lamUz'][s], dtype=np.float64)
        lWs = np.asarray(samples['lamWs'][s], dtype=np.float64)
        lWOs = float(np.asarray(samples['lamWOs'][s]).reshape(-1)[0])
        mu = mus[s]
        Sig = Sigs[s]
        for j in range(pu):
            S22 = block_cov(num, bU[:, j], lUz[j], lWs[j], lWOs, j)
            S21 = cov_cross(num.zt, xp, bU[:, j], lUz[j])
            S11 = cov_cross(xp, xp, bU[:, j], lUz[j])
            np.fill_diagonal(S11, 1.0 / lUz[j] + 1.0 / lWs[j])
            W = np.linalg.solve(S22, S21)
            sl = slice(j * npred, (j + 1) * npred)
            mu[sl] = W.T @ num.wv[j * m:(j + 1) * m, 0]
            Sig[sl, sl] = S11 - S21.T @ W
        if draw:
            U, sv, _ = np.linalg.svd(Sig)
            z = rng.normal(size=npred * pu)
            w_out[s] = (mu + U @ (np.sqrt(sv) * z)).reshape((pu, npred)).T
    return w_out, mus, Sigs


def get_y(w, K, sd, mean):
    """SepiaEmulatorPrediction.get_y: (nsamp,npred,pu) -> (nsamp,npred,n_y); dtype follows inputs."""
    if K is None:
        ystd = w
    else:
        ystd = np.tensordot(w, K, axes=[[2], [0]])
    return ystd * sd + mean
