"""Golden chains at the BASELINE.json configuration sizes (oracle/sepia_oracle.py; SEPIA parity unpinned, see its header).

    python tests/golden/make_golden_chains.py [cfg1 cfg2 cfg3]

For each configuration: the model data the sampler needs (zt, w, LamSim, per-element tables), the oracle's chain with its
per-site replay record (candidate, log aCorr, log u, validity), the accept flags, the draws, the log-posteriors, and the
decision margin of every evaluated site

    margin = (clp - lp + log aCorr) - log u          (full sums, as SepiaModel.mcmc_step forms them)

together with the same margin in the per-PC difference form the device uses: the two agree to ~1e-12 and no decision
lies anywhere near that (tests/test_oracle_cpu.py::test_accept_margins..., tests/test_gpu_chains.py).

  cfg1  experiments/synthetic-like: m=100, q=8 (d=9), pu=5, reference lamWOs override, 200 steps
  cfg2  1-D GP regression, scalar output: m=1000, d=2, pu=1, default priors, 50 steps
  cfg3  multivariate PCA emulator: m=512, q=8 (d=9), pu=10, reference lamWOs override, 30 steps
"""
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))

from helpers import so, svd_oracle, tables_from_oracle, replay_from_trace, make_scalar_problem  # noqa: E402
from gladsgp_b200 import synthetic  # noqa: E402


def pca_problem(m, q, pu, n_x, n_t, seed):
    t = synthetic.design(m, q, seed=20240318)
    y = synthetic.ensemble(t, n_x=n_x, n_t=n_t, seed=seed)
    y_std, mu, sd = synthetic.standardize(y)
    U, S, Vh = svd_oracle.randomized_svd(y_std.astype(np.float32), 25, k=0, q=1, rng=np.random.RandomState(seed))
    K = svd_oracle.k_basis(S, Vh, pu, m).astype(np.float32)
    num = so.OracleNum(t, y_std.astype(np.float32), K)
    w = num.w
    pc_prec = 1.0 / np.var(np.asarray(y_std, dtype=np.float64) - w @ np.asarray(K, dtype=np.float64))
    return num, pc_prec


def run_chain(num, n_steps, seed, pc_prec=None):
    mod = so.OracleModel(num)
    if pc_prec is not None:
        mod.override_lamWOs(pc_prec)                  # src/model.py:225-231
    tb = tables_from_oracle(mod)
    P = tb['theta'].size
    mod.trace = []
    t0 = time.time()
    mod.do_mcmc(n_steps, rng=np.random.RandomState(seed))
    dt = time.time() - t0
    replay, acc = replay_from_trace(mod.trace, n_steps, P)
    s = mod.get_samples()
    draws = np.concatenate([s['betaU'], s['lamUz'], s['lamWs'], s['lamWOs']], axis=1)
    margin = np.full((n_steps, P), np.nan)
    for i, tr in enumerate(mod.trace):
        if tr['valid']:
            margin[i // P, i % P] = (tr['clp'] - tr['lp'] + np.log(tr['aCorr'])) - np.log(tr['u2'])
    out = dict(zt=num.zt, w=num.w, LamSim=np.asarray(num.LamSim, dtype=np.float64), n_steps=np.array(n_steps),
               chain_acc=acc, chain_draws=draws, chain_lp=s['logPost'][:, 0], margin=margin,
               oracle_seconds=np.array(dt))
    out.update({'tb_' + k: v for k, v in tb.items()})
    out.update({'rp_' + k: v for k, v in replay.items()})
    return out


def main():
    which = sys.argv[1:] or ['cfg1', 'cfg2', 'cfg3']
    for cfg in which:
        if cfg == 'cfg1':
            num, pc = pca_problem(100, 8, 5, 60, 12, seed=20240319)
            out = run_chain(num, 200, seed=101, pc_prec=pc)
        elif cfg == 'cfg2':
            pr = make_scalar_problem(m=1000, seed=20240320)
            out = run_chain(pr['num'], 50, seed=102)
        elif cfg == 'cfg3':
            num, pc = pca_problem(512, 8, 10, 100, 12, seed=20240321)
            out = run_chain(num, 30, seed=103, pc_prec=pc)
        else:
            raise SystemExit('unknown configuration ' + cfg)
        path = os.path.join(HERE, 'chain_%s.npz' % cfg)
        np.savez_compressed(path, **out)
        mg = out['margin'][np.isfinite(out['margin'])]
        print(cfg, 'steps', int(out['n_steps']), 'decisions', mg.size, 'min |margin| %.3e' % np.min(np.abs(mg)),
              'accept rate %.3f' % out['chain_acc'].mean(), '%.1f s' % float(out['oracle_seconds']),
              os.path.getsize(path), 'bytes')


if __name__ == '__main__':
    main()
