"""Generates the committed golden fixtures.  Run in the BUILD container (needs /root/reference):

    python tests/golden/make_golden.py

* rsvd_reference.npz  -- outputs of the reference's OWN src/svd.py (imported from /root/reference,
  unmodified) on a seeded input with a seeded global np.random stream.  Pins oracle/svd_oracle.py
  and the CUDA rSVD.
* sepia_oracle.npz    -- outputs of oracle/sepia_oracle.py (the SEPIA restatement; parity
  unpinned, see its header) on a small seeded problem: covariance, per-PC log-likelihood terms,
  a 6-step chain with its replay tensors, prediction mu/Sigma.  Regression anchor for the oracle and
  a size-independent fixture for the GPU path.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))


def make_rsvd():
    sys.path.insert(0, '/root/reference')
    from src import svd as ref_svd                      # the reference's own implementation
    rng = np.random.default_rng(77)
    m, n = 40, 600
    a = rng.standard_normal((m, 12)) * (0.6 ** np.arange(12))
    X = (a @ rng.standard_normal((12, n)) + 0.01 * rng.standard_normal((m, n))).astype(np.float32)
    out = {'X': X}
    for tag, (p, k, q) in {'a': (8, 0, 1), 'b': (5, None, 2), 'c': (25, 0, 1)}.items():
        np.random.seed(1000 + p)
        U, S, Vh = ref_svd.randomized_svd(X, p, k=k, q=q)
        out['U_' + tag], out['S_' + tag], out['Vh_' + tag] = U, S, Vh
        out['pkq_' + tag] = np.array([p, -1 if k is None else k, q])
    np.random.seed(5)
    (U, S, Vh), err = ref_svd.randomized_svd(X, 6, k=0, q=1, return_error=True)
    out['err_bound'] = np.array(err, dtype=np.float64)
    np.savez_compressed(os.path.join(HERE, 'rsvd_reference.npz'), **out)


def make_sepia():
    from helpers import so, make_problem, tables_from_oracle, replay_from_trace, random_hypers
    from gladsgp_b200 import synthetic
    pr = make_problem(m=24, q=2, pu=2, n_x=6, n_t=5, seed=9)
    num = pr['num']
    out = dict(t=pr['t'], y=pr['y'], mu=pr['mu'], sd=pr['sd'], K=pr['K'], zt=num.zt, w=num.w, LamSim=num.LamSim,
               resid_ss=np.array(num.resid_ss))
    beta, lamz, lamws, lamwos = random_hypers(num, 4, seed=3)
    js = np.arange(4) % 2
    out.update(beta=beta, lamz=lamz, lamws=lamws, lamwos=lamwos, js=js)
    C = np.stack([so.block_cov(num, beta[b], lamz[b], lamws[b], lamwos[b], js[b]) for b in range(4)])
    ll = np.array([so.do_loglik(C[b], num.wv[js[b] * 24:(js[b] + 1) * 24, 0]) for b in range(4)])
    out.update(C=C, loglik=ll)
    mod = so.OracleModel(num)
    mod.override_lamWOs(40.0)
    tb = tables_from_oracle(mod)
    P = tb['theta'].size
    mod.trace = []
    mod.do_mcmc(6, rng=np.random.RandomState(21))
    replay, acc = replay_from_trace(mod.trace, 6, P)
    s = mod.get_samples()
    out.update({'tb_' + k: v for k, v in tb.items()})
    out.update({'rp_' + k: v for k, v in replay.items()})
    out.update(chain_acc=acc, chain_draws=np.concatenate([s['betaU'], s['lamUz'], s['lamWs'], s['lamWOs']], axis=1),
               chain_lp=s['logPost'][:, 0])
    samples = synthetic.posterior_samples(3, num.d, 2, seed=4)
    tp = synthetic.test_design(3, 2, seed=8)
    _, mu, Sig = so.w_pred(num, tp, samples, draw=False)
    out.update({'ps_' + k: v for k, v in samples.items()})
    out.update(t_pred=tp, pred_mu=mu, pred_Sigma=Sig)
    np.savez_compressed(os.path.join(HERE, 'sepia_oracle.npz'), **out)


if __name__ == '__main__':
    make_rsvd()
    make_sepia()
    for f in ('rsvd_reference.npz', 'sepia_oracle.npz'):
        print(f, os.path.getsize(os.path.join(HERE, f)), 'bytes')
