"""Developer A/B run of one library variant (GGP_LIB=<.so>): batched fused log-likelihood and the cfg3 sampler.

Prints one JSON line: evaluations/s of ggp_loglik_batched_f64 (m=512, d=9) for the batch sizes in GGP_BS, chain-steps/s and
sweep-kernel evaluations/s of the sampler (GGP_CHAINS chains x 10 PCs, GGP_STEPS steps), and SHA-1 digests of the
log-likelihood vector and of the final log-posteriors: every variant of the evaluation kernel must print the same digests.
"""
import hashlib
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gladsgp_b200 import _lib  # noqa: E402
if os.environ.get('GGP_LIB'):
    _lib.LIB_PATH = os.environ['GGP_LIB']
from gladsgp_b200 import ops, synthetic  # noqa: E402
import bench  # noqa: E402


def ev_time(fn, iters=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return min(ts), float(np.median(ts))


def digest(t):
    return hashlib.sha1(t.detach().cpu().numpy().tobytes()).hexdigest()[:12]


def main():
    res = {'lib': os.path.basename(_lib.LIB_PATH)}
    m, q, pu = 512, 8, 10
    d = q + 1
    t = synthetic.design(m, q)
    X = np.concatenate([0.5 * np.ones((m, 1)), t.astype(np.float64)], axis=1)
    rng = np.random.default_rng(0)
    flop_eval = m ** 3 / 3.0 + m * m + 2 * m + (3 * d + 2) * m * (m - 1) / 2.0
    for B in [int(x) for x in os.environ.get('GGP_BS', '592,2368').split(',') if x]:
        beta = np.exp(rng.uniform(np.log(0.05), np.log(3.0), size=(B, d)))
        lamz = rng.uniform(0.5, 2.0, B); dadd = rng.uniform(1e-3, 1e-2, B)
        W = rng.standard_normal((B, m))
        Xd = torch.as_tensor(X, device='cuda'); Wd = torch.as_tensor(W, device='cuda')
        bd = torch.as_tensor(beta, device='cuda'); ld = torch.as_tensor(lamz, device='cuda'); dd = torch.as_tensor(dadd, device='cuda')
        ws = torch.empty((B, ops._lib.load().ggp_factor_doubles(m)), dtype=torch.float64, device='cuda')
        out = ops.loglik_batched(Xd, Wd, bd, ld, dd, factor_ws=ws)
        ll = out['loglik']
        best, med = ev_time(lambda: ops.loglik_batched(Xd, Wd, bd, ld, dd, factor_ws=ws), iters=int(os.environ.get('GGP_ITERS', '7')))
        res['loglik_B%d' % B] = dict(ms=round(best, 4), kevals_s=round(B / best, 1), kevals_s_med=round(B / med, 1),
                                     tflops=round(B * flop_eval / best / 1e9, 2), sha=digest(ll))
        del ws
    chains = int(str(os.environ.get('GGP_CHAINS', '236')).split(',')[0])
    steps = int(os.environ.get('GGP_STEPS', '10'))
    if chains > 0:
        from gladsgp_b200 import svd, model as gmodel
        from sepia.SepiaData import SepiaData
        from sepia.SepiaModel import SepiaModel
        tt, y, mu, sd = bench.build_problem(400, 36, standardized=False)
        data = SepiaData(t_sim=tt, y_sim=y, y_ind_sim=np.linspace(0, 1, y.shape[1]))
        data.transform_xt(t_notrans=np.arange(q)); data.standardize_y(y_mean=mu, y_sd=sd)
        np.random.seed(1)
        U, S, Vh = svd.randomized_svd(data.sim_data.y_std, 25, k=0, q=1)
        K = ((S[:pu, None] * Vh[:pu]) / np.sqrt(m)).astype(np.float32)
        data.create_K_basis(K=K)
        model = SepiaModel(data)
        gmodel.override_lamWOs(model, gmodel.pc_precision(data.sim_data))
        for ch in [int(x) for x in str(os.environ.get('GGP_CHAINS', '236')).split(',')]:
            eng, tb = model._get_engine(ch)
            P = eng.P
            rs = np.random.RandomState(100)
            us = torch.as_tensor(rs.random_sample((ch, 2 * P * (steps + 5)))).to('cuda')
            eng.run(5, tb['step'], uniforms=us[:, :2 * P * 5].contiguous(), record=False)
            ust = us[:, 2 * P * 5:].contiguous()
            th0, s0 = eng.theta.clone(), eng.sigwl.clone()
            best = 1e30
            for rep in range(3):
                eng.theta.copy_(th0); eng.sigwl.copy_(s0)
                torch.cuda.synchronize()
                e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
                e0.record(); out = eng.run(steps, tb['step'], uniforms=ust, init_sigwl=False, record=True); e1.record()
                torch.cuda.synchronize()
                best = min(best, e0.elapsed_time(e1))
            lp_sha = digest(out['lp'][-1])
            eng.theta.copy_(th0); eng.sigwl.copy_(s0)
            out2 = eng.run(steps, tb['step'], uniforms=ust, init_sigwl=False, record=True, time_kernels=True)
            sw_ms, wo_ms = out2['kernel_ms']
            nev = int(out2['eval_count'][0]) + int(out2['eval_count'][1])      # sweep sites + lamWOs terms (one kernel since v2)
            sw_ms = sw_ms + wo_ms
            res['mcmc_c%d' % ch] = dict(chain_steps_s=round(ch * steps / best * 1e3, 1), ms_per_step=round(best / steps, 3),
                                        sweep_kevals_s=round(nev / sw_ms, 1) if sw_ms > 0 else None,
                                        sweep_frac_of_35p44=round(nev * flop_eval / (sw_ms * 1e-3) / 35.44e12, 4) if sw_ms > 0 else None,
                                        sweep_ms=round(sw_ms, 2), wos_ms=round(wo_ms, 2), lp_sha=lp_sha)
    print(json.dumps(res))


if __name__ == '__main__':
    main()
