"""Developer micro-benchmarks (not the contract bench): FP64 GEMM peak + batched loglik timing."""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gladsgp_b200 import _lib  # noqa: E402
if os.environ.get('GGP_LIB'):
    _lib.LIB_PATH = os.environ['GGP_LIB']
from gladsgp_b200 import ops, synthetic  # noqa: E402


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


def main():
    res = {}
    if '--peaks' in sys.argv:
        for dt, name in ((torch.float64, 'fp64'), (torch.float32, 'fp32')):
            torch.backends.cuda.matmul.allow_tf32 = False
            n = 8192
            a = torch.randn(n, n, dtype=dt, device='cuda'); b = torch.randn(n, n, dtype=dt, device='cuda')
            best, med = ev_time(lambda: torch.matmul(a, b), iters=5)
            res[name + '_gemm_tflops'] = 2 * n ** 3 / best / 1e9
            res[name + '_gemm_tflops_median'] = 2 * n ** 3 / med / 1e9
        x = torch.empty(1 << 28, dtype=torch.float64, device='cuda')
        best, med = ev_time(lambda: x.fill_(1.0), iters=5)
        res['hbm_write_gbs'] = x.numel() * 8 / best / 1e6
    m, q, pu = 512, 8, 10
    d = q + 1
    t = synthetic.design(m, q)
    X = np.concatenate([0.5 * np.ones((m, 1)), t.astype(np.float64)], axis=1)
    rng = np.random.default_rng(0)
    Bs = [int(x) for x in os.environ.get('GGP_BS', '1,10,148,296,444,888').split(',')]
    for B in Bs:
        beta = np.exp(rng.uniform(np.log(0.05), np.log(3.0), size=(B, d)))
        lamz = rng.uniform(0.5, 2.0, B); dadd = rng.uniform(1e-3, 1e-2, B)
        W = rng.standard_normal((B, m))
        Xd = torch.as_tensor(X, device='cuda'); Wd = torch.as_tensor(W, device='cuda')
        bd = torch.as_tensor(beta, device='cuda'); ld = torch.as_tensor(lamz, device='cuda'); dd = torch.as_tensor(dadd, device='cuda')
        ws = torch.empty((B, ops._lib.load().ggp_factor_doubles(m)), dtype=torch.float64, device='cuda')
        best, med = ev_time(lambda: ops.loglik_batched(Xd, Wd, bd, ld, dd, factor_ws=ws), iters=5)
        fl = B * (m ** 3 / 3 + m * m)
        res['loglik_m512_B%d' % B] = dict(ms=best, ms_median=med, evals_per_s=B / best * 1e3, tflops=fl / best / 1e9)
        if os.environ.get('GGP_NOCOV'):
            continue
        best, med = ev_time(lambda: ops.cov_build(Xd, bd, ld, dd), iters=5)
        res['cov_m512_B%d' % B] = dict(ms=best, gbs=B * m * m * 8 / best / 1e6)
    print(json.dumps(res, indent=1))
    os.makedirs('gpurun_out', exist_ok=True)
    json.dump(res, open('gpurun_out/quick_bench.json', 'w'), indent=1)


if __name__ == '__main__':
    main()
