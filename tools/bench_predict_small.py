"""Developer: device time of predict_kernel (+ joint covariance, draw) for small and large calls, cfg3 model (m=512, d=9, 640 blocks).
   GGP_LIB=<variant .so> python tools/bench_predict_small.py"""
import os, sys, json
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from gladsgp_b200 import _lib
if os.environ.get('GGP_LIB'):
    _lib.LIB_PATH = os.environ['GGP_LIB']
from gladsgp_b200 import ops, synthetic
m, q, pu, ns = 512, 8, 10, 64
d = q + 1
t = synthetic.design(m, q)
X = np.concatenate([0.5 * np.ones((m, 1)), t.astype(np.float64)], axis=1)
s = synthetic.posterior_samples(ns, d, pu, seed=77)
beta = np.asarray(s['betaU'], dtype=np.float64).reshape(ns, pu, d).reshape(ns * pu, d)
lamz = np.asarray(s['lamUz'], dtype=np.float64).reshape(-1)
lamw = np.asarray(s['lamWs'], dtype=np.float64).reshape(-1)
dadd = 1.0 / lamw + 1e-3
W = np.random.default_rng(0).standard_normal((ns * pu, m))
P = ops.Predictor(X, W, beta, lamz, dadd, 1.0 / lamz + 1.0 / lamw)
tp = synthetic.test_design(8192, q)
xp = torch.as_tensor(np.concatenate([0.5 * np.ones((8192, 1)), tp.astype(np.float64)], axis=1), device='cuda')
def ev(fn, reps=7):
    fn(); torch.cuda.synchronize(); best = 1e9
    for _ in range(reps):
        a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
        a.record(); out = fn(); b.record(); torch.cuda.synchronize(); best = min(best, a.elapsed_time(b))
    return best, out
res = {'lib': os.environ.get('GGP_LIB', 'default')}
ref = None
for n in (1, 4, 16, 64, 256, 8192):
    x = xp[:n].contiguous()
    ms, (mean, var) = ev(lambda: P.predict(x))
    res['predict_%d_ms' % n] = round(ms, 4)
    res['sum_%d' % n] = float(mean.sum().item()) + float(var.sum().item())
ms, (mean, var, V) = ev(lambda: P.predict(xp[:256].contiguous(), want_V=True))
res['predict_V_256_ms'] = round(ms, 4)
print(json.dumps(res))
