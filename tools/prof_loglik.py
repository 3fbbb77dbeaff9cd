"""Tiny driver for ncu: a few launches of the batched fused log-likelihood kernel at m=512, d=9."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gladsgp_b200 import ops, synthetic
m, q = 512, 8
B = int(sys.argv[1]) if len(sys.argv) > 1 else 148
d = q + 1
t = synthetic.design(m, q)
X = np.concatenate([0.5 * np.ones((m, 1)), t.astype(np.float64)], axis=1)
rng = np.random.default_rng(0)
beta = np.exp(rng.uniform(np.log(0.05), np.log(3.0), size=(B, d)))
lamz = rng.uniform(0.5, 2.0, B); dadd = rng.uniform(1e-3, 1e-2, B); W = rng.standard_normal((B, m))
Xd, Wd, bd, ld, dd = [torch.as_tensor(a, device='cuda') for a in (X, W, beta, lamz, dadd)]
ws = torch.empty((B, ops._lib.load().ggp_factor_doubles(m)), dtype=torch.float64, device='cuda')
for _ in range(3):
    out = ops.loglik_batched(Xd, Wd, bd, ld, dd, factor_ws=ws)
torch.cuda.synchronize()
print('ok', out['loglik'][:3].cpu().numpy())
