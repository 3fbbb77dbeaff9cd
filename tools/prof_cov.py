"""Tiny driver for ncu: a few launches of the materialised covariance build at m=512, d=9."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gladsgp_b200 import ops, synthetic
m, q = 512, 8
B = int(sys.argv[1]) if len(sys.argv) > 1 else 592
d = q + 1
t = synthetic.design(m, q)
X = np.concatenate([0.5 * np.ones((m, 1)), t.astype(np.float64)], axis=1)
rng = np.random.default_rng(0)
beta = np.exp(rng.uniform(np.log(0.05), np.log(3.0), size=(B, d)))
lamz = rng.uniform(0.5, 2.0, B); dadd = rng.uniform(1e-3, 1e-2, B)
Xd, bd, ld, dd = [torch.as_tensor(a, device='cuda') for a in (X, beta, lamz, dadd)]
for _ in range(3):
    out = ops.cov_build(Xd, bd, ld, dd)
torch.cuda.synchronize()
print('ok', float(out[0, 0, 0]))
