"""Host-side cost of the e2e MCMC call (SepiaModel.do_mcmc_chains) at cfg3 size: repeated timings + cProfile."""
import cProfile
import os
import pstats
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gladsgp_b200 import synthetic  # noqa: E402
from sepia.SepiaData import SepiaData  # noqa: E402
from sepia.SepiaModel import SepiaModel  # noqa: E402

m, q, pu, chains, steps = 512, 8, 10, 59, 10
t = synthetic.design(m, q, seed=1)
y = synthetic.ensemble(t, n_x=400, n_t=36, seed=2).astype(np.float32)
ys, mu, sd = synthetic.standardize(y)
d = SepiaData(t_sim=t, y_sim=y, y_ind_sim=np.linspace(0, 1, y.shape[1]))
d.transform_xt(t_notrans=np.arange(q)); d.standardize_y(y_mean=mu, y_sd=sd)
U, S, Vh = np.linalg.svd(ys, full_matrices=False)
d.create_K_basis(K=((S[:pu, None] * Vh[:pu]) / np.sqrt(m)).astype(np.float32))
model = SepiaModel(d)
np.random.seed(0)
model.do_mcmc_chains(3, chains)
for rep in range(6):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    model.do_mcmc_chains(steps, chains)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print('e2e %d steps x %d chains: %.1f ms -> %.0f chain-steps/s' % (steps, chains, dt * 1e3, steps * chains / dt))
pr = cProfile.Profile(); pr.enable()
model.do_mcmc_chains(steps, chains)
torch.cuda.synchronize(); pr.disable()
pstats.Stats(pr).sort_stats('cumulative').print_stats(18)
