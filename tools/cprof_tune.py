"""Developer: cProfile of model.tune_step_sizes(100, 5) + do_mcmc(512) at cfg3 (the reference's own workload, src/model.py:234-235)."""
import cProfile, pstats, io, os, sys, time, contextlib
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from gladsgp_b200 import svd, model as gmodel
from sepia.SepiaData import SepiaData
from sepia.SepiaModel import SepiaModel
t, y, mu, sd = bench.build_problem(400, 36, standardized=False)
data = SepiaData(t_sim=t, y_sim=y, y_ind_sim=np.linspace(0, 1, y.shape[1]))
data.transform_xt(t_notrans=np.arange(8)); data.standardize_y(y_mean=mu, y_sd=sd)
np.random.seed(1)
U, S, Vh = svd.randomized_svd(data.sim_data.y_std, 25, k=0, q=1)
data.create_K_basis(K=((S[:10, None] * Vh[:10]) / np.sqrt(512)).astype(np.float32))
model = SepiaModel(data)
gmodel.override_lamWOs(model, gmodel.pc_precision(data.sim_data))
np.random.seed(2024)
model.do_mcmc(8, prog=False); model.clear_samples(); torch.cuda.synchronize()
for rep in range(2):
    pr = cProfile.Profile()
    buf = io.StringIO()
    t0 = time.perf_counter()
    pr.enable()
    with contextlib.redirect_stdout(buf):
        model.tune_step_sizes(100, 5, prog=False)
    pr.disable()
    t1 = time.perf_counter()
    model.do_mcmc(512, prog=False); torch.cuda.synchronize()
    t2 = time.perf_counter()
    print('rep %d: tune %.3f s, do_mcmc(512) %.3f s' % (rep, t1 - t0, t2 - t1))
    s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats('cumtime').print_stats(25); print(s.getvalue()[:6000])
    model.clear_samples()
