"""Developer timing of the reference-facing prediction call (SepiaEmulatorPrediction, default joint behaviour) by phase:
   python tools/prof_predict_e2e.py [npred] [nsamp]"""
import os, sys, time, json
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from gladsgp_b200 import svd, model as gmodel, synthetic, ops
from sepia.SepiaData import SepiaData
from sepia.SepiaModel import SepiaModel
from sepia.SepiaPredict import SepiaEmulatorPrediction
npred = int(sys.argv[1]) if len(sys.argv) > 1 else 256
nsamp = int(sys.argv[2]) if len(sys.argv) > 2 else 64
t, y, mu, sd = bench.build_problem(400, 36, standardized=False)
data = SepiaData(t_sim=t, y_sim=y, y_ind_sim=np.linspace(0, 1, y.shape[1]))
data.transform_xt(t_notrans=np.arange(8)); data.standardize_y(y_mean=mu, y_sd=sd)
np.random.seed(1)
U, S, Vh = svd.randomized_svd(data.sim_data.y_std, 25, k=0, q=1)
data.create_K_basis(K=((S[:10, None] * Vh[:10]) / np.sqrt(512)).astype(np.float32))
model = SepiaModel(data)
samples = synthetic.posterior_samples(nsamp, 9, 10, seed=77)
tp = synthetic.test_design(npred * 6, 8)
res = {}
SepiaEmulatorPrediction(t_pred=tp[:npred], samples=samples, model=model)
torch.cuda.synchronize()
ts = []
for i in range(1, 6):
    t0 = time.perf_counter(); pe = SepiaEmulatorPrediction(t_pred=tp[i * npred:(i + 1) * npred], samples=samples, model=model); ts.append(time.perf_counter() - t0)
res['call_ms'] = [round(x * 1e3, 2) for x in ts]
pr = pe._pred
xp = torch.as_tensor(pe.xpredt, device='cuda')
def ev(fn, reps=5):
    fn(); torch.cuda.synchronize(); best = 1e9
    for _ in range(reps):
        a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
        a.record(); out = fn(); b.record(); torch.cuda.synchronize(); best = min(best, a.elapsed_time(b))
    return best, out
res['predict_V_ms'], (mean, var, V) = ev(lambda: pr.predict(xp, want_V=True))
res['predict_noV_ms'], _ = ev(lambda: pr.predict(xp))
res['pred_cov_ms'], Sig = ev(lambda: pr.pred_cov(xp, V))
S4 = Sig.reshape(nsamp, 10, npred, npred)
res['cholesky_ex_ms'], (L, info) = ev(lambda: torch.linalg.cholesky_ex(S4))
z = torch.randn(nsamp, 10, npred, dtype=torch.float64, device='cuda')
res['matmul_ms'], _ = ev(lambda: torch.matmul(L, z.unsqueeze(-1)))
res['chol_draw_ms'], (dev, inf2) = ev(lambda: ops.chol_draw(Sig, z.reshape(nsamp * 10, npred)))     # what the class runs
res['chol_draw_max_abs_diff_vs_torch'] = float((dev.reshape(nsamp, 10, npred) - torch.matmul(L, z.unsqueeze(-1)).squeeze(-1)).abs().max())
t0 = time.perf_counter(); zz = np.random.normal(size=nsamp * 10 * npred); res['host_normal_ms'] = (time.perf_counter() - t0) * 1e3
t0 = time.perf_counter(); k = __import__('gladsgp_b200.sepia.SepiaPredict', fromlist=['x'])._samples_key(samples, model, False); res['samples_key_ms'] = (time.perf_counter() - t0) * 1e3
res['pairs_per_call'] = nsamp * npred
res['e2e_pairs_per_s'] = nsamp * npred / np.median(ts)
print(json.dumps(res))
