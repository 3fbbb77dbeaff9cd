"""Developer: cProfile of the reference-facing prediction call (host side).   python tools/cprof_predict.py [npred]"""
import cProfile, pstats, os, sys, io
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from gladsgp_b200 import svd, synthetic
from sepia.SepiaData import SepiaData
from sepia.SepiaModel import SepiaModel
from sepia.SepiaPredict import SepiaEmulatorPrediction
npred = int(sys.argv[1]) if len(sys.argv) > 1 else 256
t, y, mu, sd = bench.build_problem(400, 36, standardized=False)
data = SepiaData(t_sim=t, y_sim=y, y_ind_sim=np.linspace(0, 1, y.shape[1]))
data.transform_xt(t_notrans=np.arange(8)); data.standardize_y(y_mean=mu, y_sd=sd)
np.random.seed(1)
U, S, Vh = svd.randomized_svd(data.sim_data.y_std, 25, k=0, q=1)
data.create_K_basis(K=((S[:10, None] * Vh[:10]) / np.sqrt(512)).astype(np.float32))
model = SepiaModel(data)
samples = synthetic.posterior_samples(64, 9, 10, seed=77)
tp = synthetic.test_design(npred * 4, 8)
for i in range(3):
    SepiaEmulatorPrediction(t_pred=tp[:npred], samples=samples, model=model)
torch.cuda.synchronize()
pr = cProfile.Profile()
pr.enable()
for i in range(40):
    SepiaEmulatorPrediction(t_pred=tp[(i % 4) * npred:(i % 4 + 1) * npred], samples=samples, model=model)
pr.disable()
s = io.StringIO()
pstats.Stats(pr, stream=s).sort_stats('tottime').print_stats(22)
print(s.getvalue())
