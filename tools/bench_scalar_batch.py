"""fit_scalar_models.py workload (15 scalar QoI models, m = 512, 8 parameters: tune_step_sizes(100, 10) + do_mcmc(512)
each): one by one through SepiaModel vs together through gladsgp_b200.batch.ModelBatch."""
import io
import json
import os
import sys
import time
from contextlib import redirect_stdout

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gladsgp_b200 import synthetic  # noqa: E402
from gladsgp_b200.batch import ModelBatch  # noqa: E402
from sepia.SepiaData import SepiaData  # noqa: E402
from sepia.SepiaModel import SepiaModel  # noqa: E402

m, q, nmodels = 512, 8, int(sys.argv[1]) if len(sys.argv) > 1 else 15
t = synthetic.design(m, q, seed=3)
rng = np.random.default_rng(0)
ys = [np.sin(t @ rng.uniform(0.5, 3, size=q)) + 0.05 * rng.standard_normal(m) for _ in range(nmodels)]


def make(y):
    d = SepiaData(t_sim=t, y_sim=y.astype(np.float32))
    d.transform_xt(); d.standardize_y()
    return SepiaModel(d)


res = {'models': nmodels, 'm': m}
mm = make(ys[0])
with redirect_stdout(io.StringIO()):
    np.random.seed(1); mm.tune_step_sizes(3, 3, prog=False)           # warm the path
    mm = make(ys[0])
    torch.cuda.synchronize(); t0 = time.perf_counter()
    np.random.seed(1); mm.tune_step_sizes(100, 10, prog=False); mm.do_mcmc(512, prog=False)
    torch.cuda.synchronize(); res['one_model_s'] = time.perf_counter() - t0
models = [make(y) for y in ys]
b = ModelBatch(models, seeds=range(nmodels))
torch.cuda.synchronize(); t0 = time.perf_counter()
b.tune_step_sizes(100, 10); b.do_mcmc(512)
torch.cuda.synchronize(); res['batch_s'] = time.perf_counter() - t0
res['one_by_one_s_extrapolated'] = res['one_model_s'] * nmodels
res['speedup'] = res['one_by_one_s_extrapolated'] / res['batch_s']
res['steps_per_model'] = 10 + 1000 + 512
print(json.dumps(res))
