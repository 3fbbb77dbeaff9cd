"""Sobol' / Saltelli indices with bootstrap limits at the reference's size (sensitivity_indices.py:96: n_dim = 8, m = 8 ->
N = 256 designs per block, 10 PCs, 9999 resamples x 4 statistics): oracle (the reference's algorithm through
scipy.stats.bootstrap, CPU) vs gladsgp_b200.sensitivity (statistics on the GPU).  python tools/bench_sobol.py [--cpu-only]"""
import json
import os
import sys
import time
import warnings

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import sobol_oracle as sob  # noqa: E402

n_dim, m, p = 8, 8, 10
rng = np.random.default_rng(0)
Wt = rng.standard_normal((n_dim, p))


def func(x):            # stand-in for the emulator: smooth, vector valued
    return np.sin(2.0 * x @ Wt) + (x ** 2) @ np.abs(Wt) * 0.5


pcvar = np.linspace(1.0, 0.1, p); pcvar /= pcvar.sum()
AB = sob.sobol_matrix(n_dim, m, seed=1)
res = {}
t0 = time.perf_counter()
with warnings.catch_warnings():
    warnings.simplefilter('ignore')
    ref = sob.PCA_saltelli_sensitivity_indices(func, n_dim, m, pcvar, AB=AB, rng=np.random.default_rng(5))
res['cpu_oracle_s'] = time.perf_counter() - t0
if '--cpu-only' not in sys.argv:
    import torch
    from gladsgp_b200 import sensitivity
    sensitivity.PCA_saltelli_sensitivity_indices(func, n_dim, 4, pcvar, n_resamples=99, rng=np.random.default_rng(5))   # warm-up
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    got = sensitivity.PCA_saltelli_sensitivity_indices(func, n_dim, m, pcvar, AB=AB, rng=np.random.default_rng(5))
    torch.cuda.synchronize()
    res['gpu_s'] = time.perf_counter() - t0
    diffs = [np.abs(np.array(got[4][k].confidence_interval) - np.array(ref[4][k].confidence_interval)) for k in ref[4]]
    res['max_abs_diff_ci'] = float(max(np.nanmax(d) for d in diffs))        # NaN limits (degenerate statistics) are NaN on both sides
    res['nan_pattern_equal'] = bool(all(np.array_equal(np.isnan(np.array(got[4][k].confidence_interval)), np.isnan(np.array(ref[4][k].confidence_interval))) for k in ref[4]))
    res['speedup'] = res['cpu_oracle_s'] / res['gpu_s']
print(json.dumps(res))
