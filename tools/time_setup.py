"""Times the initialisation passes of init_model / fit_models (src/model.py:20-107, 218-224) at cfg3 size on the
device path, stage by stage, and the same NumPy lines on the host for a slice of the columns (extrapolated)."""
import json
import os
import sys
import tempfile
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gladsgp_b200 import ingest, ops, svd, synthetic, model as gmodel  # noqa: E402


def ev(fn, it=3):
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(it):
        a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
        a.record(); r = fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return min(ts), r


def main():
    m, q, pu = 512, 8, 10
    nx, nt = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (4000, 365)
    t = synthetic.design(m, q, seed=20240318)
    y = synthetic.ensemble(t, n_x=nx, n_t=nt, seed=3).astype(np.float32)
    n = y.shape[1]
    res = {'m': m, 'n_y': n, 'gb': y.nbytes / 1e9}
    torch.zeros(1, device='cuda')
    t0 = time.perf_counter(); yd = ingest.upload(y); torch.cuda.synchronize(); res['upload_s'] = time.perf_counter() - t0
    t0 = time.perf_counter(); yd = ingest.upload(y); torch.cuda.synchronize(); res['upload2_s'] = time.perf_counter() - t0
    ytd = yd.t().contiguous()
    ms, (mu, sd) = ev(lambda: ops.colstats(yd, sd_floor=1e-6)); res['colstats_ms'] = ms; res['colstats_gbs'] = y.nbytes / ms / 1e6
    ms, _ = ev(lambda: ops.colstats(ytd, transposed=True, sd_floor=1e-6)); res['colstats_t_ms'] = ms
    out = torch.empty_like(yd)
    ms, ys = ev(lambda: ops.standardize(yd, mu, sd, out=out)); res['standardize_ms'] = ms; res['standardize_gbs'] = 2 * y.nbytes / ms / 1e6
    ms, _ = ev(lambda: ops.standardize(ytd, mu, sd, transposed=True, out=out)); res['standardize_t_ms'] = ms
    del ytd
    np.random.seed(1)
    t0 = time.perf_counter(); U, S, Vh = svd.randomized_svd(ys, 25, k=0, q=1); torch.cuda.synchronize(); res['rsvd_s'] = time.perf_counter() - t0
    K = ((S[:pu, None] * Vh[:pu]) / np.sqrt(m)).astype(np.float32)
    Kd = ingest.upload(K)
    ms, _ = ev(lambda: ops.project(ys, Kd)); res['project_ms'] = ms
    res['project_gflops_fp64'] = 2.0 * m * n * (pu + 2) / ms / 1e6
    t0 = time.perf_counter(); proj = ingest.project_basis(ys, Kd); res['project_basis_s'] = time.perf_counter() - t0
    res['pc_prec'] = ingest.pc_precision_from(proj)
    del yd, ys, out, Kd
    torch.cuda.empty_cache()
    # whole init_model + pc_precision through the mirror of src/model.py (device path)
    with tempfile.TemporaryDirectory() as td:
        np.random.seed(1)
        t0 = time.perf_counter()
        data, mdl = gmodel.init_model(t, y, 'bench', pu, data_dir=td, recompute=True)
        pc = gmodel.pc_precision(data.sim_data)
        torch.cuda.synchronize()
        res['init_model_device_s'] = time.perf_counter() - t0
        res['pc_prec_api'] = pc
    # the reference's NumPy lines on a column slice (src/model.py:60-72, 219-223), extrapolated to n_y
    ns = min(n, 100000)
    ysl = np.ascontiguousarray(y[:, :ns])
    t0 = time.perf_counter()
    mu_h = np.mean(ysl, axis=0); sd_h = np.std(ysl, ddof=1, axis=0); sd_h[sd_h < 1e-6] = 1e-6
    ystd_h = (ysl - mu_h) / sd_h
    Ks = K[:, :ns]
    w = np.dot(np.linalg.pinv(Ks).T, ystd_h.T).T
    pc_var = np.var(ystd_h - np.dot(w, Ks))
    res['host_numpy_lines_s_extrapolated'] = (time.perf_counter() - t0) * n / ns
    print(json.dumps(res))


if __name__ == '__main__':
    main()
