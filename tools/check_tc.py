"""Bring-up check of the tcgen05 rSVD passes against FP64 NumPy and the SIMT kernels (accuracy + time)."""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gladsgp_b200 import _lib  # noqa: E402
if os.environ.get('GGP_LIB'):
    _lib.LIB_PATH = os.path.abspath(os.environ['GGP_LIB'])
from gladsgp_b200 import ops  # noqa: E402


def ev(fn, it=5):
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(it):
        a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return min(ts)


def main():
    rng = np.random.default_rng(0)
    res = {}
    for (m, n, r) in [(128, 64, 8), (256, 4096, 25), (512, 20001, 25), (100, 5003, 25), (700, 9000, 40)]:
        X = (rng.standard_normal((m, n)) * rng.uniform(0.1, 3, size=(1, n))).astype(np.float32)
        Om = rng.standard_normal((r, n)).astype(np.float32)
        Xd, Od = torch.as_tensor(X, device='cuda'), torch.as_tensor(Om, device='cuda')
        ref = X.astype(np.float64) @ Om.astype(np.float64).T
        scale = np.abs(ref).max()
        y_tc = ops.rsvd_sketch_tc(Xd, Od).cpu().numpy()
        e_tc = float(np.abs(y_tc - ref).max() / scale)
        e_simt = None
        if m <= 1024:
            e_simt = float(np.abs(ops.rsvd_sketch(Xd, Od).cpu().numpy() - ref).max() / scale)
        # xty: Bt = Y^T X with Y (m, r)
        Yh = rng.standard_normal((m, r)).astype(np.float32)
        Yd = torch.as_tensor(Yh, device='cuda')
        refb = Yh.astype(np.float64).T @ X.astype(np.float64)
        sb = np.abs(refb).max()
        ex_tc = float(np.abs(ops.rsvd_xty_tc(Xd, Yd).cpu().numpy() - refb).max() / sb)
        ex_simt = float(np.abs(ops.rsvd_xty(Xd, Yd).cpu().numpy() - refb).max() / sb) if m <= 1024 else None
        res['%dx%dx%d' % (m, n, r)] = dict(err_tc=e_tc, err_simt=e_simt, xty_err_tc=ex_tc, xty_err_simt=ex_simt)
        print(m, n, r, 'sketch err tc', e_tc, 'simt', e_simt, '| xty err tc', ex_tc, 'simt', ex_simt, flush=True)
    if '--time' in sys.argv:
        m, n, r = 512, 1460000, 25
        Xd = torch.randn((m, n), dtype=torch.float32, device='cuda')
        Od = torch.randn((r, n), dtype=torch.float32, device='cuda')
        t_tc = ev(lambda: ops.rsvd_sketch_tc(Xd, Od))
        t_simt = ev(lambda: ops.rsvd_sketch(Xd, Od))
        res['time_ms'] = dict(tc=t_tc, simt=t_simt, tc_gbs=4.0 * m * n / t_tc / 1e6, simt_gbs=4.0 * m * n / t_simt / 1e6)
        d = (ops.rsvd_sketch_tc(Xd, Od) - ops.rsvd_sketch(Xd, Od)).abs().max().item()
        res['tc_vs_simt_maxabs'] = d
        Yd = torch.randn((m, r), dtype=torch.float32, device='cuda')
        t_tc = ev(lambda: ops.rsvd_xty_tc(Xd, Yd))
        t_simt = ev(lambda: ops.rsvd_xty(Xd, Yd))
        res['xty_time_ms'] = dict(tc=t_tc, simt=t_simt, tc_gbs=4.0 * m * n / t_tc / 1e6, simt_gbs=4.0 * m * n / t_simt / 1e6)
    if 'time_ms' in res:
        print('TIMES sketch tc %.3f ms simt %.3f | xty tc %.3f ms simt %.3f' % (res['time_ms']['tc'], res['time_ms']['simt'], res['xty_time_ms']['tc'], res['xty_time_ms']['simt']))
    print(json.dumps(res))


if __name__ == '__main__':
    main()
