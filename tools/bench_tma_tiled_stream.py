"""Developer experiment: the sketch pass's pure TMA stream (dry mode 9: no split, no MMA) over the row-major ensemble (256 rows x 128 B
boxes, rows 5.8 MB apart) against the same bytes described as contiguous 32 KB tiles (GGP_TMA_TILED=1) -- what a tile-major copy of
the ensemble would stream.  Wrong results by construction; timing only."""
import json, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from gladsgp_b200 import _lib
if os.environ.get('GGP_LIB'):
    _lib.LIB_PATH = os.environ['GGP_LIB']
from gladsgp_b200 import ops
m, n, r = 512, 1460000 - 1460000 % 256, 25        # m * n a multiple of 32 * 256
g = torch.Generator(device='cuda'); g.manual_seed(0)
X = torch.randn((m, n), dtype=torch.float32, device='cuda', generator=g)
omT = torch.randn((r, n), dtype=torch.float32, device='cuda', generator=g)
ws = torch.empty(_lib.load().ggp_rsvd_tc_workspace_bytes(m), dtype=torch.uint8, device='cuda')
def ev(fn, reps=7):
    fn(); torch.cuda.synchronize(); best = 1e9
    for _ in range(reps):
        a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); best = min(best, a.elapsed_time(b))
    return best
res = {}
for mode in ('8', '9'):
    for tiled in ('', '1'):
        os.environ['GGP_TMA'] = mode
        if tiled: os.environ['GGP_TMA_TILED'] = '1'
        else: os.environ.pop('GGP_TMA_TILED', None)
        ms = ev(lambda: ops.rsvd_sketch_tc(X, omT, ws))
        res['dry%s%s' % (mode, '_tiled' if tiled else '')] = dict(ms=ms, gbs=4.0 * m * n / ms / 1e6, frac=4.0 * m * n / ms / 1e6 / 6533.8)
os.environ.pop('GGP_TMA_TILED', None); os.environ['GGP_TMA'] = '3'
ms = ev(lambda: ops.rsvd_sketch_tc(X, omT, ws)); res['real_mode3'] = dict(ms=ms, gbs=4.0 * m * n / ms / 1e6)
ms = ev(lambda: X.view(-1).sum()); res['torch_sum_read_only'] = dict(ms=ms, gbs=4.0 * m * n / ms / 1e6)
print(json.dumps(res))
