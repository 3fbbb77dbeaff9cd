"""Developer timing of the two rSVD passes at cfg3 size (512 x 1.46 M float32, r = 25): GGP_TMA=0 (register-staged) / 1 / 2 / 3."""
import json, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from gladsgp_b200 import _lib
if os.environ.get('GGP_LIB'):                      # developer A/B runs against a variant library
    _lib.LIB_PATH = os.environ['GGP_LIB']
from gladsgp_b200 import ops
m, n, r = 512, 1460000, 25
g = torch.Generator(device='cuda'); g.manual_seed(0)
X = torch.randn((m, n), dtype=torch.float32, device='cuda', generator=g)
omT = torch.randn((r, n), dtype=torch.float32, device='cuda', generator=g)
Y = torch.randn((m, r), dtype=torch.float32, device='cuda', generator=g)
ws = torch.empty(_lib.load().ggp_rsvd_tc_workspace_bytes(m), dtype=torch.uint8, device='cuda')
def ev(fn, reps=7):
    fn(); torch.cuda.synchronize(); best = 1e9
    for _ in range(reps):
        a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); best = min(best, a.elapsed_time(b))
    return best
res = {}
for tma in ('0', '1', '2', '3', '8', '9'):
    os.environ['GGP_TMA'] = tma
    ms = ev(lambda: ops.rsvd_sketch_tc(X, omT, ws))
    res['sketch_tma' + tma] = dict(ms=ms, gbs=4.0 * m * n / ms / 1e6, frac=4.0 * m * n / ms / 1e6 / 6533.8)
x1 = X.view(-1)
ms = ev(lambda: x1.sum())
res['torch_sum_read_only'] = dict(ms=ms, gbs=4.0 * m * n / ms / 1e6)
for tma in ('0', '1', '8', '9'):
    os.environ['GGP_TMA_XTY'] = tma
    ms = ev(lambda: ops.rsvd_xty_tc(X, Y))
    res['xty_tma' + tma] = dict(ms=ms, gbs=4.0 * m * n / ms / 1e6, frac=4.0 * m * n / ms / 1e6 / 6533.8)
ref = X.double()[:, :200000] @ omT.double()[:, :200000].T
for tma in ('0', '3'):
    os.environ['GGP_TMA'] = tma
    got = ops.rsvd_sketch_tc(X[:, :200000].contiguous(), omT[:, :200000].contiguous()).double()
    res['relerr_tma' + tma] = float((got - ref).abs().max() / ref.abs().max())
print(json.dumps(res))
