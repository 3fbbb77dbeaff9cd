"""Developer diagnostic: per-stage busy / barrier-wait cycles of the four warp roles of the look-ahead evaluation kernel, block 0
(needs the -DGGP_PHASES build of tools/build_phases.sh).   python tools/stage_timing_la.py B"""
import ctypes as C, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
so = os.path.join(ROOT, 'build_var', 'libggp_phases.so')
from gladsgp_b200 import _lib
_lib.LIB_PATH = so
from gladsgp_b200 import ops, synthetic
lib = _lib.load()
m, q = 512, 8
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1
d = q + 1
t = synthetic.design(m, q)
X = np.concatenate([0.5 * np.ones((m, 1)), t.astype(np.float64)], axis=1)
rng = np.random.default_rng(0)
beta = np.exp(rng.uniform(np.log(0.05), np.log(3.0), size=(B, d)))
lamz = rng.uniform(0.5, 2.0, B); dadd = rng.uniform(1e-3, 1e-2, B); W = rng.standard_normal((B, m))
Xd, Wd, bd, ld, dd = [torch.as_tensor(a, device='cuda') for a in (X, W, beta, lamz, dadd)]
ws = torch.empty((B, lib.ggp_factor_doubles(m)), dtype=torch.float64, device='cuda')
raw = C.CDLL(so)
buf = (C.c_ulonglong * 512)()
for _ in range(2):
    ops.loglik_batched(Xd, Wd, bd, ld, dd, factor_ws=ws)
raw.ggp_debug_stage_cycles(buf, 1)
ops.loglik_batched(Xd, Wd, bd, ld, dd, factor_ws=ws)
raw.ggp_debug_stage_cycles(buf, 0)
a = np.array(list(buf), dtype=np.float64).reshape(4, 64, 2)[:, :16, :]
print('B =', B, ' (k cycles; role 0 = factor warp, role 1 = inverse warp, roles 2, 3 = pool)')
print('stage   length |  idle before the end-of-stage barrier: role0  role1  role2  role3 | idle share')
tot_len = tot_wait = 0.0
for jc in range(16):
    busy = a[:, jc, 0]                          # cycles from the start of the stage to the role's arrival at barrier (E)
    length = busy.max()
    w = length - busy
    tot_len += length; tot_wait += w.sum()
    print('%5d %8.1f | %44.1f %6.1f %6.1f %6.1f | %5.1f%%' % (jc, length / 1e3, w[0] / 1e3, w[1] / 1e3, w[2] / 1e3, w[3] / 1e3,
                                                              100 * w.sum() / max(4 * length, 1)))
print('total %8.1f | idle share %.1f%%' % (tot_len / 1e3, 100 * tot_wait / max(4 * tot_len, 1)))
