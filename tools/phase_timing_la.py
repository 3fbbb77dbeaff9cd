"""Developer diagnostic: per-warp activity cycles of the look-ahead evaluation kernel, block 0 (needs the -DGGP_PHASES
build of tools/build_phases.sh).   python tools/phase_timing_la.py B"""
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
buf = (C.c_ulonglong * 128)()
for _ in range(2):
    ops.loglik_batched(Xd, Wd, bd, ld, dd, factor_ws=ws)
raw.ggp_debug_phase2_cycles(buf, 1)
ops.loglik_batched(Xd, Wd, bd, ld, dd, factor_ws=ws)
raw.ggp_debug_phase2_cycles(buf, 0)
names = {0: 'stage set-up', 1: 'barrier T', 2: 'task fetch / misc', 3: 'named barriers A/B/C', 4: 'inverse (w1)', 5: 'GEMM', 6: 'cov final (mask, P=C-S)',
         7: 'D write', 8: 'cov distances (DMMA)', 9: 'cov exponentials', 10: 'factor (w0)', 11: 'w solve + diag rows (w0)', 12: 'TRSM + store', 13: 'barrier E'}
print('B =', B)
for w in range(4):
    tot = sum(buf[16 * w + i] for i in range(16))
    print('warp %d: total %d cycles' % (w, tot))
    for i in range(14):
        v = buf[16 * w + i]
        if v:
            print('   %-28s %10d %5.1f%%' % (names[i], v, 100.0 * v / max(tot, 1)))
