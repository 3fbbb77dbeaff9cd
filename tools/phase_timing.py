"""Developer diagnostic: per-phase cycles of the fused evaluation kernel (needs a -DGGP_PHASES build).
   python tools/phase_timing.py B"""
import ctypes as C, os, subprocess, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
so = os.path.join(ROOT, 'build_var', 'libggp_phases.so')
if not os.path.exists(so):
    raise SystemExit('build first: nvcc -DGGP_PHASES ... (see tools/build_phases.sh)')
from gladsgp_b200 import _lib
_lib.LIB_PATH = so
from gladsgp_b200 import ops, synthetic
lib = _lib.load()
m, q = int(os.environ.get('M', 512)), int(os.environ.get('Q', 8))          # M=4096 Q=16 with B=20: the cfg5 shape (16-CTA clusters)
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
buf = (C.c_ulonglong * 32)()
for _ in range(2):
    ops.loglik_batched(Xd, Wd, bd, ld, dd, factor_ws=ws)
raw.ggp_debug_phase_cycles(buf, 1)
ops.loglik_batched(Xd, Wd, bd, ld, dd, factor_ws=ws)
raw.ggp_debug_phase_cycles(buf, 1)
names = ['fill+sync', 'pair0 gemm+cov', 'wait A', 'factor(w0)/idle', 'wait B0', 'usolve(w0)/minv/idle', 'wait B', 'finish+rest', 'wait C']
nP = (m + 31) // 32
print('precise (data-dependent) serial timers, cycles per panel (%d panels):' % nP)
for slot, n in ((9, 'w0: other'), (15, 'w0: factor blk0'), (10, 'w0: factor blk1-3'), (11, 'w0: B0 barrier'), (12, 'w0: usolve+store'), (13, 'w1: other'), (14, 'w1: Minv')):
    print('   %-20s %9.0f' % (n, buf[slot] / float(nP)))
tot2 = sum(buf[i] for i in (20, 21, 22, 23))
print('worker warp 2 of block 0: total', tot2)
for slot, n in ((20, 'DMMA update'), (21, 'covariance'), (22, 'TRSM+store'), (23, 'other (barriers, serial wait)')):
    print('   %-30s %10d %5.1f%%' % (n, buf[slot], 100.0 * buf[slot] / max(tot2, 1)))
for w, off in (('warp0', 0),):
    tot = sum(buf[off + i] for i in range(9))
    print(w, 'total cycles', tot)
    for i, n in enumerate(names):
        print('   %-24s %9d  %5.1f%%' % (n, buf[off + i], 100.0 * buf[off + i] / max(tot, 1)))
