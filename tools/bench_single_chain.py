"""Developer check: single-chain (and few-chain) steps/s at cfg3 shape with the speculative step kernel off / on.
   python tools/bench_single_chain.py            (prints one JSON line)"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

if len(sys.argv) > 1 and sys.argv[1] == '--child':
    import numpy as np
    import torch
    from gladsgp_b200 import _lib
    if os.environ.get('GGP_LIB'):                      # developer A/B runs against a variant library
        _lib.LIB_PATH = os.environ['GGP_LIB']
    from gladsgp_b200 import ops
    g = np.load(os.path.join(ROOT, 'tests', 'golden', 'chain_cfg3.npz'))
    tb = {k[3:]: g[k] for k in g.files if k.startswith('tb_')}
    P = tb['theta'].size
    res = {}
    for chains in (1, 2, 4):
        steps = 200
        us = np.random.RandomState(3).random_sample((chains, 2 * P * steps))
        eng = ops.McmcEngine(g['zt'], np.ascontiguousarray(g['w'].T), g['LamSim'], tb, n_chains=chains)
        eng.set_state(tb['theta'])
        eng.run(20, tb['step'], uniforms=us[:, :2 * P * 20])
        eng.set_state(tb['theta'])
        torch.cuda.synchronize()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); o = eng.run(steps, tb['step'], uniforms=us); e1.record(); torch.cuda.synchronize()
        res['chains_%d' % chains] = dict(steps_per_s_per_chain=steps / (e0.elapsed_time(e1) * 1e-3), lp_last=float(o['lp'][-1, 0].item()))
    print(json.dumps(res))
else:
    out = {}
    for spec in ('0', '1'):
        env = dict(os.environ, GGP_SPEC=spec)
        r = subprocess.run([sys.executable, os.path.abspath(__file__), '--child'], capture_output=True, text=True, env=env, timeout=600)
        try:
            out['GGP_SPEC=' + spec] = json.loads(r.stdout.strip().splitlines()[-1])
        except Exception:
            out['GGP_SPEC=' + spec] = {'failed': r.stderr[-1500:]}
    print(json.dumps(out))
