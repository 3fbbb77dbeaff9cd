"""Multi-GPU check of the PC-sharded sampler (torchrun, one rank per GPU, NCCL):
   python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 tools/check_by_pc.py
Every rank replays the golden cfg1 and cfg3 chains (tests/golden/chain_cfg*.npz) with its share of the PCs and must end with
the oracle's decisions and draws bit for bit; then the cfg3 chain is timed against the single-GPU step kernel."""
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from gladsgp_b200 import ops, dist as gdist  # noqa: E402


def main():
    rank, world, local = int(os.environ['RANK']), int(os.environ['WORLD_SIZE']), int(os.environ['LOCAL_RANK'])
    torch.cuda.set_device(local)
    dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    res = {'world': world}
    for cfg in ('cfg1', 'cfg3'):
        g = np.load(os.path.join(ROOT, 'tests', 'golden', 'chain_%s.npz' % cfg))
        tb = {k[3:]: g[k] for k in g.files if k.startswith('tb_')}
        replay = {k[3:]: g[k] for k in g.files if k.startswith('rp_')}
        n = int(g['n_steps'])
        eng = ops.McmcEngine(g['zt'], np.ascontiguousarray(g['w'].T), g['LamSim'], tb, n_chains=1)
        eng.set_state(tb['theta'])
        out = gdist.mcmc_by_pc(eng, n, tb['step'], replay=replay, record_accept=True)
        ok = bool(np.array_equal(out['accepted'].cpu().numpy(), g['chain_acc']) and
                  np.array_equal(out['draws'].cpu().numpy()[:, 0, :], g['chain_draws']))
        t = torch.tensor([1.0 if ok else 0.0], device='cuda')
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
        res[cfg + '_bit_identical_on_all_ranks'] = bool(t.item() == 1.0)
        res[cfg + '_collectives'] = out['collective_calls']
        if cfg == 'cfg3':
            P = tb['theta'].size
            steps = 60
            us = np.random.RandomState(7).random_sample((1, 2 * P * steps))
            for name in ('one_gpu_step_kernel', 'by_pc'):
                eng.set_state(tb['theta'])
                if name == 'by_pc':
                    gdist.mcmc_by_pc(eng, 5, tb['step'], uniforms=us[:, :2 * P * 5])
                else:
                    eng.run(5, tb['step'], uniforms=us[:, :2 * P * 5])
                eng.set_state(tb['theta'])
                torch.cuda.synchronize(); dist.barrier()
                e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
                e0.record()
                o = gdist.mcmc_by_pc(eng, steps, tb['step'], uniforms=us) if name == 'by_pc' else eng.run(steps, tb['step'], uniforms=us)
                e1.record(); torch.cuda.synchronize()
                ms = torch.tensor([e0.elapsed_time(e1)], device='cuda')
                dist.all_reduce(ms, op=dist.ReduceOp.MAX)
                res['cfg3_single_chain_steps_per_s_' + name] = steps / (float(ms.item()) * 1e-3)
                res['cfg3_lp_last_' + name] = float(o['lp'][-1, 0].item())
    if rank == 0:
        print(json.dumps(res))
    dist.barrier()
    dist.destroy_process_group()


if __name__ == '__main__':
    main()
