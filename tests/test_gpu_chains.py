"""Chain identity at the BASELINE.json configuration sizes: the device sampler replays the oracle's recorded proposals and
uniforms (tests/golden/chain_cfg*.npz, made by tests/golden/make_golden_chains.py) and must reproduce every accept
decision and every draw bit for bit -- cfg1 m=100/q=8/pu=5 x 200 steps, cfg2 m=1000 scalar x 50 steps, cfg3
m=512/q=8/pu=10 x 30 steps; 13 576 evaluated decisions in all.  Each fixture is run under the library's own schedule
for a single chain (a thread-block cluster per matrix) and with one CTA per matrix (the look-ahead schedule of the bench).
"""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')


def _replay(g, n_chains=1):
    from gladsgp_b200 import ops
    tb = {k[3:]: g[k] for k in g.files if k.startswith('tb_')}
    eng = ops.McmcEngine(g['zt'], np.ascontiguousarray(g['w'].T), g['LamSim'], tb, n_chains=n_chains)
    eng.set_state(tb['theta'])
    replay = {k[3:]: np.repeat(g[k], n_chains, axis=1) for k in g.files if k.startswith('rp_')}
    n = int(g['n_steps'])
    return eng.run(n, tb['step'], replay=replay, record_accept=True)


@pytest.mark.parametrize('cluster', [None, '1'])
@pytest.mark.parametrize('cfg', ['cfg1', 'cfg2', 'cfg3'])
def test_chain_identical_to_oracle_at_config_sizes(cuda, monkeypatch, cfg, cluster):
    if cluster is not None:
        monkeypatch.setenv('GGP_CLUSTER', cluster)
    g = np.load(os.path.join(GOLD, 'chain_%s.npz' % cfg))
    out = _replay(g)
    acc = out['accepted'].cpu().numpy()
    assert np.array_equal(acc, g['chain_acc'])                                  # every decision
    assert np.array_equal(out['draws'].cpu().numpy()[:, 0, :], g['chain_draws'])  # bit-identical chain
    np.testing.assert_allclose(out['lp'].cpu().numpy()[:, 0], g['chain_lp'], rtol=1e-9)


def test_many_replicas_of_one_chain_are_identical(cuda):
    """74 replicas of the cfg3 chain (740 matrices: more than one wave of CTAs, look-ahead schedule, CTAs of one chain
    finishing in any order): every replica reproduces the oracle's chain -- the close of a step by whichever CTA arrives
    last is order-independent."""
    g = np.load(os.path.join(GOLD, 'chain_cfg3.npz'))
    steps = 6
    gg = {k: g[k] for k in g.files}
    for k in list(gg):
        if k.startswith('rp_'):
            gg[k] = gg[k][:steps]
    gg['n_steps'] = np.array(steps)

    class G(dict):
        files = list(gg.keys())
    out = _replay(G(gg), n_chains=74)
    draws = out['draws'].cpu().numpy()
    for c in range(74):
        assert np.array_equal(draws[:, c, :], g['chain_draws'][:steps])
    assert np.array_equal(out['accepted'].cpu().numpy()[:, 0, :], g['chain_acc'][:steps, 0, :])


@pytest.mark.parametrize('spec', ['0', '1'])
@pytest.mark.parametrize('splits', [[(0, 5)], [(0, 3), (3, 2)], [(0, 2), (2, 2), (4, 1)]])
def test_pc_sharded_stepping_is_bit_identical(cuda, monkeypatch, splits, spec):
    """SURVEY 8e ii: the PCs of a chain swept by several shards (here: one after the other on one GPU, standing in for
    the ranks), rows exchanged, step closed from the gathered rows -- same decisions and draws as the one-kernel step."""
    from gladsgp_b200 import ops
    monkeypatch.setenv('GGP_SPEC', spec)            # shards of a single chain run the speculative step kernel when it fits
    g = np.load(os.path.join(GOLD, 'chain_cfg1.npz'))
    tb = {k[3:]: g[k] for k in g.files if k.startswith('tb_')}
    steps = 40
    replay = {k[3:]: g[k][:steps] for k in g.files if k.startswith('rp_')}
    eng = ops.McmcEngine(g['zt'], np.ascontiguousarray(g['w'].T), g['LamSim'], tb, n_chains=1)
    eng.set_state(tb['theta'])
    out = eng.run_by_pc(steps, tb['step'], replay=replay, record_accept=True, shards=splits)
    assert np.array_equal(out['accepted'].cpu().numpy(), g['chain_acc'][:steps])
    assert np.array_equal(out['draws'].cpu().numpy()[:, 0, :], g['chain_draws'][:steps])
    np.testing.assert_allclose(out['lp'].cpu().numpy()[:, 0], g['chain_lp'][:steps], rtol=1e-9)


def test_pc_sharded_stepping_follows_the_uniform_stream(cuda):
    """Stream mode, several chains: candidates drawn on the device in the close of the previous step."""
    from gladsgp_b200 import ops
    g = np.load(os.path.join(GOLD, 'chain_cfg1.npz'))
    tb = {k[3:]: g[k] for k in g.files if k.startswith('tb_')}
    P = tb['theta'].size
    steps, chains = 12, 3
    us = np.random.RandomState(5).random_sample((chains, 2 * P * steps))
    outs = []
    for mode in ('fused', 'sharded'):
        eng = ops.McmcEngine(g['zt'], np.ascontiguousarray(g['w'].T), g['LamSim'], tb, n_chains=chains)
        eng.set_state(tb['theta'])
        if mode == 'fused':
            o = eng.run(steps, tb['step'], uniforms=us, record_accept=True)
        else:
            o = eng.run_by_pc(steps, tb['step'], uniforms=us, record_accept=True, shards=[(0, 2), (2, 3)])
        outs.append({k: o[k].cpu().numpy() for k in ('draws', 'lp', 'accepted', 'consumed')})
    for k in ('draws', 'lp', 'accepted', 'consumed'):
        assert np.array_equal(outs[0][k], outs[1][k]), k


@pytest.mark.parametrize('spec', ['0', '1'])
@pytest.mark.parametrize('cfg,chains', [('cfg1', 1), ('cfg1', 2), ('cfg2', 1), ('cfg3', 1)])
def test_speculative_step_kernel_is_bit_identical(cuda, monkeypatch, cfg, chains, spec):
    """Few chains: three clusters per (PC, chain) retire two evaluations per round (the second one under both outcomes of
    the first).  Decisions and draws must be those of the sequential sweep, with the kernel switched off (GGP_SPEC=0) and
    on; in stream mode the two must also agree with each other bit for bit."""
    from gladsgp_b200 import ops
    monkeypatch.setenv('GGP_SPEC', spec)
    g = np.load(os.path.join(GOLD, 'chain_%s.npz' % cfg))
    n = min(int(g['n_steps']), 60)
    gg = {k: (g[k][:n] if k.startswith('rp_') else g[k]) for k in g.files}
    gg['n_steps'] = np.array(n)

    class G(dict):
        files = list(gg.keys())
    out = _replay(G(gg), n_chains=chains)
    for c in range(chains):
        assert np.array_equal(out['draws'].cpu().numpy()[:, c, :], g['chain_draws'][:n])
    assert np.array_equal(out['accepted'].cpu().numpy()[:, 0, :], g['chain_acc'][:n, 0, :])
    np.testing.assert_allclose(out['lp'].cpu().numpy()[:, 0], g['chain_lp'][:n], rtol=1e-9)
    # stream mode from the same start: digest of the chain is independent of the kernel choice
    tb = {k[3:]: g[k] for k in g.files if k.startswith('tb_')}
    P = tb['theta'].size
    us = np.random.RandomState(9).random_sample((chains, 2 * P * 25))
    eng = ops.McmcEngine(g['zt'], np.ascontiguousarray(g['w'].T), g['LamSim'], tb, n_chains=chains)
    eng.set_state(tb['theta'])
    o = eng.run(25, tb['step'], uniforms=us, record_accept=True)
    key = (cfg, chains)
    cur = (o['draws'].cpu().numpy().tobytes(), o['lp'].cpu().numpy().tobytes(), o['consumed'].cpu().numpy().tobytes())
    prev = _STREAM_RESULTS.setdefault(key, cur)
    assert prev == cur


_STREAM_RESULTS = {}
