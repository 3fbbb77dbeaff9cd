"""CPU: the reference arm of bench.py (`--impl reference`, the oracle port timed on the host cores) prints one JSON line with
the contract's keys; on rank > 0 it prints nothing."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(env_extra):
    env = dict(os.environ, **env_extra)
    return subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference', '--steps', '1', '--warmup', '1',
                           '--ref-nx', '40', '--ref-nt', '10'], capture_output=True, text=True, env=env, timeout=600)


def test_reference_arm_prints_contract_line():
    r = _run({'RANK': '0'})
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith('{')]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d['impl'] == 'reference' and d['metric'] == 'mcmc_steps_per_s' and d['unit'] == 'chain-steps/s'
    assert d['higher_is_better'] is True and d['value'] > 0 and d['steps'] == 1 and d['warmup'] == 1
    assert d['cpu_baseline']['kind'] == 'port' and d['cpu_baseline']['cores'] >= 1 and d['cpu_baseline']['value'] == d['value']
    assert d['e2e'] == {'value': d['value'], 'unit': d['unit'], 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}
    assert 'workload' in d['config'] and d['dtype'] == 'f64'
    # both host modes are reported (one chain x all BLAS threads; one single-thread chain per core), value = the better one
    modes = d['cpu_baseline']['modes']
    assert set(modes) == {'one_chain', 'per_core_chains'}
    assert modes['one_chain']['chains'] == 1 and modes['per_core_chains']['blas_threads'] == 1
    assert d['value'] == max(v['value'] for v in modes.values())


def test_reference_arm_is_silent_on_other_ranks():
    r = _run({'RANK': '1', 'WORLD_SIZE': '2'})
    assert r.returncode == 0 and r.stdout.strip() == ''
