"""bench.py's reference arm runs on the host CPU (oracle port), so its JSON contract can be checked without a GPU."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run_bench(*args, env=None):
    e = dict(os.environ)
    e.update(env or {})
    r = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), *args], capture_output=True, text=True, timeout=600, env=e)
    assert r.returncode == 0, r.stderr[-2000:]
    return r.stdout


def test_reference_arm_prints_the_contract_line():
    out = run_bench('--impl', 'reference', '--steps', '1', '--warmup', '0')
    lines = [l for l in out.strip().split('\n') if l.startswith('{')]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d['impl'] == 'reference' and d['metric'].startswith('rays/sec') and d['unit'] == 'rays/s'
    assert d['higher_is_better'] is True and d['vs_baseline'] is None and d['n_gpus'] == 1 and d['steps'] == 1
    assert d['value'] > 0 and d['ms_per_step'] > 0
    assert d['cpu_baseline']['kind'] == 'port' and d['cpu_baseline']['cores'] >= 1 and d['cpu_baseline']['value'] == d['value']
    assert d['e2e'] == {'value': d['value'], 'unit': 'rays/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}
    assert 'workload' in d['config'] and 'model' not in d['config']
    assert d['gpu_launches'] == 0


def test_reference_arm_other_ranks_exit_quietly():
    """Under torchrun only rank 0 measures the CPU arm; the other ranks print nothing and exit 0."""
    out = run_bench('--impl', 'reference', '--gpus', '2', '--steps', '1', '--warmup', '0', env={'RANK': '1', 'WORLD_SIZE': '2', 'LOCAL_RANK': '1'})
    assert out.strip() == ''
