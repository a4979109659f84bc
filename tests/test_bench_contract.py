"""bench.py's reference arm runs on the host CPU (the reference's own modules from /root/reference or oracle/_ref, else the oracle
port), so its JSON contract can be checked without a GPU; so can the oracle/_ref recipe."""
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
    sys.path.insert(0, os.path.join(ROOT, 'oracle'))
    import ref_import
    assert d['cpu_baseline']['kind'] == ('reference' if ref_import.usable() else 'port')
    assert d['cpu_baseline']['cores'] >= 1 and d['cpu_baseline']['value'] == d['value'] and d['device'] == 'cpu'
    assert d['e2e'] == {'value': d['value'], 'unit': 'rays/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}
    assert 'workload' in d['config'] and 'model' not in d['config']
    assert d['gpu_launches'] == 0


def test_reference_arm_other_ranks_exit_quietly():
    """Under torchrun only rank 0 measures the CPU arm; the other ranks print nothing and exit 0."""
    out = run_bench('--impl', 'reference', '--gpus', '2', '--steps', '1', '--warmup', '0', env={'RANK': '1', 'WORLD_SIZE': '2', 'LOCAL_RANK': '1'})
    assert out.strip() == ''


def test_reference_bytecode_build_is_importable_without_the_tree():
    """oracle/build_ref.py: the render-path modules of the reference, byte-compiled where they lie into oracle/_ref (no source copied);
    the build imports sourceless -- what the GPU box, which has no /root/reference, relies on -- and renders what the oracle renders."""
    sys.path.insert(0, os.path.join(ROOT, 'oracle'))
    import pytest
    import ref_import
    if not ref_import.available() and not ref_import.built():
        pytest.skip('neither the reference tree nor its bytecode build is present')
    if ref_import.available():
        import build_ref
        assert build_ref.build() is not None
    assert ref_import.built()
    files = sorted(os.listdir(os.path.join(ROOT, 'oracle', '_ref', 'utils')))
    assert all(f.endswith('.pyc') or f == 'BUILT_FROM' for f in files), files          # bytecode only, never a source file
    code = ("import sys; sys.path.insert(0, %r); import ref_import; ref_import.REF_ROOT = '/nonexistent'; RN, RH = ref_import.load(); "
            "import torch; assert RN.__file__.endswith('.pyc'); print(float(RH.get_embedder(10, 0)[1]))" % os.path.join(ROOT, 'oracle'))
    r = subprocess.run([sys.executable, '-c', code], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and r.stdout.strip() == '63.0', r.stderr[-2000:]
