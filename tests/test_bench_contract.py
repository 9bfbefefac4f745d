"""bench.py on a machine without a GPU: the reference arm prints the contract's JSON line (CPU work only), the B200 arm
refuses to run instead of falling back to anything."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(*args):
    env = dict(os.environ, CUDA_VISIBLE_DEVICES='')
    return subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py')] + list(args), capture_output=True, text=True, env=env,
                          cwd=ROOT, timeout=600)


def test_reference_arm_prints_the_contract_line():
    r = _run('--impl', 'reference', '--steps', '1', '--warmup', '0', '--ref-windows', '2')
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith('{')]
    assert len(lines) == 1, r.stdout[-2000:]
    d = json.loads(lines[0])
    assert d['impl'] == 'reference' and d['unit'] == 'frames/s' and d['higher_is_better'] is True
    assert d['metric'] == 'frames/sec LGD-RNN-12 N=4 ws=32' and d['value'] > 0 and d['steps'] == 1
    assert d['cpu_baseline']['kind'] in ('reference', 'port') and d['cpu_baseline']['cores'] >= 1
    assert d['cpu_baseline']['value'] == d['value'] == d['e2e']['value']
    assert d['e2e']['h2d_bytes_per_step'] == 0 and d['e2e']['d2h_bytes_per_step'] == 0
    assert d['config']['windows_per_step'] == 2 and d['vs_baseline'] is None


def test_b200_arm_has_no_cpu_fallback():
    r = _run('--steps', '1', '--warmup', '1')
    assert r.returncode != 0
    assert 'no CPU fallback' in (r.stdout + r.stderr)
    assert not [l for l in r.stdout.splitlines() if l.startswith('{')]
