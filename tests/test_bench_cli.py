"""bench.py contract checks that need no GPU: the reference arm prints exactly one JSON line with
the keys the driver reads, and the B200 arm refuses to run without a CUDA device (no CPU path)."""

import json
import os
import subprocess
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_json_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference', '--steps', '1',
                        '--warmup', '3', '--basins', '4'], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, r.stdout
    d = json.loads(lines[0])
    assert d['impl'] == 'reference' and d['unit'] == 'basin-timesteps/s' and d['higher_is_better'] is True
    assert d['metric'].startswith('basin-timesteps/sec') and d['dtype'] == 'f32' and d['vs_baseline'] is None
    # the same `config` as the B200 arm prints (same workload, full size, every step); what ran on
    # the CPU — the unmodified reference from baseline/_ref, else the oracle port — is in cpu_baseline
    assert 'workload' in d['config'] and d['config']['basins_per_gpu'] == 4 and 'sample' in d['cpu_baseline']
    have_ref = os.path.isdir(os.path.join(ROOT, 'baseline', '_ref', 'hydrodl2'))
    assert d['cpu_baseline']['kind'] == ('reference' if have_ref else 'port') and d['cpu_baseline']['cores'] >= 1
    assert d['cpu_baseline']['value'] == d['value'] == d['e2e']['value'] > 0
    assert d['e2e']['h2d_bytes_per_step'] == 0 and d['e2e']['d2h_bytes_per_step'] == 0


def test_b200_arm_needs_cuda():
    if torch.cuda.is_available():
        return
    r = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--steps', '1'],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode != 0 and r.stdout.strip() == ''
    assert 'no CPU fallback' in r.stderr
