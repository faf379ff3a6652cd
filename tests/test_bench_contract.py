"""The bench.py output contract: the reference arm is run here on the CPU (it only needs the oracle), the product arm is checked
on the bench lines committed under profiles/ (they were produced on a B200 by the same script)."""
import glob
import json
import os
import subprocess
import sys

from conftest import ROOT

BASE_KEYS = {'metric', 'value', 'unit', 'n_gpus', 'steps', 'warmup', 'ms_per_step', 'higher_is_better', 'scaling', 'vs_baseline', 'dtype',
             'data', 'config', 'e2e'}


def test_reference_arm_prints_one_contract_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference', '--steps', '1', '--warmup', '1'],
                         capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith('{')]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d['impl'] == 'reference' and BASE_KEYS <= set(d) and d['metric'] == 'nav-step decisions/sec' and d['unit'] == 'decisions/s'
    assert d['value'] > 0 and d['higher_is_better'] is True and d['vs_baseline'] is None
    assert d['e2e'] == {'value': d['value'], 'unit': d['unit'], 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}
    cb = d['cpu_baseline']
    assert cb['kind'] in ('reference', 'port') and cb['cores'] >= 1 and cb['value'] == d['value'] and 'sample' in cb
    assert 'workload' in d['config'] and 'model' not in d['config']


def test_committed_product_bench_lines_follow_the_contract():
    files = sorted(glob.glob(os.path.join(ROOT, 'profiles', 'r01b_bench_*.json')))
    assert files
    for f in files:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        assert BASE_KEYS | {'clocks', 'gpu_launches', 'roofline'} <= set(d), f
        assert d['gpu_launches'] > 0 and d['value'] > 0 and d['dtype'] in ('bf16', 'fp32')
        r = d['roofline']
        assert {'bound', 'achieved', 'peak', 'unit', 'frac', 'traffic'} <= set(r) and abs(r['frac'] - r['achieved'] / r['peak']) < 1e-9
        e = d['e2e']
        assert {'value', 'unit', 'h2d_bytes_per_step', 'd2h_bytes_per_step'} <= set(e) and e['value'] != d['value']
        assert not set(d['clocks']['reasons']) & {'hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown'}
        if d['n_gpus'] == 1 and 'train' not in f:
            assert {'value', 'unit', 'cores', 'kind', 'sample'} <= set(d['cpu_baseline'])
