"""The JSON line bench.py prints, held to the driver's contract: the committed record of the B200 arm (profiles/r2_bench_gen256.json, written on the
GPU box) carries every required key with a sane value, and the reference arm -- run live here on the host cores, one bounded step -- prints the same
metric / unit / config object, so the driver compares like with like."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _record():
    with open(os.path.join(ROOT, 'profiles', 'r2_bench_gen256.json')) as fh:
        return json.loads(fh.read().strip().splitlines()[-1])


def test_committed_b200_line_has_the_contract_keys():
    d = _record()
    for k in ('metric', 'value', 'unit', 'n_gpus', 'steps', 'warmup', 'ms_per_step', 'higher_is_better', 'scaling', 'vs_baseline', 'dtype', 'data', 'config',
              'e2e', 'gpu_launches', 'clocks', 'roofline', 'cpu_baseline'):
        assert k in d, k
    assert d['n_gpus'] == 1 and d['warmup'] >= 3 and d['higher_is_better'] is True and d['scaling'] == 'weak' and d['vs_baseline'] is None
    assert d['data'] == 'synthetic' and 'workload' in d['config'] and 'model' not in d['config']
    assert abs(d['value'] - d['config']['global_batch'] * 1e3 / d['ms_per_step']) < 1e-6 * d['value']
    e = d['e2e']
    assert e['h2d_bytes_per_step'] > 0 and e['d2h_bytes_per_step'] > 0 and 0 < e['value'] <= d['value'] * 1.02 and e['unit'] == d['unit']
    r = d['roofline']
    assert r['bound'] in ('hbm', 'tensor') and abs(r['frac'] - r['achieved'] / r['peak']) < 1e-9 and 0 < r['frac'] < 1 and 'traffic' in r
    c = d['cpu_baseline']
    assert c['kind'] in ('reference', 'port') and c['cores'] >= 1 and c['value'] > 0 and c['sample']
    assert d['gpu_launches'] == d['gpu_launches_per_step'] * d['steps'] > 0
    assert set(d['clocks']) >= {'sm_mhz', 'sm_max_mhz', 'reasons'}
    assert not {'hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown'} & set(d['clocks']['reasons'])


def test_reference_arm_prints_the_same_metric_and_config():
    r = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference', '--gpus', '1', '--steps', '1', '--warmup', '0'],
                       capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-1500:]
    ref = json.loads([ln for ln in r.stdout.splitlines() if ln.startswith('{')][-1])
    d = _record()
    assert ref['impl'] == 'reference' and ref['gpu_launches'] == 0
    for k in ('metric', 'unit', 'higher_is_better', 'scaling', 'config'):
        assert ref[k] == d[k], k
    assert ref['e2e'] == dict(value=ref['value'], unit=ref['unit'], h2d_bytes_per_step=0, d2h_bytes_per_step=0)
    assert ref['cpu_baseline']['value'] == ref['value'] and ref['cpu_baseline']['kind'] in ('reference', 'port')
    assert ref['value'] > 0 and ref['steps'] == 1
