"""The bench line committed under profiles/ (the last `python bench.py` on a B200 with the build in the tree) carries
every key the measurement contract asks for, and its derived figures are consistent with each other.  CPU-only."""
import json
import os

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _line(name):
    p = os.path.join(ROOT, 'profiles', name)
    if not os.path.exists(p):
        pytest.skip(name + ' not present')
    return json.loads(open(p).read().strip().splitlines()[-1])


@pytest.mark.parametrize('name,n', [('bench_r2g_n1.json', 1), ('bench_r2g_n2.json', 2)])
def test_committed_bench_line_keeps_the_contract(name, n):
    d = _line(name)
    for k in ('metric', 'value', 'unit', 'n_gpus', 'steps', 'warmup', 'ms_per_step', 'higher_is_better', 'scaling',
              'vs_baseline', 'dtype', 'data', 'config', 'e2e', 'gpu_launches', 'clocks', 'roofline'):
        assert k in d, k
    assert d['metric'] == 'lm_iterations_per_s' and d['n_gpus'] == n and d['dtype'] == 'f64'
    assert d['higher_is_better'] is True and d['scaling'] == 'weak' and d['vs_baseline'] is None
    assert d['warmup'] >= 3 and d['data'] == 'synthetic' and 'workload' in d['config'] and 'model' not in d['config']
    assert d['gpu_launches'] > 0
    # value = 2M-observation equivalents per second of the device-timed step
    nobs = 2.0e6 if n == 1 else 5.0e6 * n
    assert d['value'] == pytest.approx(1e3 / d['ms_per_step'] * nobs / 2.0e6, rel=1e-9)
    e = d['e2e']
    assert e['unit'] == d['unit'] and e['h2d_bytes_per_step'] > 0 and e['d2h_bytes_per_step'] > 0
    assert 0 < e['value'] < d['value']                  # host buffers in the timed region cost something
    r = d['roofline']
    for k in ('bound', 'achieved', 'peak', 'unit', 'frac', 'traffic'):
        assert k in r, k
    assert r['bound'] in ('hbm', 'tensor') and r['frac'] == pytest.approx(r['achieved'] / r['peak'], rel=1e-12)
    c = d['clocks']
    assert c['reasons'] == [] and c['sm_mhz'] >= 0.9 * c['sm_max_mhz']
    # the phases the rooflines are computed from add up to the step
    ph = d['phases_ms_last_step']
    assert sum(v for k, v in ph.items() if k != 'total') == pytest.approx(ph['total'], rel=0.05)
    if n == 1:
        b = d['cpu_baseline']
        for k in ('value', 'unit', 'cores', 'kind', 'sample'):
            assert k in b, k
        assert b['kind'] == 'port' and b['unit'] == d['unit'] and 0 < b['value'] < d['value']
        # per-phase rooflines: the streaming phases are quoted against an HBM peak (MEASURED_PEAKS.json of that run)
        for k, v in d['rooflines'].items():
            if v['bound'] == 'hbm':
                assert v['unit'] == 'GB/s' and v['peak'] > 1000 and 0 < v['frac'] < 1
                assert v['achieved'] == pytest.approx(v['bytes_per_obs'] * 2.0e6 / (v['ms'] * 1e-3) / 1e9, rel=1e-9)
