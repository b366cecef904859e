"""The committed bench lines under profiles/ are the evidence DESIGN.md quotes: check that they carry the contract's
keys, that the strong-scaling runs walked the same chain (equal ensemble checksums at 1, 2, 4 and 8 GPUs) and that the
reference arm was run on the product arm's configuration."""
import json
import os

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _line(name):
    path = os.path.join(ROOT, 'profiles', name)
    if not os.path.exists(path):
        pytest.skip(f'{name} not committed')
    return json.loads(open(path).read().strip().splitlines()[-1])


def test_scaling_lines_share_one_chain_and_the_contract_keys():
    lines = {n: _line(f'bench_r2_n{n}.json') for n in (1, 2, 4, 8)}
    assert len({d['ensemble_checksum'] for d in lines.values()}) == 1
    for n, d in lines.items():
        assert d['n_gpus'] == n and d['metric'] and d['unit'] and d['higher_is_better'] is True and d['scaling'] == 'strong'
        assert d['steps'] == lines[1]['steps'] and d['warmup'] >= 3 and d['dtype'] == 'f64' and d['data'] == 'synthetic'
        assert d['config'] == lines[1]['config'] and 'workload' in d['config']
        r = d['roofline']
        assert {'bound', 'achieved', 'peak', 'unit', 'frac', 'traffic'} <= set(r) and r['frac'] == pytest.approx(r['achieved'] / r['peak'])
        assert 0.0 < r['binding']['frac'] < 1.0
        assert {'value', 'unit', 'h2d_bytes_per_step', 'd2h_bytes_per_step'} <= set(d['e2e']) and d['e2e']['h2d_bytes_per_step'] > 0
        assert d['gpu_launches'] > 0
        # (lines saved before the sampler learnt to keep the loop going for short timed regions may hold no in-window sample)
        if d.get('clocks'):
            assert not set(d['clocks']['reasons']) & {'hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown'}
            assert d['clocks']['sm_mhz'] >= 0.9 * d['clocks']['sm_max_mhz']
        assert d['value'] == pytest.approx(65536 * d['steps'] / (d['ms_per_step'] * d['steps'] / 1e3), rel=1e-6)
    assert lines[1]['cpu_baseline']['kind'] == 'port' and lines[1]['cpu_baseline']['cores'] >= 1
    # more GPUs, more throughput (strong scaling of the fixed ensemble)
    assert lines[1]['value'] < lines[2]['value'] < lines[4]['value'] < lines[8]['value']


def test_reference_arm_ran_the_product_arms_configuration():
    ref, prod = _line('bench_r2_reference_arm.json'), _line('bench_r2_n1.json')
    assert ref['impl'] == 'reference' and ref['metric'] == prod['metric'] and ref['unit'] == prod['unit']
    ref_cfg = {k: v for k, v in ref['config'].items() if k != 'reference_step'}
    assert ref_cfg == prod['config']
    assert ref['e2e']['h2d_bytes_per_step'] == 0 and ref['e2e']['d2h_bytes_per_step'] == 0 and ref['e2e']['value'] == ref['value']
