"""CPU: the committed measurement artefacts belong to the committed sources.

bench.py reports `roofline.traffic` (DRAM bytes of the convolution kernels per step, from an ncu launch list) only while
profiles/roofline_traffic.json carries the sha of the convolution sources it was captured on; this test makes a source edit
without a re-capture visible before the GPU run does."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def test_roofline_traffic_matches_kernel_sources():
    import bench
    tj = json.load(open(os.path.join(ROOT, 'profiles', 'roofline_traffic.json')))
    assert tj.get('conv_source_sha') == bench.conv_source_sha(), \
        'profiles/roofline_traffic.json was captured on other convolution sources: re-run profiles/gpu_r2w2.sh and make_roofline_traffic.py'
    assert tj['traffic'] > 0 and tj['launches_per_step'] > 0


def test_bench_final_line_has_the_contract_keys():
    d = json.load(open(os.path.join(ROOT, 'profiles', 'r2_bench_final.json')))
    for k in ('metric', 'value', 'unit', 'n_gpus', 'steps', 'warmup', 'ms_per_step', 'higher_is_better', 'scaling', 'vs_baseline', 'dtype',
              'data', 'config', 'clocks', 'e2e', 'gpu_launches', 'roofline', 'cpu_baseline'):
        assert k in d, k
    r = d['roofline']
    for k in ('bound', 'achieved', 'peak', 'unit', 'frac', 'traffic'):
        assert k in r, k
    assert abs(r['frac'] - r['achieved'] / r['peak']) < 1e-9
    assert d['e2e']['h2d_bytes_per_step'] > 0 and d['gpu_launches'] > 0 and d['clocks']['samples'] >= 2
    assert d['cpu_baseline']['kind'] in ('port', 'reference') and d['cpu_baseline']['cores'] >= 1


def test_clock_sampler_cuts_samples_to_the_timed_region(tmp_path):
    """bench.ClockSampler: nvidia-smi lines carry a timestamp; only the samples inside [begin(), end()] count."""
    import datetime
    import time
    import bench

    class _Proc:
        def terminate(self): pass
        def wait(self, timeout=None): pass

    cs = bench.ClockSampler.__new__(bench.ClockSampler)
    cs.path, cs.proc = str(tmp_path / 'smi.csv'), _Proc()
    now = time.time()
    with open(cs.path, 'w') as f:
        for i in range(10):
            ts = datetime.datetime.fromtimestamp(now + 0.02 * i).strftime('%Y/%m/%d %H:%M:%S.%f')[:-3]
            f.write('%s, %d, 1965, 900.1, 0x4, Not Active, Not Active, Not Active, %s\n' % (ts, 1900 + i, 'Active' if i == 4 else 'Not Active'))
    cs.t0, cs.t1 = now + 0.05, now + 0.15
    out = cs.stop()
    assert out['samples'] == 5 and out['sm_mhz'] == 1905.0 and out['sm_max_mhz'] == 1965.0 and out['reasons'] == ['sw_power_cap']
