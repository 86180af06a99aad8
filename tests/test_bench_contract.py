"""bench.py's reference arm runs without a GPU: check that it prints one JSON
line with the keys the driver reads (the GPU arm prints the same keys plus
`roofline`; it cannot run here).  Also the clocks sampler's parsing."""

import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def test_reference_arm_json_line():
    out = subprocess.run(
        [sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference',
         '--steps', '1', '--warmup', '1', '--cpu-rows', '20000',
         '--rows', '20000'],
        capture_output=True, check=True, text=True, timeout=600).stdout
    lines = [l for l in out.splitlines() if l.startswith('{')]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for k in ('metric', 'value', 'unit', 'n_gpus', 'steps', 'warmup',
              'ms_per_step', 'higher_is_better', 'scaling', 'vs_baseline',
              'dtype', 'data', 'config', 'cpu_baseline', 'e2e', 'impl'):
        assert k in d, k
    assert d['impl'] == 'reference' and d['metric'] == 'json_records_per_sec'
    assert d['value'] > 0 and d['cpu_baseline']['kind'] in ('port', 'reference')
    assert d['e2e']['h2d_bytes_per_step'] == 0
    assert 'workload' in d['config']


def test_clock_sampler_summarises_what_it_saw():
    import bench
    s = bench.ClockSampler([0, 1])
    s.proc = object()            # pretend nvidia-smi is running
    s.lines = [(1.0, '0, 1965, 1965, 400.1, Not Active, Not Active, Not Active, Not Active'),
               (1.0, '1, 1950, 1965, 410.0, Not Active, Not Active, Not Active, Active'),
               (5.0, '0, 1800, 1965, 420.0, Not Active, Not Active, Not Active, Not Active'),
               (5.0, '1, 1965, 1965, 415.0, Not Active, Not Active, Active, Not Active')]
    s.t0 = 4.0
    r = s.since_mark()
    assert r['samples'] == 2 and r['sm_max_mhz'] == 1965.0
    assert r['reasons'] == ['sw_thermal_slowdown']
    s.t0 = 9.0                   # nothing since the mark: the latest lines
    r = s.since_mark()
    assert r['samples'] == 2 and r['sm_mhz'] in (1800.0, 1965.0)
