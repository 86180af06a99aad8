"""CPU-side checks of the C-ABI library: it loads, exports every symbol
include/dragnet_gpu.h declares, compiles/rejects plans, refuses to run without
a GPU (no CPU fallback), and the shard-merge helpers (pure host code) work."""

import ctypes
import os
import re
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(__file__))


@pytest.fixture(scope='module')
def native():
    import __graft_entry__
    __graft_entry__.build()
    from dragnet_b200 import native as n
    return n


def test_library_exports_every_declared_symbol(native):
    hdr = open(os.path.join(ROOT, 'include', 'dragnet_gpu.h')).read()
    hdr = re.sub(r'/\*.*?\*/', '', hdr, flags=re.S)
    declared = set(re.findall(r'\b(dng_[a-z0-9_]+)\s*\(', hdr))
    assert len(declared) >= 30
    L = native.lib()
    for name in sorted(declared):
        assert hasattr(L, name), name
    assert declared == set(native.SYMBOLS)


def test_plan_create_accepts_and_rejects(native):
    p = native.Plan('{"breakdowns":[{"name":"req.method","field":"req.method"}]}')
    p.close()
    for bad, frag in [('{', 'invalid plan'),
                      ('{"breakdowns":[],"filter":{"junk":["a","b"]}}',
                       'unknown operator "junk"'),
                      ('{"breakdowns":[{"name":"x","aggr":"lquantize"}]}',
                       'requires "step"'),
                      ('{"breakdowns":[],"format":"junk"}',
                       'unsupported format: "junk"')]:
        with pytest.raises(native.DngError) as ei:
            native.Plan(bad)
        assert frag in str(ei.value)


def test_no_cpu_fallback(native):
    if native.lib().dng_device_count() > 0:
        pytest.skip('a GPU is present')
    p = native.Plan('{"breakdowns":[]}')
    with pytest.raises(native.DngError) as ei:
        native.Scan(p)
    assert ei.value.code == -2


def test_generator_shape(native):
    import json
    g = native.gen_params(total_records=1000)
    data = native.gen_host(g, 0, 1000)
    lines = data.split(b'\n')
    assert lines[-1] == b'' and len(lines) == 1001
    o = json.loads(lines[0])
    assert list(o) == ['time', 'host', 'req', 'operation', 'res', 'latency',
                       'dataLatency', 'dataSize']
    assert o['time'] == '2014-05-31T21:00:00.000Z'
    assert 200 < len(data) / 1000.0 < 240
    # any sub-range reproduces the same bytes
    assert native.gen_host(g, 100, 50) == b'\n'.join(lines[100:150]) + b'\n'
