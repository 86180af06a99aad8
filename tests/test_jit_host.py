"""The run-time compiled matcher (dragnet_b200/csrc/jit.cpp) without a GPU:
the source generated for a scan's templates must build for the device (NVRTC)
and link against the relocatable kernel embedded in the library (nvJitLink),
and the same generated code, compiled for the host by tests/hostcheck, must
reproduce the oracle (the 'jit' mode of test_hostcheck_*.py covers the goldens
and the edge corpus; this file covers the device build and the synthetic
shape)."""

import json
import os
import subprocess
import sys
import tempfile

sys.path.insert(0, os.path.dirname(__file__))
import corpus  # noqa: E402
from engines import build_hostcheck, canon_points, py_engine  # noqa: E402
from engines import hostcheck_engine  # noqa: E402


def test_generated_matcher_builds_for_the_device_and_matches(tmp_path,
                                                             monkeypatch):
    from dragnet_b200 import native
    n = 3000
    data = native.gen_host(native.gen_params(total_records=n), 0, n)
    p = tmp_path / 'syn.log'
    p.write_bytes(data)
    plan = corpus.make_plan(['-b', 'req.method,res.statusCode', '-f',
                             '{"eq":["req.method","GET"]}'])
    exe = build_hostcheck()
    pf = tmp_path / 'plan.json'
    pf.write_text(json.dumps(plan))
    env = dict(os.environ, DNG_HOSTCHECK_F='1', DNG_HOSTCHECK_JIT='2')
    r = subprocess.run([exe, str(pf), str(p)], capture_output=True, env=env)
    assert r.returncode == 0, r.stderr.decode()
    assert b'nvrtc' in r.stderr and b'link' in r.stderr, r.stderr
    doc = json.loads(r.stdout)
    assert doc['nfmatch'] == n and doc['nfmiss'] == 0, doc
    # and the results are the oracle's
    monkeypatch.setenv('DNG_HOSTCHECK_F', '1')
    monkeypatch.setenv('DNG_HOSTCHECK_JIT', '1')
    exp_p, exp_c = py_engine(plan, [str(p)])
    act_p, act_c = hostcheck_engine(plan, [str(p)])
    assert canon_points(act_p) == canon_points(exp_p)
    assert act_c == exp_c


def test_three_column_plans_link_with_shared_scanners(tmp_path):
    """Plans of three columns and more get the wildcard scanners as shared,
    called copies (jit.cpp: their loop does not fit the instruction cache
    otherwise); that source must build and link for the device, and the linked
    code must really hold the two functions."""
    from dragnet_b200 import native
    n = 2000
    data = native.gen_host(native.gen_params(total_records=n), 0, n)
    p = tmp_path / 'syn.log'
    p.write_bytes(data)
    exe = build_hostcheck()
    for argv, shared in ((['-b', 'operation,req.method,host'], True),
                         (['-b', 'req.method,host'], False)):
        plan = corpus.make_plan(argv)
        pf = tmp_path / 'plan.json'
        pf.write_text(json.dumps(plan))
        dump = str(tmp_path / 'jm')
        env = dict(os.environ, DNG_HOSTCHECK_F='1', DNG_HOSTCHECK_JIT='2',
                   DNG_HOSTCHECK_JIT_DUMP=dump)
        env.pop('DNG_JIT_SHARED', None)
        r = subprocess.run([exe, str(pf), str(p)], capture_output=True,
                           env=env)
        assert r.returncode == 0, r.stderr.decode()
        src = open(dump + '.cu').read()
        assert ('#define DNG_JIT_SHARED_SCAN' in src) == shared
        sass = subprocess.run(['cuobjdump', '-sass', dump + '.cubin'],
                              capture_output=True, text=True).stdout
        fns = [l for l in sass.splitlines() if 'Function :' in l]
        assert any('fscan_str_p' in l for l in fns) == shared
        assert any('fscan_bare_p' in l for l in fns) == shared
