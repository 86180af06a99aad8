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
