"""The host build of the per-record device logic and of the template builder
(tests/hostcheck, TEST ONLY) under AddressSanitizer + UBSan, over the edge,
json-skinner, scalar-form, fuzz and random corpora in template mode: memory
errors and undefined behaviour in code that also runs on the device."""

import json
import os
import subprocess
import sys

import pytest

sys.path.insert(0, os.path.dirname(__file__))
import corpus  # noqa: E402

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, 'dragnet_b200', 'csrc')
HC = os.path.join(ROOT, 'tests', 'hostcheck')


@pytest.fixture(scope='module')
def asan_exe(tmp_path_factory):
    exe = str(tmp_path_factory.mktemp('asan') / 'hostcheck_asan')
    subprocess.check_call(['make', '-s', '-C', CSRC, 'build/jit_blob.o'])
    subprocess.check_call(
        ['g++', '-std=c++17', '-O1', '-g', '-fsanitize=address,undefined',
         '-fno-omit-frame-pointer', '-I/usr/local/cuda/include', '-o', exe,
         os.path.join(HC, 'hostcheck.cpp'), os.path.join(CSRC, 'plan.cpp'),
         os.path.join(CSRC, 'result.cpp'), os.path.join(CSRC, 'tmpl.cpp'),
         os.path.join(CSRC, 'fast.cpp'), os.path.join(CSRC, 'jit.cpp'),
         os.path.join(CSRC, 'build', 'jit_blob.o'), '-ldl'])
    return exe


def test_no_memory_errors_or_ub(asan_exe, tmp_path):
    files = {
        'edge': corpus.EDGE_LINES,
        'skinner': corpus.SKINNER_LINES,
        'scalars': [b'{"a":7,"s":"k0"}'] * 3 + corpus.scalar_lines(),
        'fuzz': corpus.template_fuzz_lines(3, 500),
        'random': corpus.random_lines(2, 300),
    }
    jobs = [('edge', corpus.EDGE_QUERIES[::2]),
            ('skinner', corpus.SKINNER_QUERIES),
            ('scalars', corpus.EDGE_QUERIES[:6]),
            ('fuzz', corpus.EDGE_QUERIES[:30:4]),
            ('random', corpus.EDGE_QUERIES[:30:6])]
    # (+ the F path: fast.cuh's matcher, stages and piece-wise key functions)
    env = dict(os.environ, DNG_HOSTCHECK_TMPL='1', DNG_HOSTCHECK_FAST='1',
               DNG_HOSTCHECK_F='1', ASAN_OPTIONS='detect_leaks=0')
    n = 0
    for name, queries in jobs:
        path = tmp_path / (name + '.log')
        path.write_bytes(b'\n'.join(files[name]) + b'\n')
        for argv, ds in queries:
            pf = tmp_path / 'plan.json'
            pf.write_text(json.dumps(corpus.make_plan(argv, ds)))
            r = subprocess.run([asan_exe, str(pf), str(path)],
                               capture_output=True, env=env)
            err = r.stderr.decode('utf-8', 'replace')
            assert r.returncode == 0 and 'runtime error' not in err and \
                'AddressSanitizer' not in err, (name, argv, err[:2000])
            n += 1
    assert n > 40
