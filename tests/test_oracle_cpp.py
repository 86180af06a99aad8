"""Pins the C++ oracle (oracle/dn_oracle.cpp, the timed CPU baseline) to the
reference's goldens and cross-checks it against the Python oracle on the
edge-case corpus, single- and multi-threaded."""

import os
import sys

import pytest

sys.path.insert(0, os.path.dirname(__file__))
import corpus  # noqa: E402
from engines import canon_points, cpp_engine, py_engine  # noqa: E402
from golden_harness import check_section  # noqa: E402


@pytest.mark.parametrize('suite', ['scan_file', 'scan_fileset', 'empty',
                                   'scan_manta'])
def test_cpp_oracle_matches_reference_goldens(suite, goldens, datadir):
    n = 0
    for i, sec in enumerate(goldens['suites'][suite]):
        if sec['cmd'] != 'scan':
            continue
        if suite == 'scan_manta' and ('--counters' in sec['argv'] or
                                      '--dry-run' in sec['argv'] or
                                      '-n' in sec['argv']):
            continue
        check_section(cpp_engine, suite, i, sec, datadir)
        n += 1
    assert n >= 8


def _write(tmp_path, name, lines):
    p = tmp_path / name
    p.write_bytes(b'\n'.join(lines) + b'\n')
    return str(p)


@pytest.mark.parametrize('qi', range(len(corpus.EDGE_QUERIES)))
def test_edge_lines(qi, tmp_path):
    argv, ds = corpus.EDGE_QUERIES[qi]
    plan = corpus.make_plan(argv, ds)
    path = _write(tmp_path, 'edge.log', corpus.EDGE_LINES)
    exp_p, exp_c = py_engine(plan, [path])
    for threads in (1, 3):
        act_p, act_c = cpp_engine(plan, [path], threads)
        assert canon_points(act_p) == canon_points(exp_p)
        assert act_c == exp_c


@pytest.mark.parametrize('qi', range(len(corpus.SKINNER_QUERIES)))
def test_skinner_lines(qi, tmp_path):
    argv, ds = corpus.SKINNER_QUERIES[qi]
    plan = corpus.make_plan(argv, ds)
    path = _write(tmp_path, 'sk.log', corpus.SKINNER_LINES)
    exp_p, exp_c = py_engine(plan, [path])
    act_p, act_c = cpp_engine(plan, [path])
    assert canon_points(act_p) == canon_points(exp_p)
    assert act_c == exp_c


@pytest.mark.parametrize('seed', range(3))
def test_random_lines(seed, tmp_path):
    path = _write(tmp_path, 'rand.log', corpus.random_lines(100 + seed, 500))
    for argv, ds in corpus.EDGE_QUERIES[1:24:3]:
        plan = corpus.make_plan(argv, ds)
        exp_p, exp_c = py_engine(plan, [path])
        act_p, act_c = cpp_engine(plan, [path], 4)
        assert canon_points(act_p) == canon_points(exp_p), argv
        assert act_c == exp_c, argv
