"""Device per-record logic (host-compiled, TEST ONLY) vs the Python oracle on
edge-case and randomized inputs: JSON grammar corners, escapes, duplicate
keys, pluck precedence, JS number/string coercions, dates, json-skinner."""

import os
import sys

import pytest


@pytest.fixture(params=['general', 'fast'], autouse=True)
def parser_mode(request, monkeypatch):
    """Run every case with the general parser only, and with the lock-step
    fast automaton (+ fallback) the kernel uses."""
    if request.param == 'fast':
        monkeypatch.setenv('DNG_HOSTCHECK_FAST', '1')
    else:
        monkeypatch.delenv('DNG_HOSTCHECK_FAST', raising=False)

sys.path.insert(0, os.path.dirname(__file__))
import corpus  # noqa: E402
from engines import canon_points, hostcheck_engine, py_engine  # noqa: E402


def _write(tmp_path, name, lines):
    p = tmp_path / name
    p.write_bytes(b'\n'.join(lines) + b'\n')
    return str(p)


def _compare(plan, path):
    exp_p, exp_c = py_engine(plan, [path])
    act_p, act_c = hostcheck_engine(plan, [path])
    assert canon_points(act_p) == canon_points(exp_p), plan
    assert act_c == exp_c, plan


@pytest.mark.parametrize('qi', range(len(corpus.EDGE_QUERIES)))
def test_edge_lines(qi, tmp_path):
    argv, ds = corpus.EDGE_QUERIES[qi]
    path = _write(tmp_path, 'edge.log', corpus.EDGE_LINES)
    _compare(corpus.make_plan(argv, ds), path)


@pytest.mark.parametrize('qi', range(len(corpus.SKINNER_QUERIES)))
def test_skinner_lines(qi, tmp_path):
    argv, ds = corpus.SKINNER_QUERIES[qi]
    path = _write(tmp_path, 'sk.log', corpus.SKINNER_LINES)
    _compare(corpus.make_plan(argv, ds), path)


@pytest.mark.parametrize('seed', range(6))
def test_random_lines(seed, tmp_path):
    path = _write(tmp_path, 'rand.log', corpus.random_lines(seed, 400))
    for argv, ds in corpus.EDGE_QUERIES[:24:3]:
        _compare(corpus.make_plan(argv, ds), path)
