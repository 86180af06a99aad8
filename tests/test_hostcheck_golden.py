"""The device per-record code (record.cuh, plan.cpp) compiled for the host by
tests/hostcheck -- a TEST-ONLY harness, not a product path -- must reproduce
every reference golden too.  This is what lets kernel logic be iterated here
(no GPU) before the `-m gpu` parity tests run through libdragnet_gpu.so."""

import os
import sys

import pytest


@pytest.fixture(params=['general', 'fast', 'tmpl', 'fpath', 'jit'], autouse=True)
def parser_mode(request, monkeypatch):
    """Run every case with the general parser only, with the lock-step fast
    automaton (+ fallback), and with record templates learned from the input
    in front of both -- the three tiers the general kernels use -- and with the
    F path (fast.cuh: path-indexed templates, lean stages, piece-wise keys) in
    front of all of them, as scan_kernel_f does."""
    monkeypatch.delenv('DNG_HOSTCHECK_FAST', raising=False)
    monkeypatch.delenv('DNG_HOSTCHECK_TMPL', raising=False)
    monkeypatch.delenv('DNG_HOSTCHECK_F', raising=False)
    monkeypatch.delenv('DNG_HOSTCHECK_JIT', raising=False)
    if request.param in ('fpath', 'jit'):
        monkeypatch.setenv('DNG_HOSTCHECK_F', '1')
    if request.param == 'jit':
        # + the matcher jit.cpp generates for the templates: built for the
        # device (NVRTC + nvJitLink, no GPU needed) and run on the host
        monkeypatch.setenv('DNG_HOSTCHECK_JIT', '1')
    if request.param in ('fast', 'tmpl'):
        monkeypatch.setenv('DNG_HOSTCHECK_FAST', '1')
    if request.param == 'tmpl':
        monkeypatch.setenv('DNG_HOSTCHECK_TMPL', '1')

sys.path.insert(0, os.path.dirname(__file__))
from engines import hostcheck_engine  # noqa: E402
from golden_harness import check_section  # noqa: E402


@pytest.mark.parametrize('suite', ['scan_file', 'scan_fileset', 'empty',
                                   'scan_manta'])
def test_device_record_logic_matches_reference_goldens(suite, goldens,
                                                       datadir):
    n = 0
    for i, sec in enumerate(goldens['suites'][suite]):
        if sec['cmd'] != 'scan':
            continue
        if suite == 'scan_manta' and ('--counters' in sec['argv'] or
                                      '--dry-run' in sec['argv'] or
                                      '-n' in sec['argv']):
            continue
        check_section(hostcheck_engine, suite, i, sec, datadir)
        n += 1
    assert n >= 8
