"""Pins the Python oracle (oracle/dn_oracle.py) to every `dn scan` golden the
reference's tests hold for the raw-scan path."""

import os
import sys

import pytest

sys.path.insert(0, os.path.join(os.path.dirname(__file__), '..', 'oracle'))
import dn_oracle  # noqa: E402

from golden_harness import check_section  # noqa: E402


def py_engine(plan, files):
    def chunks():
        for p in files:
            with open(p, 'rb') as f:
                while True:
                    b = f.read(16834)   # lib/datasource-file.js:264
                    if not b:
                        break
                    yield b
    return dn_oracle.scan(plan, chunks())


@pytest.mark.parametrize('suite', ['scan_file', 'scan_fileset', 'empty',
                                   'scan_manta'])
def test_python_oracle_matches_reference_goldens(suite, goldens, datadir):
    n = 0
    for i, sec in enumerate(goldens['suites'][suite]):
        if sec['cmd'] != 'scan':
            continue
        if suite == 'scan_manta' and ('--counters' in sec['argv'] or
                                      '--dry-run' in sec['argv'] or
                                      '-n' in sec['argv']):
            continue
        check_section(py_engine, suite, i, sec, datadir)
        n += 1
    assert n >= 8
