"""Record templates (csrc/tmpl.*), host side: the lexical skeletons, the trie
builder and the matcher compiled for the host by tests/hostcheck (TEST ONLY).
What they produce is compared with the oracle in test_hostcheck_*.py; here the
point is coverage: that the shapes one expects to be templated are."""

import json
import os
import subprocess
import sys

import pytest

sys.path.insert(0, os.path.dirname(__file__))
import corpus  # noqa: E402
from engines import build_hostcheck  # noqa: E402

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def run_tmpl(plan, path, tmp_path):
    exe = build_hostcheck()
    pf = tmp_path / 'plan.json'
    pf.write_text(json.dumps(plan))
    env = dict(os.environ, DNG_HOSTCHECK_TMPL='1', DNG_HOSTCHECK_FAST='1',
               DNG_HOSTCHECK_TMPL_DEBUG='1')
    r = subprocess.run([exe, str(pf), path], capture_output=True, env=env,
                       check=True)
    return json.loads(r.stdout), r.stderr.decode()


@pytest.mark.parametrize('name', sorted(corpus.BASELINE_QUERIES))
def test_generated_workload_is_fully_templated(name, tmp_path):
    """mktestdata-shaped records have three shapes (caller string / null /
    absent): every record must be taken by a template, whatever the query."""
    from dragnet_b200 import native
    n = 5000
    p = tmp_path / 'gen.log'
    p.write_bytes(native.gen_host(native.gen_params(total_records=n), 0, n))
    argv, ds = corpus.BASELINE_QUERIES[name]
    doc, dbg = run_tmpl(corpus.make_plan(argv, ds), str(p), tmp_path)
    assert doc['counters']['lines'] == n
    assert doc['ntmpl'] == n, dbg
    assert '3 leaves' in dbg, dbg


def test_shapes_that_cannot_be_templates_fall_through(tmp_path):
    """A plan slot that is an array length (an inline constant) or an invalid
    sample line yields no template; escapes and odd numbers in VALUES do not
    stop a record from matching one."""
    lines = [b'{"a":[1,2,3],"s":"x"}'] * 40 + \
            [b'{"a":"v\\u0041","n":1.5e3}', b'{"a":"w\\\\\\"","n":-0}'] * 20 + \
            [b'{"a":"v","n":}'] * 5
    p = tmp_path / 'mix.log'
    p.write_bytes(b'\n'.join(lines) + b'\n')
    doc, dbg = run_tmpl(corpus.make_plan(['-b', 'a.length']), str(p), tmp_path)
    # the array shape needs index semantics: never templated
    assert doc['ntmpl'] == 40, (doc['ntmpl'], dbg)
    assert doc['counters']['invalid_json'] == 5
    doc, dbg = run_tmpl(corpus.make_plan(['-b', 'a,n']), str(p), tmp_path)
    assert doc['ntmpl'] == 80, (doc['ntmpl'], dbg)
