"""The reference goldens that tests/golden_harness.py does not cover, replayed
through every engine (both oracles, the host build of the device code, and --
under `-m gpu` -- the CUDA path through the C ABI):

  tst.format_skinner.sh.out  `dn scan --points` output fed back as json-skinner
                             input once, twice, three times: weights add up
                             (250 / 500 / 750) and the flat "req.method" key of
                             a point is plucked whole-key-first (222/162/183/183)
  tst.scan_250k.sh.out       250 000 generated records in small chunks: 250000
  tst.badargs.sh.out         the error messages of bad breakdowns / filters /
                             formats (as far as the scan path produces them)

Fixtures: tests/golden/scan_goldens.json, generated from /root/reference by
tests/golden/make_golden.py."""

import io
import os
import sys

import pytest

sys.path.insert(0, os.path.dirname(__file__))
from engines import cpp_engine, hostcheck_engine, py_engine  # noqa: E402

from hostmirror import dn as mod_dn  # noqa: E402
from dragnet_b200 import query as mod_query  # noqa: E402


def gpu_engine(plan, files):
    from dragnet_b200 import datasource_gpu
    r = datasource_gpu.run_plan(plan, files=files)
    return r.points, r.counters


ENGINES = [pytest.param(py_engine, id='oracle-py'),
           pytest.param(cpp_engine, id='oracle-cpp'),
           pytest.param(hostcheck_engine, id='device-code-on-host'),
           pytest.param(gpu_engine, id='gpu', marks=pytest.mark.gpu)]


def scan(engine, argv, files, data_format='json'):
    options = mod_dn.dnParseArgs(list(argv))
    q = mod_dn.dnQueryConfig(options)['query']
    plan = mod_query.scan_plan(q, data_format=data_format)
    points, _ = engine(plan, files)
    return mod_dn.render_scan(q, options, points, 'stdin')


@pytest.mark.parametrize('engine', ENGINES)
def test_format_skinner_golden(engine, goldens, datadir, tmp_path):
    """/root/reference/tests/dn/local/tst.format_skinner.sh:27-40 (the scan
    part; the index part is out of scope) against its .out."""
    one = os.path.join(datadir, '2014/05-01/one.log')
    exp = goldens['format_skinner']
    exp = exp[:exp.index('building index')]
    out = []

    def cat(text, n, name):
        p = tmp_path / name
        p.write_text(text * n)
        return [str(p)]

    # points with no fields
    pts = scan(engine, ['--points'], [one])
    for n in (1, 2, 3):
        out.append('# dn scan stdin-skinner\n' +
                   scan(engine, [], cat(pts, n, 'p0_%d' % n), 'json-skinner'))
    # points with a couple of fields
    pts = scan(engine, ['--points', '-b', 'req.method,res.statusCode'], [one])
    out.append(scan(engine, ['-b', 'req.method'], [one]))
    out.append('# dn scan stdin-skinner\n' +
               scan(engine, [], cat(pts, 3, 'p1'), 'json-skinner'))
    out.append('# dn scan stdin-skinner -b req.method\n' +
               scan(engine, ['-b', 'req.method'], cat(pts, 3, 'p2'),
                    'json-skinner'))
    assert ''.join(out) == exp


@pytest.mark.parametrize('engine', [ENGINES[1], ENGINES[3]])
def test_scan_250k_golden(engine, goldens, tmp_path):
    """tst.scan_250k.sh:30-43: 250 000 generated records, counted.  (The
    reference asserts a memory ceiling as well: the scan's memory does not
    depend on the number of records here either -- fixed device buffers.)"""
    from dragnet_b200 import native
    n = 250000
    data = native.gen_host(native.gen_params(total_records=n), 0, n)
    p = tmp_path / 'in.log'
    p.write_bytes(data)
    assert scan(engine, [], [str(p)]).split() == goldens['scan_250k'].split()


@pytest.mark.gpu
def test_scan_250k_in_small_chunks_through_the_c_abi(goldens):
    """The same count with the input fed in 16834-byte pieces (the reference's
    read size, lib/datasource-file.js:264): lstream's carry across chunks."""
    from dragnet_b200 import datasource_gpu, native
    n = 250000
    data = native.gen_host(native.gen_params(total_records=n), 0, n)
    plan = mod_query.scan_plan(mod_dn.dnQueryConfig(
        mod_dn.dnParseArgs([]))['query'])
    chunks = [data[i:i + 16834] for i in range(0, len(data), 16834)]
    r = datasource_gpu.run_plan(plan, chunks=chunks)
    assert [v for _, v in r.points] == [n]
    assert str(n) in goldens['scan_250k']


def test_badargs_golden(goldens, datadir):
    """tst.badargs.sh:27-38 against its .out, through dn.main (argument
    parsing, query validation, datasource format check: the messages a scan
    can produce; nothing here reaches the device)."""
    exp = goldens['badargs'].split('\n')
    one = os.path.join(datadir, '2014/05-01/one.log')
    ds = {'input': {'backend': 'file', 'path': one, 'dataFormat': 'json'}}

    def run(argv, datasources=ds):
        out, err = io.StringIO(), io.StringIO()
        mod_dn.main(['scan'] + argv + ['input'], datasources, out, err)
        return err.getvalue().split('\n')

    got = []
    got += run(['-b', 'host', '-b', 'req.method,x[=bar]'])[:2]
    got += run(['-b', 'host', '-b', 'req.method,[]'])[:2]
    got += run(['-b', 'host', '-b', 'req.method,foo['])[:2]
    assert got == exp[:6]
    # JSON.parse's message for '{' is V8's ("Unexpected end of input"): only
    # its prefix is ours to reproduce
    e = run(['-f', '{'])
    assert e[0].startswith('dn: invalid filter: ') and \
        e[1] == 'usage: dn SUBCOMMAND [OPTIONS] ARGS'
    e = run(['-f', '{ "junk": [ "foo", "bar" ] }'])
    assert e[0] == exp[8]
    junk = {'input': {'backend': 'file', 'path': one, 'dataFormat': 'junk'}}
    e = run([], junk)
    assert e[0] == exp[11]
