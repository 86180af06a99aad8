"""Replays the reference's golden `dn scan` sections (tests/dn/scan_testcases.sh
as run by tst.scan_file.sh / tst.scan_fileset.sh / tst.empty.sh /
tst.scan_manta.sh) against any scan engine.

An engine is ``fn(plan_dict, [paths]) -> (points, counters)`` with the oracle's
point shape: ([(name, bytes|float), ...], value).
"""

import os

from hostmirror import dn as mod_dn
from hostmirror import find as mod_find
from dragnet_b200 import query as mod_query

OUR_STAGES = set(mod_dn.STAGE_ORDER) | {'Flattener'}


def suite_dsconfig(suite, index, datadir):
    if suite == 'scan_file':
        ds = {'path': os.path.join(datadir, '2014/05-01/one.log')}
        if index >= 26:
            ds['filter'] = {'eq': ['req.method', 'GET']}
        return ds
    if suite in ('scan_fileset', 'scan_manta'):
        ds = {'path': datadir, 'timeFormat': '%Y/%m-%d', 'timeField': 'time'}
        if suite == 'scan_manta' and index >= 38:
            ds['filter'] = {'eq': ['req.method', 'GET']}
        return ds
    if suite == 'empty':
        return {'path': '/dev/null'}
    raise KeyError(suite)


FOREIGN_STAGES = ('FindStart ', 'FindStatter ', 'FindTraverser ',
                  'FindFeedback ', 'PathEnumerator ')


def _strip_foreign_counters(text):
    """Drop counters of stages outside the scan hot path (file finder, path
    enumerator): SURVEY.md section 8 scopes them out."""
    return '\n'.join(l for l in text.split('\n')
                     if not l.startswith(FOREIGN_STAGES))


def expected_text(section):
    text = section['text']
    cut = text.find('#\n# This is a GNUplot')
    if cut != -1:
        text = text[:cut]
    return _strip_foreign_counters(text).rstrip('\n')


def run_section(engine, suite, index, section, datadir):
    """Returns (actual, expected, is_points)."""
    argv = list(section['argv'])
    options = mod_dn.dnParseArgs(argv)
    ds = suite_dsconfig(suite, index, datadir)
    scanargs = mod_dn.dnQueryConfig(options)
    q = scanargs['query']
    files = mod_find.find_files(ds['path'], ds.get('timeFormat'),
                                q.qc_after, q.qc_before)
    exp = expected_text(section)
    if scanargs['dryRun']:
        root = os.path.dirname(os.path.dirname(datadir))
        act = 'would scan files:\n' + ''.join(
            '    %s\n' % os.path.relpath(f, root) for f in files)
        return act.rstrip('\n'), exp, False
    plan = mod_query.scan_plan(q, ds_filter=ds.get('filter'),
                               time_field=ds.get('timeField'))
    points, counters = engine(plan, files)
    out = mod_dn.render_scan(q, options, points, 'test_input')
    err = ''
    if options.get('counters'):
        err = mod_dn.format_counters(counters, not options.get('points'))
    if options.get('points'):
        act = err + out        # stderr is not sorted by the test's `sort -d`
    else:
        act = out + err
    return act.rstrip('\n'), exp, bool(options.get('points'))


def check_section(engine, suite, index, section, datadir):
    act, exp, is_points = run_section(engine, suite, index, section, datadir)
    if is_points:
        assert sorted(act.split('\n')) == sorted(exp.split('\n')), \
            (suite, index, section['header'])
    else:
        assert act == exp, (suite, index, section['header'])
