"""Multi-metric fan-out (the scan half of `dn build` / index-scan,
lib/datasource-file.js:321-433): one pass computing several metrics must give,
for every metric, exactly what a separate `dn scan` of that metric gives --
checked against the oracle per metric (CPU, host build of the device logic)
and through the C ABI on the GPU."""

import os
import sys

import pytest

sys.path.insert(0, os.path.dirname(__file__))
import corpus  # noqa: E402
from engines import canon_points, hostcheck_multi, py_engine  # noqa: E402
from dragnet_b200 import query as mod_query  # noqa: E402

METRICS = [
    {'filter': None, 'breakdowns': []},
    {'filter': {'eq': ['req.method', 'GET']},
     'breakdowns': [{'name': 'operation', 'field': 'operation'},
                    {'name': 'res.statusCode', 'field': 'res.statusCode'}]},
    {'filter': None,
     'breakdowns': [{'name': 'latency', 'field': 'latency',
                     'aggr': 'quantize'},
                    {'name': 'host', 'field': 'host'}]},
    {'filter': {'ne': ['req.caller', 'admin']},
     'breakdowns': [{'name': 'req.caller', 'field': 'req.caller'}]},
]

DROPS = ['user_filtered', 'user_failedeval', 'synth_undef', 'synth_baddate',
         'time_filtered', 'time_failedeval']


def _plans(interval, after=None, before=None):
    queries = [mod_query.metricQuery(m, after, before, interval, 'time')
               for m in METRICS]
    multi = mod_query.scan_plan_multi(queries, ds_filter={'ne': ['host',
                                                                 'ralph']},
                                      time_field='time')
    singles = [mod_query.scan_plan(q, ds_filter={'ne': ['host', 'ralph']},
                                   time_field='time') for q in queries]
    return multi, singles


def _files(datadir):
    return [os.path.join(datadir, '2014/05-02/one.log'),
            os.path.join(datadir, '2014/05-05/more.log')]


@pytest.mark.parametrize('fast', [False, True])
@pytest.mark.parametrize('interval,after,before', [
    ('all', None, None), ('hour', None, None),
    ('day', 1398988800000, 1399334400000)])
def test_fanout_equals_separate_scans_host(datadir, interval, after, before,
                                           fast):
    multi, singles = _plans(interval, after, before)
    per, flats = hostcheck_multi(multi, _files(datadir), fast)
    for m, single in enumerate(singles):
        exp_p, exp_c = py_engine(single, _files(datadir))
        assert canon_points(per.get(m, [])) == canon_points(exp_p), m
        if m:                     # metric 0's counters are the main ones
            f = flats[m]
            exp_user = exp_c.get('User filter', {})
            assert f['user_filtered'] == exp_user.get('nfilteredout', 0)
            assert f['user_failedeval'] == exp_user.get('nfailedeval', 0)
            assert f['aggr'] == exp_c.get('Aggregator', {}).get('ninputs', 0)


@pytest.mark.gpu
@pytest.mark.parametrize('interval', ['all', 'hour'])
def test_fanout_through_c_abi(datadir, interval, tmp_path):
    from dragnet_b200 import datasource_gpu, native
    n = 30000
    data = native.gen_host(native.gen_params(total_records=n), 0, n)
    p = tmp_path / 'syn.log'
    p.write_bytes(data)
    files = _files(datadir) + [str(p)]
    multi, singles = _plans(interval)
    res = datasource_gpu.run_plan(multi, files=files)
    for m, single in enumerate(singles):
        exp_p, exp_c = py_engine(single, files)
        got = [(f[:-1], v) for f, v in res.points if f[-1] == ('__dn_metric',
                                                               m)]
        assert canon_points(got) == canon_points(exp_p), m
        assert res.counters[m] == exp_c, m


@pytest.mark.gpu
def test_index_scan_datasource(datadir):
    """DatasourceGpu.indexScan mirrors the reference's build scan: hourly
    __dn_ts buckets are prepended unless interval is 'all'."""
    from dragnet_b200 import datasource_gpu
    from hostmirror import find as mod_find
    ds = datasource_gpu.datasourceForConfig({'dsconfig': {
        'backend': 'gpu', 'backend_config': {
            'path': datadir, 'timeFormat': '%Y/%m-%d', 'timeField': 'time'}},
        'findFiles': mod_find.find_files})
    res = ds.indexScan({'metrics': METRICS[:2], 'interval': 'day',
                        'dryRun': False})
    total = [(f, v) for f, v in res.points if f[-1][1] == 0]
    assert sorted(v for _, v in total) == [250, 500, 500, 500, 500]
    assert [f[0][0] for f, _ in total] == ['__dn_ts'] * 5
