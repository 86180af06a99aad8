"""Query plan: host-side mirror of the reference's QueryConfig / queryLoad.

Reference: lib/dragnet.js:28-77 (QueryConfig), :103-144 (queryLoad),
:151-186 (parseTimeBounds), :191-244 (parseFields/parseField);
lib/dragnet-impl.js:66-125 (queryAggrStreamConfig, queryTimeBoundsFilter);
lib/stream-scan.js:56-86 (stage order, ``dn_ts`` synthetic).

Errors are returned (``isinstance(x, Exception)``) following the reference's
convention; programmer errors assert.

``scan_plan()`` serialises everything the GPU scan needs into the plan JSON
accepted by ``dng_plan_create`` (include/dragnet_gpu.h) and by oracle/.
"""

import copy
import json
import math

from . import krill
from . import jsdate


class DnError(Exception):
    """VError-like: message chaining with ': '."""

    def __init__(self, msg, cause=None):
        if cause is not None:
            msg = '%s: %s' % (msg, cause)
        Exception.__init__(self, msg)
        self.message = msg


def _js_parse_int(v):
    """parseInt(v, 10): returns int or None for NaN."""
    if isinstance(v, bool):
        return None
    if isinstance(v, (int, float)):
        if isinstance(v, float) and (math.isnan(v) or math.isinf(v)):
            return None
        s = _js_num_str(v)
    else:
        s = str(v)
    s = s.lstrip(' \t\n\r\v\f\u00a0\u1680\u2000\u2001\u2002\u2003\u2004\u2005'
                 '\u2006\u2007\u2008\u2009\u200a\u2028\u2029\u202f\u205f\u3000\ufeff')
    i = 0
    sign = 1
    if i < len(s) and s[i] in '+-':
        sign = -1 if s[i] == '-' else 1
        i += 1
    j = i
    while j < len(s) and s[j] in '0123456789':
        j += 1
    if j == i:
        return None
    return sign * int(s[i:j])


def _js_num_str(v):
    if isinstance(v, int):
        return str(v)
    if v == int(v) and abs(v) < 1e21:
        return str(int(v))
    return repr(v)


class QueryConfig(object):
    """Immutable parameters of one query (lib/dragnet.js:28-77)."""

    def __init__(self, filter=None, breakdowns=None, timeBefore=None,
                 timeAfter=None, timeField=None):
        assert filter is None or isinstance(filter, dict)
        assert isinstance(breakdowns, list)
        self.qc_filter = filter or None
        self.qc_breakdowns = copy.deepcopy(breakdowns)
        self.qc_before = timeBefore     # integer ms since epoch, or None
        self.qc_after = timeAfter
        self.qc_fieldsbyname = {}
        self.qc_bucketizers = {}
        self.qc_synthetic = []

        if timeField:
            self.qc_synthetic.append(
                {'name': timeField, 'field': timeField, 'date': ''})

        for fieldconf in self.qc_breakdowns:
            self.qc_fieldsbyname[fieldconf['name']] = fieldconf
            if 'date' in fieldconf:
                self.qc_synthetic.append(fieldconf)
            if 'aggr' not in fieldconf:
                continue
            if fieldconf['aggr'] == 'quantize':
                self.qc_bucketizers[fieldconf['name']] = P2Bucketizer()
                continue
            assert fieldconf['aggr'] == 'lquantize'
            assert isinstance(fieldconf['step'], (int, float))
            self.qc_bucketizers[fieldconf['name']] = \
                LinearBucketizer(fieldconf['step'])

        if self.qc_before is not None:
            assert self.qc_after is not None
        else:
            assert self.qc_after is None


class P2Bucketizer(object):
    """skinner.makeP2Bucketizer() (call site lib/dragnet.js:63): ordinal 0
    holds v < 1, ordinal i >= 1 holds [2^(i-1), 2^i)."""
    kind = 'p2'

    def bucketMin(self, i):
        if isinstance(i, float) and math.isnan(i):
            return float('nan')
        return 0 if i == 0 else 2 ** (int(i) - 1)


class LinearBucketizer(object):
    """skinner.makeLinearBucketizer(step) (call site lib/dragnet.js:70)."""
    kind = 'linear'

    def __init__(self, step):
        self.step = step

    def bucketMin(self, i):
        return i * self.step


def queryLoad(args):
    """lib/dragnet.js:103-144.  ``args`` = {'query': {...}, 'allowReserved'}."""
    assert isinstance(args, dict)
    q = args['query']
    assert isinstance(q, dict)
    assert isinstance(q.get('breakdowns'), list)

    if q.get('filter'):
        flt = q['filter']
        try:
            krill.createPredicate(flt)
        except krill.KrillError as ex:
            return DnError('invalid query: invalid filter', ex)
    else:
        flt = None

    breakdowns = parseFields(q['breakdowns'],
                             {'allowReserved': args.get('allowReserved')})
    if isinstance(breakdowns, Exception):
        return DnError('invalid query', breakdowns)

    tb = parseTimeBounds({'timeAfter': q.get('timeAfter'),
                          'timeBefore': q.get('timeBefore')})
    if isinstance(tb, Exception):
        return tb

    return QueryConfig(filter=flt, breakdowns=breakdowns,
                       timeAfter=tb['timeAfter'], timeBefore=tb['timeBefore'],
                       timeField=q.get('timeField'))


def _to_ms(v):
    """new Date(v).getTime(): number -> itself, string -> Date.parse."""
    if isinstance(v, bool):
        return None
    if isinstance(v, (int, float)):
        if isinstance(v, float) and (math.isnan(v) or math.isinf(v)):
            return None
        return int(v)
    return jsdate.date_parse_ms(str(v))


def parseTimeBounds(args):
    """lib/dragnet.js:151-186: both or neither; after <= before."""
    after = before = None
    if args.get('timeAfter'):
        if not args.get('timeBefore'):
            return DnError('"after" requires specifying "before" too')
        after = _to_ms(args['timeAfter'])
        if after is None:
            return DnError('"after": not a valid date: "%s"' %
                           args['timeAfter'])
        before = _to_ms(args['timeBefore'])
        if before is None:
            return DnError('"before": not a valid date: "%s"' %
                           args['timeBefore'])
        if after > before:
            return DnError('"after" timestamp may not come after "before"')
    elif args.get('timeBefore'):
        return DnError('"before" requires specifying "after" too')
    return {'timeAfter': after, 'timeBefore': before}


def parseFields(inputs, options=None):
    fields = []
    for i, b in enumerate(inputs):
        ret = parseField(b, options)
        if isinstance(ret, Exception):
            return DnError('field %d ("%s") is invalid' % (i, b), ret)
        fields.append(ret)
    return fields


def parseField(b, options=None):
    """lib/dragnet.js:210-244 (mutates and returns ``b`` like the reference)."""
    assert not isinstance(b, str)
    assert isinstance(b['name'], str)
    if 'aggr' in b:
        if b['aggr'] != 'quantize' and b['aggr'] != 'lquantize':
            return DnError('unsupported aggr: "%s"' % b['aggr'])
        if b['aggr'] == 'lquantize':
            if 'step' not in b:
                return DnError('aggr "lquantize" requires "step"')
            step = _js_parse_int(b['step'])
            if step is None:
                return DnError('aggr "lquzntize": invalid value for "step": '
                               '"%s"' % b['step'])
            b['step'] = step
    if not (options and options.get('allowReserved')) and \
            b['name'].startswith('__dn'):
        return DnError('field names starting with "__dn" are reserved')
    if 'field' not in b:
        b['field'] = b['name']
    return b


def queryTimeBoundsFilter(query, timefield):
    """lib/dragnet-impl.js:94-125: [ceil(after/1000), ceil(before/1000))."""
    if query.qc_before is not None:
        assert query.qc_after is not None
        return {'and': [
            {'ge': [timefield, -((-query.qc_after) // 1000)]},
            {'lt': [timefield, -((-query.qc_before) // 1000)]}]}
    assert query.qc_after is None
    return None


def queryAggrStreamConfig(query, options=None):
    """lib/dragnet-impl.js:66-89."""
    rv = dict(options or {})
    rv['bucketizers'] = query.qc_bucketizers
    rv['decomps'] = [b['name'] for b in query.qc_breakdowns]
    rv['ordinalBuckets'] = True
    return rv


def scan_plan(query, ds_filter=None, time_field=None, data_format='json'):
    """Serialise a scan (QueryConfig + datasource properties) as the plan JSON
    that crosses the C ABI.  This is the moral equivalent of the arguments
    ``DatasourceFile.scan`` hands to ``StreamScan``
    (lib/datasource-file.js:72-108, lib/stream-scan.js:40-94).
    """
    # StreamScan appends the dn_ts synthetic when time bounds are present
    # (lib/stream-scan.js:62-69); it needs the datasource's timeField.
    synthetic = [{'name': s['name'], 'field': s['field']}
                 for s in query.qc_synthetic]
    bounds = None
    if query.qc_before is not None or query.qc_after is not None:
        assert isinstance(time_field, str)
        synthetic.append({'name': 'dn_ts', 'field': time_field})
        f = queryTimeBoundsFilter(query, 'dn_ts')
        bounds = {'field': 'dn_ts', 'ge': f['and'][0]['ge'][1],
                  'lt': f['and'][1]['lt'][1]}
    bds = []
    for b in query.qc_breakdowns:
        e = {'name': b['name'], 'field': b['field']}
        if 'date' in b:
            e['date'] = True
        if 'aggr' in b:
            e['aggr'] = b['aggr']
            if b['aggr'] == 'lquantize':
                e['step'] = b['step']
        bds.append(e)
    return {
        'format': data_format,
        'ds_filter': ds_filter or None,
        'filter': query.qc_filter,
        'synthetic': synthetic,
        'time_bounds': bounds,
        'breakdowns': bds,
    }


def metricQuery(metric, after, before, interval, timefield):
    """lib/dragnet-impl.js:290-323: the query that computes one metric of an
    index; unless interval is 'all' a `__dn_ts[lquantize, step, date]`
    breakdown is prepended.  ``metric`` = {'filter', 'breakdowns': [...]}
    (config form: name/field[/date/aggr/step])."""
    qconf = {'filter': metric.get('filter'),
             'breakdowns': copy.deepcopy(metric.get('breakdowns') or [])}
    if interval != 'all':
        step = {'hour': 3600, 'day': 86400}[interval]
        qconf['breakdowns'].insert(0, {
            'name': '__dn_ts', 'aggr': 'lquantize', 'step': step,
            'field': timefield, 'date': ''})
    if after:
        qconf['timeAfter'] = after
    if before:
        qconf['timeBefore'] = before
    q = queryLoad({'allowReserved': True, 'query': qconf})
    assert not isinstance(q, Exception), q
    return q


def scan_plan_multi(queries, ds_filter=None, time_field=None,
                    data_format='json'):
    """One pass, several metrics: what DatasourceFile.indexScanImpl wires up
    (lib/datasource-file.js:386-432): the datasource filter sits on the
    parser, every metric gets its own StreamScan."""
    metrics = []
    for q in queries:
        p = scan_plan(q, ds_filter=None, time_field=time_field,
                      data_format=data_format)
        metrics.append({'filter': p['filter'], 'synthetic': p['synthetic'],
                        'time_bounds': p['time_bounds'],
                        'breakdowns': p['breakdowns']})
    return {'format': data_format, 'ds_filter': ds_filter or None,
            'metrics': metrics}


def scan_plan_json(*args, **kwargs):
    return json.dumps(scan_plan(*args, **kwargs), separators=(',', ':'))
