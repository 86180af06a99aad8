"""The GPU datasource: what `datasourceForConfig` returns for backend "gpu"
(and, as a drop-in, for "file").

Host-side mirror of the reference's Datasource interface for the scan path:
  lib/dragnet.js:288-304        datasourceForConfig (backend dispatch)
  lib/datasource-file.js:31-56  createDatasource / DatasourceFile fields
  lib/datasource-file.js:72-108 scan({query, dryRun})
  lib/datasource-file.js:124-192 scanInit (timeField check, dry run)
The reference returns an object-mode stream of skinner points; here scan()
returns a ScanResult holding the points, the per-stage counters and the list of
files.  build/query/index* are outside the GPU hot path
and raise NotImplementedError.
"""

from . import native
from . import query as mod_query


class DsError(Exception):
    pass


class ScanResult(object):
    def __init__(self, points, counters, files, flat=None, stats=None):
        self.points = points        # [([(name, bytes|float)...], value)]
        self.counters = counters    # {stage: {counter: n}} vstream style
        self.files = files
        self.flat_counters = flat   # dng_counters as a dict
        self.stats = stats


def stage_counters(plan, c, npoints):
    """dng_counters -> vstream-style per-stage counters (bin/dn:911-916)."""
    out = {}

    def put(stage, name, n):
        if n:
            out.setdefault(stage, {})[name] = n

    n = c['lines']
    put('json parser', 'ninputs', n)
    put('json parser', 'invalid json', c['invalid_json'])
    n -= c['invalid_json']
    put('json parser', 'noutputs', n)
    if plan.get('format', 'json') == 'json':
        put('SkinnerAdapterStream', 'ninputs', n)
        put('SkinnerAdapterStream', 'noutputs', n)
    else:
        put('json parser', 'invalid point', c.get('invalid_point', 0))
        n -= c.get('invalid_point', 0)

    def filt(stage, present, kf, ke):
        nonlocal n
        if not present:
            return
        put(stage, 'ninputs', n)
        put(stage, 'nfilteredout', c[kf])
        put(stage, 'nfailedeval', c[ke])
        n -= c[kf] + c[ke]
        put(stage, 'noutputs', n)

    filt('Datasource filter', plan.get('ds_filter'), 'ds_filtered',
         'ds_failedeval')
    filt('User filter', plan.get('filter'), 'user_filtered',
         'user_failedeval')
    if plan.get('synthetic'):
        put('Datetime parser', 'ninputs', n)
        put('Datetime parser', 'undef', c['synth_undef'])
        put('Datetime parser', 'baddate', c['synth_baddate'])
        n -= c['synth_undef'] + c['synth_baddate']
        put('Datetime parser', 'noutputs', n)
    filt('Time filter', plan.get('time_bounds'), 'time_filtered',
         'time_failedeval')
    # (what is left must be what the aggregator counted)
    naggr = c.get('aggr_ninputs', c.get('aggr'))
    if naggr is not None and n != naggr:
        raise DsError('stage counters do not add up: %r' % (c,))
    put('Aggregator', 'ninputs', n)
    put('Aggregator', 'noutputs', npoints)
    return out


def run_plan(plan, files=(), chunks=None, device=0, device_buffers=None,
             templates=None):
    """Drive one scan through the C ABI.  Input is any of: files (read by the
    library), host byte chunks, or (ptr, len) device buffers."""
    import json
    p = native.Plan(json.dumps(plan, separators=(',', ':')))
    s = native.Scan(p, device)
    try:
        if templates is not None:
            s.set_templates(templates)
        for f in files:
            s.feed_file(f)
        for c in (chunks or ()):
            s.feed(c)
        for ptr, n in (device_buffers or ()):
            s.feed_device(ptr, n)
        res = s.finish()
        flat = s.counters()
        stats = s.kernel_stats()
        stats.update(s.template_stats())
        pts = []
        if 'metrics' in plan:
            # fan-out: points are tagged like the reference tags them
            # (fields.__dn_metric = qi, lib/datasource-file.js:412-419)
            mflat = [s.counters(m) for m in range(len(plan['metrics']))]
            for m, cols, value in res.points(with_metric=True):
                names = [b['name'] for b in plan['metrics'][m]['breakdowns']]
                pts.append((list(zip(names, cols)) + [('__dn_metric', m)],
                            value))
            res.close()
            counters = []
            for m, mp in enumerate(plan['metrics']):
                one = dict(mp, format=plan.get('format', 'json'),
                           ds_filter=plan.get('ds_filter'))
                npts = sum(1 for f, _ in pts if f[-1][1] == m)
                counters.append(stage_counters(one, mflat[m], npts))
            return ScanResult(pts, counters, list(files), mflat, stats)
        names = [b['name'] for b in plan['breakdowns']]
        for cols, value in res.points():
            pts.append((list(zip(names, cols)), value))
        res.close()
        return ScanResult(pts, stage_counters(plan, flat, len(pts)),
                          list(files), flat, stats)
    finally:
        s.close()
        p.close()


class DatasourceGpu(object):
    def __init__(self, args):
        dsconfig = args['dsconfig']
        bc = dsconfig.get('backend_config') or dsconfig
        self.ds_format = dsconfig.get('dataFormat') or \
            dsconfig.get('ds_format') or 'json'
        self.ds_timeformat = bc.get('timeFormat') or None
        self.ds_timefield = bc.get('timeField') or None
        self.ds_datapath = bc.get('path')
        self.ds_filter = dsconfig.get('filter') or None
        self.ds_device = dsconfig.get('device', 0)
        # (root, timeFormat, after_ms, before_ms) -> [paths]: the reference's
        # own enumeration (lib/datasource-file.js:218-246 findStream), which
        # the real integration keeps; the tests pass their mirror of it
        self.ds_find = args.get('findFiles')

    def close(self):
        pass

    def scan(self, args):
        """lib/datasource-file.js:72-108.  Returns ScanResult or an Error."""
        query = args['query']
        dry = args['dryRun']
        assert isinstance(dry, bool)
        if self.ds_timefield is None and (query.qc_before is not None or
                                          query.qc_after is not None):
            return DsError('datasource is missing "timefield" for "before" '
                           'and "after" constraints')
        if self.ds_format not in ('json', 'json-skinner'):
            return DsError('unsupported format: "%s"' % self.ds_format)
        if self.ds_find is None:
            return DsError('no file enumerator ("findFiles")')
        files = self.ds_find(self.ds_datapath, self.ds_timeformat,
                             query.qc_after, query.qc_before)
        if dry:
            return ScanResult([], {}, files)
        plan = mod_query.scan_plan(query, ds_filter=self.ds_filter,
                                   time_field=self.ds_timefield,
                                   data_format=self.ds_format)
        return run_plan(plan, files=files, device=self.ds_device)

    def indexScan(self, args):
        """The scan half of build/index-scan (lib/datasource-file.js:321-433):
        one pass over the raw data computing every metric; returns the points
        an IndexSink would store, each tagged ('__dn_metric', i).  args:
        metrics [{'filter', 'breakdowns'}], interval, dryRun, timeAfter,
        timeBefore (ms)."""
        after, before = args.get('timeAfter'), args.get('timeBefore')
        if self.ds_timefield is None and (args['interval'] != 'all' or
                                          after or before):
            return DsError('datasource is missing "timefield"')
        if self.ds_find is None:
            return DsError('no file enumerator ("findFiles")')
        files = self.ds_find(self.ds_datapath, self.ds_timeformat,
                             after, before)
        if args.get('dryRun'):
            return ScanResult([], {}, files)
        queries = [mod_query.metricQuery(m, after, before, args['interval'],
                                         self.ds_timefield)
                   for m in args['metrics']]
        plan = mod_query.scan_plan_multi(queries, ds_filter=self.ds_filter,
                                         time_field=self.ds_timefield,
                                         data_format=self.ds_format)
        return run_plan(plan, files=files, device=self.ds_device)

    def build(self, *a, **k):
        raise NotImplementedError('writing sqlite indexes stays on the '
                                  'reference path (IndexSink); use indexScan '
                                  'for the points')

    query = indexRead = build


def createDatasource(args):
    dsconfig = args['dsconfig']
    bc = dsconfig.get('backend_config') or dsconfig
    if not isinstance(bc.get('path'), str):
        return DsError('expected datasource "path" to be a string')
    return DatasourceGpu(args)


def datasourceForConfig(args):
    """lib/dragnet.js:288-304 with the one added branch INTEGRATION.md shows."""
    bename = args['dsconfig'].get('backend', 'gpu')
    if bename in ('gpu', 'file'):
        return createDatasource(args)
    return DsError('unknown datasource backend: "%s"' % bename)
