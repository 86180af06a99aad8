"""`dn scan`-shaped front end over a Datasource: argument parsing, query
construction and result rendering, so that parity tests read like the
reference's own golden-file tests (tests/dn/scan_testcases.sh).

Host-side mirror of the scan-related parts of the reference CLI:
  bin/dn:247-309  dnParseArgs / dnExpandArray
  bin/dn:696-724  dnQueryConfig
  bin/dn:924-967  dnOutput (points vs flattened)
  bin/dn:972-1199 raw / pretty / quantized printers
  bin/dn:1205-1274 gnuplot printer
  bin/dn:911-916  --counters dump (vstream vsDumpCounters format)
  lib/skinner-flattener.js (re-aggregation into flat rows)

The CLI itself is out of scope of the GPU hot path: this
module never touches record data, only the <= #tuples points a scan emits.  It
is TEST INFRASTRUCTURE (the harness that replays the reference's golden
files), not part of the product package.
"""

import json
import math
import sys

from dragnet_b200 import jsdate
from dragnet_b200 import query as mod_query
from .attr_parser import attrsParse
from . import find as mod_find


class UsageError(Exception):
    pass


class FatalError(Exception):
    pass


SCAN_BOOL_OPTS = {'--raw': 'raw', '--points': 'points',
                  '--counters': 'counters', '--warnings': 'warnings',
                  '--gnuplot': 'gnuplot', '--dry-run': 'dry_run',
                  '-n': 'dry_run'}
SCAN_VALUE_OPTS = {'--before': 'before', '-B': 'before', '--after': 'after',
                   '-A': 'after', '--filter': 'filter', '-f': 'filter',
                   '--breakdowns': 'breakdowns', '-b': 'breakdowns',
                   '--assetroot': 'assetroot'}


def _parse_date_opt(name, v):
    """dashdash 'date' option type: epoch seconds or an ISO-8601 stamp."""
    if v.isdigit():
        return int(v) * 1000
    ms = jsdate.date_parse_ms(v)
    if ms is None:
        raise UsageError('arg for "--%s" is not a valid date format: "%s"'
                         % (name, v))
    return ms


def dnParseArgs(argv):
    """bin/dn:247-271 for the `scan` option set.  Returns an options dict."""
    opts = {'breakdowns': [], '_args': [], 'dry_run': False}
    i = 0
    while i < len(argv):
        a = argv[i]
        if a in SCAN_BOOL_OPTS:
            opts[SCAN_BOOL_OPTS[a]] = True
        elif a in SCAN_VALUE_OPTS or (a.startswith('--') and '=' in a and
                                      a.split('=', 1)[0] in SCAN_VALUE_OPTS):
            if '=' in a and a.startswith('--'):
                a, v = a.split('=', 1)
            else:
                i += 1
                if i >= len(argv):
                    raise UsageError('do not have enough args for "%s" '
                                     'option' % a)
                v = argv[i]
            name = SCAN_VALUE_OPTS[a]
            if name == 'breakdowns':
                opts['breakdowns'].append(v)
            elif name in ('before', 'after'):
                opts[name] = _parse_date_opt(name, v)
            else:
                opts[name] = v
        elif a.startswith('-') and a != '-':
            raise UsageError('unknown option: "%s"' % a)
        else:
            opts['_args'].append(a)
        i += 1

    dnExpandArray(opts, 'breakdowns')
    if opts.get('filter'):
        try:
            opts['filter'] = json.loads(opts['filter'])
        except ValueError as ex:
            msg = str(ex)
            # V8's wording for the one case the reference's goldens pin
            # (tests/dn/local/tst.badargs.sh.out:7).
            if opts['filter'].strip() in ('{', '[', '') or \
                    'Expecting' in msg and msg.endswith(
                        '(char %d)' % len(opts['filter'])):
                msg = 'Unexpected end of input'
            raise UsageError('invalid filter: %s' % msg)
    return opts


def dnExpandArray(options, field):
    """bin/dn:283-309: split each -b on the attribute grammar."""
    tmp = options[field]
    options[field] = []
    for v in tmp:
        lst = attrsParse(v)
        if isinstance(lst, Exception):
            raise UsageError('bad value for "%s" ("%s"): %s' %
                             (field, v, lst))
        for s in lst:
            if not s.get('field'):
                s['field'] = s['name']
            if 'step' in s:
                step = mod_query._js_parse_int(s['step'])
                if step is None:
                    raise UsageError('field "%s": "step" must be a number'
                                     % s['name'])
                s['step'] = step
            options[field].append(s)


def dnQueryConfig(options):
    """bin/dn:696-724."""
    qconf = {'breakdowns': options['breakdowns']}
    if options.get('after'):
        qconf['timeAfter'] = options['after']
    if options.get('before'):
        qconf['timeBefore'] = options['before']
    if options.get('filter'):
        qconf['filter'] = options['filter']
    qc = mod_query.queryLoad({'query': qconf})
    if isinstance(qc, Exception):
        raise FatalError(str(qc))
    if options.get('gnuplot') and len(qc.qc_breakdowns) != 1:
        raise FatalError('--gnuplot can only be used with exactly one '
                         'breakdown')
    return {'query': qc, 'dryRun': bool(options.get('dry_run'))}


# ---------------------------------------------------------------------------
# number / value formatting (JS semantics for what the printers print)
# ---------------------------------------------------------------------------

def js_number_str(v):
    """String(number) for the magnitudes the printers meet."""
    if isinstance(v, int):
        return str(v)
    if v != v:
        return 'NaN'
    if math.isinf(v):
        return 'Infinity' if v > 0 else '-Infinity'
    if v == int(v) and abs(v) < 1e21:
        return str(int(v))
    r = repr(v)
    if 'e' not in r:
        return r
    mant, exp = r.split('e')
    e = int(exp)
    if -7 <= e < 21:
        from decimal import Decimal
        return format(Decimal(r), 'f')
    return '%se%s%d' % (mant.rstrip('0').rstrip('.') if '.' in mant else mant,
                        '+' if e >= 0 else '-', abs(e))


def _json_str(s):
    return json.dumps(s, ensure_ascii=False)


def point_json(fields, value):
    """JSON.stringify of one skinner point (bin/dn:972-975)."""
    parts = []
    for name, v in fields:
        if isinstance(v, (bytes, str)):
            if isinstance(v, bytes):
                try:
                    v = v.decode('utf-8')
                except UnicodeDecodeError:
                    v = v.decode('latin-1')
            vs = _json_str(v)
        elif v != v or math.isinf(v):
            vs = 'null'
        else:
            vs = js_number_str(v)
        parts.append('%s:%s' % (_json_str(name), vs))
    return '{"fields":{%s},"value":%s}' % (','.join(parts),
                                           js_number_str(value))


# ---------------------------------------------------------------------------
# flattening (lib/skinner-flattener.js) and printers
# ---------------------------------------------------------------------------

def _p2_ordinal(x):
    if x != x:
        return x
    if x < 1:
        return 0
    if math.isinf(x):
        return x
    return math.frexp(x)[1]


def _ordinal(bucketizer, x):
    if bucketizer.kind == 'p2':
        return _p2_ordinal(x)
    q = x / bucketizer.step if bucketizer.step else float('nan')
    if q != q or math.isinf(q):
        return q
    return int(math.floor(q))


def flatten(query, points):
    """Points -> rows [k1, ..., kn, value]; quantized columns become bucket
    ORDINALS again (the flattener re-bucketizes the bucket minima, which is
    idempotent); with no breakdowns the single row is the bare total."""
    # skinner keeps nested objects keyed by successive decomposition values,
    # so rows sharing a prefix come out contiguous, in first-seen order.
    root = {}
    total = 0
    for fields, value in points:
        key = []
        for (name, v), b in zip(fields, query.qc_breakdowns):
            if b['name'] in query.qc_bucketizers:
                key.append(_ordinal(query.qc_bucketizers[b['name']], v))
            else:
                if isinstance(v, bytes):
                    try:
                        v = v.decode('utf-8')
                    except UnicodeDecodeError:
                        v = v.decode('latin-1')
                key.append(v)
        total += value
        node = root
        for k in key[:-1]:
            node = node.setdefault(k, {})
        if key:
            node[key[-1]] = node.get(key[-1], 0) + value
    if not query.qc_breakdowns:
        return [total] if points else []
    rows = []

    def walk(node, prefix, depth):
        for k, v in node.items():
            if depth == len(query.qc_breakdowns) - 1:
                rows.append(prefix + [k, v])
            else:
                walk(v, prefix + [k], depth + 1)
    walk(root, [], 0)
    return rows


def _locale_key(s):
    return (s.casefold(), s.swapcase())


def _sort_rows(rows):
    def key(row):
        return tuple(_locale_key(x) if isinstance(x, str) else
                     (float('inf') if x != x else x) for x in row)
    return sorted(rows, key=key)


def _expand_values(query, rows):
    coldefs = query.qc_breakdowns
    quantized = len(coldefs) > 0 and coldefs[-1].get('aggr')
    for j, c in enumerate(coldefs):
        if quantized and j == len(coldefs) - 1:
            continue
        if c['name'] in query.qc_bucketizers:
            bz = query.qc_bucketizers[c['name']]
            for row in rows:
                row[j] = bz.bucketMin(row[j])
        if 'date' in c:
            for row in rows:
                row[j] = jsdate.to_iso_string(row[j] * 1000)


def _emit_table(out, columns, rows):
    """mod_tab.emitTable: space-separated, padded columns, header row."""
    def fmt(vals):
        cells = []
        for c, v in zip(columns, vals):
            s = v if isinstance(v, str) else js_number_str(v)
            cells.append(s.rjust(c['width']) if c.get('align') == 'right'
                         else s.ljust(c['width']))
        return ' '.join(cells)
    out.append(fmt([c['label'] for c in columns]))
    for r in rows:
        out.append(fmt(r))


def output_pretty(query, rows):
    out = []
    _expand_values(query, rows)
    coldefs = query.qc_breakdowns
    quantized = len(coldefs) > 0 and coldefs[-1].get('aggr')
    if quantized:
        return _output_pretty_quantized(query, rows)
    cols = [{'label': c['name'].upper(), 'width': len(c['name'])}
            for c in coldefs]
    cols.append({'label': 'VALUE', 'width': 5, 'align': 'right'})
    if not rows:
        return ''
    if len(rows) == 1 and not isinstance(rows[0], list):
        rows[0] = [rows[0]]
    for row in rows:
        for j in range(len(coldefs)):
            if not isinstance(row[j], str):
                cols[j]['align'] = 'right'
            w = len(row[j] if isinstance(row[j], str)
                    else js_number_str(row[j]))
            cols[j]['width'] = max(cols[j]['width'], w)
        cols[-1]['width'] = max(cols[-1]['width'],
                                len(js_number_str(row[-1])))
    _emit_table(out, cols, _sort_rows(rows))
    return '\n'.join(out) + '\n'


def _output_pretty_quantized(query, rows):
    coldefs = query.qc_breakdowns
    qcol = coldefs[-1]
    bz = query.qc_bucketizers[qcol['name']]
    groups = []
    last = None
    distr = []
    n = len(coldefs)
    for row in rows:
        key = ', '.join(x if isinstance(x, str) else js_number_str(x)
                        for x in row[:n - 1]) + '\n'
        if distr and key != last:
            groups.append({'label': last, 'distr': distr})
        if key != last:
            last = key
            distr = []
        distr.append([row[n - 1], row[n]])
    if last is not None:
        groups.append({'label': last, 'distr': distr})
    groups.sort(key=lambda g: _locale_key(g['label']))
    out = []
    for i, g in enumerate(groups):
        if i != 0:
            out.append('\n')
        out.append(g['label'])
        out.append(_print_distribution(g['distr'], bz, 'date' in qcol))
    return ''.join(out)


def _print_distribution(distr, bz, asdate):
    out = []
    if asdate:
        out.append('          ')
    out.append('           ')
    out.append('value  ------------- Distribution ------------- count\n')
    if not distr:
        return ''.join(out)
    distr = sorted(distr, key=lambda d: d[0])
    total = sum(d[1] for d in distr)
    bi = distr[0][0] if distr[0][0] > 100 else 0
    di = 0
    while di < len(distr) + 1:
        if di == len(distr):
            count = 0
            di += 1
        elif distr[di][0] == bi:
            count = distr[di][1]
            di += 1
        else:
            count = 0
        normalized = int(math.floor(40.0 * count / total + 0.5))
        dots = '@' * normalized + ' ' * (40 - normalized)
        mn = bz.bucketMin(bi)
        if asdate:
            out.append('  %24s |%s %s\n' % (jsdate.to_iso_string(mn * 1000),
                                            dots, js_number_str(count)))
        else:
            out.append('%16s |%s %s\n' % (js_number_str(mn), dots,
                                          js_number_str(count)))
        bi += 1
    return ''.join(out)


def output_gnuplot(query, rows, title):
    coldefs = query.qc_breakdowns
    o = ['#\n',
         '# This is a GNUplot input file generated automatically\n',
         '# by the Dragnet "dn" command.  You can use it to create\n',
         '# a graph as a PNG image (as file "graph.png") using:\n',
         '#\n', '#     gnuplot < this_file > graph.png\n', '#\n',
         'set terminal png size 1200,600\n',
         'set title "' + title + '"\n']
    if 'date' in coldefs[0]:
        o += ['# Configure plots to use the x-axis as time.\n',
              'set xdata time;\n', 'set timefmt "%s";\n',
              'set format x "%m/%d\\n%H:%MZ"\n']
    o += ['# Add 10% padding at the top of the graph.\n',
          'set offsets graph 0, 0, 0.1, 0\n',
          '# The y-axis should always start at zero.\n',
          'set yrange [0:*]\n', 'set ylabel "Count"\n', 'set ytics\n']
    assert len(coldefs) == 1
    xquant = coldefs[0]['name'] in query.qc_bucketizers
    if xquant:
        o.append('plot "-" using 1:2 with linespoints title "Value"\n')
    else:
        o.append('plot "-" using (column(0)):2:xtic(1) with linespoints '
                 'title "Value"\n')
    for row in _sort_rows(rows):
        x = query.qc_bucketizers[coldefs[0]['name']].bucketMin(row[0]) \
            if xquant else row[0]
        o.append('\t%s %s\n' % (x if isinstance(x, str) else js_number_str(x),
                                js_number_str(row[1])))
    o.append('\te\n')
    return ''.join(o)


# stage names in pipeline order, as vsWalk visits them (bin/dn:911-916)
STAGE_ORDER = ['json parser', 'SkinnerAdapterStream', 'Datasource filter',
               'User filter', 'Datetime parser', 'Time filter', 'Aggregator']


def format_counters(counters, flattened):
    """vstream vsDumpCounters: '%-18s %-14s %6d' per non-zero counter, stages
    in pipeline order and counter names sorted within a stage."""
    lines = []
    stages = list(STAGE_ORDER)
    cs = dict(counters)
    if flattened and 'Aggregator' in cs:
        cs['Flattener'] = {'ninputs': cs['Aggregator'].get('noutputs', 0),
                           'noutputs': 1}
        stages.append('Flattener')
    for st in stages:
        for name in sorted(cs.get(st, {})):
            n = cs[st][name]
            if n:
                lines.append('%-18s %-14s %6d' % (st, name + ':', n))
    return ''.join(l + '\n' for l in lines)


def render_scan(query, options, points, title='datasource'):
    """What `dn scan` writes to stdout for these points."""
    if options.get('points'):
        return ''.join(point_json(f, v) + '\n' for f, v in points)
    rows = flatten(query, points)
    if options.get('raw'):
        return ''.join(json.dumps(r) + '\n' for r in rows)
    if options.get('gnuplot'):
        return output_gnuplot(query, rows, title)
    return output_pretty(query, rows)


def main(argv=None, datasources=None, out=None, err=None):
    """`dn scan [opts] DATASOURCE` against an in-memory datasource table:
    ``datasources`` maps name -> dsconfig dict (see datasource_gpu)."""
    from dragnet_b200 import datasource_gpu
    argv = list(sys.argv[1:] if argv is None else argv)
    out = out or sys.stdout
    err = err or sys.stderr
    if not argv or argv[0] != 'scan':
        err.write('dn: only the "scan" subcommand is provided by the GPU '
                  'path\n')
        return 2
    try:
        options = dnParseArgs(argv[1:])
        if len(options['_args']) < 1:
            raise UsageError('missing arguments')
        if len(options['_args']) > 1:
            raise UsageError('extra arguments')
    except UsageError as ex:
        err.write('dn: %s\n' % ex)
        err.write('usage: dn SUBCOMMAND [OPTIONS] ARGS\n')
        return 2
    dsname = options['_args'][0]
    try:
        if not datasources or dsname not in datasources:
            raise FatalError('unknown datasource: "%s"' % dsname)
        ds = datasource_gpu.datasourceForConfig(
            {'dsconfig': datasources[dsname],
             'findFiles': mod_find.find_files})
        if isinstance(ds, Exception):
            raise FatalError(str(ds))
        scanargs = dnQueryConfig(options)
        res = ds.scan(scanargs)
        if isinstance(res, Exception):
            raise FatalError(str(res))
    except FatalError as ex:
        err.write('dn: %s\n' % ex)
        return 1
    if scanargs['dryRun']:
        err.write('would scan files:\n')
        for p in res.files:
            err.write('    %s\n' % p)
        return 0
    out.write(render_scan(scanargs['query'], options, res.points, dsname))
    if options.get('counters'):
        err.write(format_counters(res.counters, not options.get('points')))
    ds.close()
    return 0
