"""CPU oracle (Python): restatement of dragnet's raw-data scan path.

TEST INFRASTRUCTURE ONLY.  Nothing under dragnet_b200/ may import this module;
only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg use
oracle/.  The product path is the CUDA library (dragnet_b200/csrc).

What it restates (reference file:line, all under /root/reference):
  * line splitting + JSON decode + point adapter ... lib/format-json.js:26-98
    (third-party lstream@0.0.4, vstream-json-parser@1.0.0 = JSON.parse per line)
  * datasource filter stage ........................ lib/datasource-file.js:154-163
  * krill predicate stage .......................... lib/krill-skinner-stream.js:29-52
    (third-party krill@^1.0.0: eq ne lt le gt ge and or, JS loose comparison)
  * synthetic date fields .......................... lib/stream-synthetic.js:37-85
  * time-bounds filter ............................. lib/dragnet-impl.js:94-125
  * stage order .................................... lib/stream-scan.js:56-86
  * aggregation (third-party skinner#dragnet) ...... lib/dragnet-impl.js:48-89,
    bucketizers lib/dragnet.js:52-71
  * dotted lookup (third-party jsprim@^1.3.0 pluck)  lib/stream-synthetic.js:47

The third-party modules are not vendored in the reference tree and node is not
installed, so their behaviour is restated from their published semantics
(plain ECMAScript operations) and PINNED by the reference's own golden outputs
(tests/dn/local/*.out, README tables): see tests/test_oracle_golden.py.  Parts
no reference test exercises are marked "unpinned" in DESIGN.md.

Input: plan dict (dragnet_b200.query.scan_plan) + iterable of byte chunks.
Output: (points, counters).  A point is (fields, value) where fields is a list
of (name, v); v is ``bytes`` (UTF-8 of the JS string) for discrete breakdowns
or a float (bucketMin(ordinal), may be nan/inf) for quantized ones.
"""

import json
import math
import re

UNDEF = type('Undefined', (), {'__repr__': lambda s: 'undefined'})()

JS_WS = (' \t\n\r\v\f\u00a0\u1680\u2000\u2001\u2002\u2003\u2004\u2005'
         '\u2006\u2007\u2008\u2009\u200a\u2028\u2029\u202f\u205f\u3000\ufeff')


# --------------------------------------------------------------------------
# ECMAScript primitives
# --------------------------------------------------------------------------

def js_number_to_string(v):
    """Number::toString(v) (ECMA-262 7.1.12.1), returned as str."""
    if v != v:
        return 'NaN'
    if v == 0:
        return '0'
    if math.isinf(v):
        return 'Infinity' if v > 0 else '-Infinity'
    sign = '-' if v < 0 else ''
    v = abs(v)
    r = repr(float(v))          # shortest round-trip digits (David Gay)
    mant, _, exp = r.partition('e')
    exp = int(exp) if exp else 0
    ip, _, fp = mant.partition('.')
    digits = ip + fp
    n = len(ip) + exp           # value = 0.<digits> * 10**n
    stripped = digits.lstrip('0')
    n -= len(digits) - len(stripped)
    digits = stripped.rstrip('0')
    k = len(digits)
    if k <= n <= 21:
        return sign + digits + '0' * (n - k)
    if 0 < n <= 21:
        return sign + digits[:n] + '.' + digits[n:]
    if -6 < n <= 0:
        return sign + '0.' + '0' * (-n) + digits
    e = n - 1
    es = ('+' if e >= 0 else '-') + str(abs(e))
    if k == 1:
        return sign + digits + 'e' + es
    return sign + digits[0] + '.' + digits[1:] + 'e' + es


_STRNUM = re.compile(r'^[+-]?(?:Infinity|(?:\d+\.?\d*(?:[eE][+-]?\d+)?|\.\d+(?:[eE][+-]?\d+)?))$')


def js_string_to_number(b):
    """ToNumber(string) (ECMA-262 7.1.3.1, ES2015 grammar)."""
    try:
        s = b.decode('utf-8')
    except UnicodeDecodeError:
        s = b.decode('latin-1')
    s = s.strip(JS_WS)
    if s == '':
        return 0.0
    if len(s) > 2 and s[0] == '0' and s[1] in 'xXoObB':
        base = {'x': 16, 'o': 8, 'b': 2}[s[1].lower()]
        body = s[2:]
        ok = {16: '0123456789abcdefABCDEF', 8: '01234567', 2: '01'}[base]
        if all(c in ok for c in body):
            return float(int(body, base))
        return float('nan')
    if not _STRNUM.match(s):
        return float('nan')
    if 'Infinity' in s:
        return float('-inf') if s[0] == '-' else float('inf')
    return float(s)


def js_to_string(v):
    """ToString(v) as UTF-8 bytes."""
    if v is UNDEF:
        return b'undefined'
    if v is None:
        return b'null'
    if v is True:
        return b'true'
    if v is False:
        return b'false'
    if isinstance(v, float):
        return js_number_to_string(v).encode()
    if isinstance(v, bytes):
        return v
    if isinstance(v, dict):
        return b'[object Object]'
    if isinstance(v, list):
        return b','.join(b'' if (x is None or x is UNDEF) else js_to_string(x)
                         for x in v)
    raise TypeError(v)


def js_to_number(v):
    if v is UNDEF:
        return float('nan')
    if v is None:
        return 0.0
    if v is True:
        return 1.0
    if v is False:
        return 0.0
    if isinstance(v, float):
        return v
    if isinstance(v, bytes):
        return js_string_to_number(v)
    return js_string_to_number(js_to_string(v))   # object -> ToPrimitive


def _utf16_units(b):
    try:
        return b.decode('utf-8', 'surrogatepass').encode('utf-16-be',
                                                          'surrogatepass')
    except UnicodeDecodeError:
        return None


def js_loose_eq(a, b):
    """a == b where b is a string / number / boolean constant."""
    if isinstance(a, (dict, list)):
        a = js_to_string(a)
    if a is None or a is UNDEF:
        return False                      # constants are never null/undefined
    if isinstance(a, bytes) and isinstance(b, bytes):
        return a == b
    x, y = js_to_number(a), js_to_number(b)
    return x == y


def js_relational(op, a, b):
    if isinstance(a, (dict, list)):
        a = js_to_string(a)
    if isinstance(a, bytes) and isinstance(b, bytes):
        ua, ub = _utf16_units(a), _utf16_units(b)
        if ua is None or ub is None:
            ua, ub = a, b
        x, y = ua, ub
    else:
        x, y = js_to_number(a), js_to_number(b)
        if x != x or y != y:
            return False
    if op == 'lt':
        return x < y
    if op == 'le':
        return x <= y
    if op == 'gt':
        return x > y
    return x >= y


# --------------------------------------------------------------------------
# JSON.parse (strict ECMA-404) on one line of bytes
# --------------------------------------------------------------------------

class InvalidJson(Exception):
    pass


def _reject_constant(name):
    raise InvalidJson(name)


def _to_bytes_tree(v, enc):
    if isinstance(v, str):
        return v.encode(enc, 'surrogatepass')
    if isinstance(v, dict):
        return {k.encode(enc, 'surrogatepass'): _to_bytes_tree(x, enc)
                for k, x in v.items()}
    if isinstance(v, list):
        return [_to_bytes_tree(x, enc) for x in v]
    if isinstance(v, int) and not isinstance(v, bool):
        return float(v)
    return v


def json_parse_line(line):
    """JSON.parse(line).  Strings come back as bytes (UTF-8; invalid UTF-8 in
    the input is preserved byte-for-byte -- documented deviation from node's
    U+FFFD replacement, unpinned)."""
    try:
        text = line.decode('utf-8')
        enc = 'utf-8'
    except UnicodeDecodeError:
        if b'\\u' in line:
            raise NotImplementedError('oracle(py): \\u escape in a line that '
                                      'is not valid UTF-8')
        text = line.decode('latin-1')
        enc = 'latin-1'
    try:
        v = json.loads(text, parse_constant=_reject_constant,
                       parse_int=float, parse_float=float)
    except (ValueError, RecursionError) as ex:
        raise InvalidJson(str(ex))
    return _to_bytes_tree(v, enc)


# --------------------------------------------------------------------------
# jsprim.pluck
# --------------------------------------------------------------------------

def _has_own(obj, key):
    if isinstance(obj, dict):
        return key in obj
    # arrays: canonical index strings and "length"
    if key == b'length':
        return True
    if re.match(rb'^(0|[1-9][0-9]*)$', key):
        return int(key) < len(obj)
    return False


def _get_own(obj, key):
    if isinstance(obj, dict):
        return obj[key]
    if key == b'length':
        return float(len(obj))
    return obj[int(key)]


def pluck(obj, key):
    """jsprim.pluck: whole key first, else split at the FIRST dot."""
    if not isinstance(obj, (dict, list)):
        return UNDEF
    if _has_own(obj, key):
        return _get_own(obj, key)
    i = key.find(b'.')
    if i == -1:
        return UNDEF
    k1 = key[:i]
    if not _has_own(obj, k1):
        return UNDEF
    return pluck(_get_own(obj, k1), key[i + 1:])


def pluck_point(point, key):
    """pluck on a point whose top-level object may carry synthetic fields
    assigned by the Datetime parser stage (``chunk.fields[name] = v``,
    lib/stream-synthetic.js:61,80): an own property that shadows the JSON."""
    fields = point['fields']
    synth = point.get('synth')
    if not synth or not isinstance(fields, (dict, list)):
        return pluck(fields, key)
    if key in synth:
        return synth[key]
    if _has_own(fields, key):
        return _get_own(fields, key)
    i = key.find(b'.')
    if i == -1:
        return UNDEF
    k1 = key[:i]
    if k1 in synth:
        return pluck(synth[k1], key[i + 1:])
    if not _has_own(fields, k1):
        return UNDEF
    return pluck(_get_own(fields, k1), key[i + 1:])


# --------------------------------------------------------------------------
# krill
# --------------------------------------------------------------------------

class EvalError(Exception):
    pass


def _const(c):
    if isinstance(c, str):
        return c.encode('utf-8', 'surrogatepass')
    if isinstance(c, bool):
        return c
    return float(c)


def krill_eval(pred, point):
    if not pred:
        return True
    key = next(iter(pred))
    if key == 'and':
        for sub in pred[key]:
            if not krill_eval(sub, point):
                return False
        return True
    if key == 'or':
        for sub in pred[key]:
            if krill_eval(sub, point):
                return True
        return False
    field, const = pred[key]
    val = pluck_point(point, field.encode('utf-8', 'surrogatepass'))
    if val is UNDEF:
        raise EvalError('no value for field "%s"' % field)
    c = _const(const)
    if key == 'eq':
        return js_loose_eq(val, c)
    if key == 'ne':
        return not js_loose_eq(val, c)
    return js_relational(key, val, c)


# --------------------------------------------------------------------------
# Date.parse (ISO format only; see dragnet_b200/jsdate.py for the grammar)
# --------------------------------------------------------------------------

_ISO = re.compile(
    rb'^([+-]\d{6}|\d{4})(?:-(\d{2})(?:-(\d{2}))?)?'
    rb'(?:T(\d{2}):(\d{2})(?::(\d{2})(?:\.(\d+))?)?(Z|[+-]\d{2}:\d{2})?)?$')


def _days_from_civil(y, m, d):
    y -= m <= 2
    era = y // 400
    yoe = y - era * 400
    doy = (153 * (m + (-3 if m > 2 else 9)) + 2) // 5 + d - 1
    doe = yoe * 365 + yoe // 4 - yoe // 100 + doy
    return era * 146097 + doe - 719468


class Unsupported(Exception):
    """Input whose reference behaviour this restatement does not cover."""


_MONTHS = (b'jan', b'feb', b'mar', b'apr', b'may', b'jun', b'jul', b'aug',
           b'sep', b'oct', b'nov', b'dec')


def date_maybe_legacy(b):
    """Does this (non-ISO) string look like one of the forms V8's legacy
    Date.parse exists for: a month name, two date separators between digits,
    a clock time?  Those are not restated, so they are not called NaN either
    (dragnet_b200/csrc/jsdate.cuh dng_date_maybe_legacy)."""
    if len(re.findall(rb'(?=[0-9][-/][0-9])', b)) >= 2:
        return True
    if re.search(rb'[0-9]:[0-9]', b):
        return True
    for m in re.finditer(rb'[A-Za-z]+', b):
        if m.group(0)[:3].lower() in _MONTHS and len(m.group(0)) >= 3:
            return True
    return False


def date_parse_ms(b):
    m = _ISO.match(b)
    if not m:
        return None
    ys, mo, dd, hh, mi, ss, frac, tz = m.groups()
    if ys == b'-000000':
        return None
    y = int(ys)
    mo = int(mo) if mo else 1
    dd = int(dd) if dd else 1
    # V8's parser takes any day 1..31 and lets MakeDay carry it over
    if not (1 <= mo <= 12) or not (1 <= dd <= 31):
        return None
    h = int(hh) if hh else 0
    mn = int(mi) if mi else 0
    sc = int(ss) if ss else 0
    ms = int((frac + b'00')[:3]) if frac else 0
    if h > 24 or mn > 59 or sc > 59 or (h == 24 and (mn or sc or ms)):
        return None
    t = _days_from_civil(y, mo, dd) * 86400000 + \
        ((h * 60 + mn) * 60 + sc) * 1000 + ms
    if tz and tz != b'Z':
        oh, om = int(tz[1:3]), int(tz[4:6])
        if oh > 23 or om > 59:
            return None
        off = (oh * 60 + om) * 60000
        t = t - off if tz[:1] == b'+' else t + off
    if abs(t) > 8640000000000000:
        return None
    return t


# --------------------------------------------------------------------------
# bucketizers
# --------------------------------------------------------------------------

def p2_ordinal(v):
    x = js_to_number(v)
    if x != x:
        return float('nan')
    if x < 1:
        return 0.0
    if math.isinf(x):
        return float('inf')
    m, e = math.frexp(x)       # x = m * 2**e, 0.5 <= m < 1  => floor(log2 x) = e-1
    return float(e)            # 1 + floor(log2 x)


def linear_ordinal(v, step):
    x = js_to_number(v)
    step = float(step)
    if step == 0:
        q = float('nan') if (x == 0 or x != x) else math.copysign(float('inf'), x) * math.copysign(1.0, step)
    else:
        q = x / step
    if q != q or math.isinf(q):
        return q
    return float(math.floor(q)) + 0.0


def p2_bucket_min(o):
    if o != o:
        return float('nan')
    if o == 0:
        return 0.0
    if math.isinf(o):
        return float('inf')
    try:
        return math.ldexp(1.0, int(o) - 1)
    except OverflowError:
        return float('inf')


def linear_bucket_min(o, step):
    if math.isinf(o) and step == 0:
        return float('nan')
    return o * float(step)


# --------------------------------------------------------------------------
# the scan
# --------------------------------------------------------------------------

def iter_lines(chunks):
    """lstream: split on '\\n'; every line (including empty ones) is emitted;
    a final unterminated non-empty remainder is emitted at end of input."""
    buf = b''
    for c in chunks:
        buf += c
        parts = buf.split(b'\n')
        buf = parts.pop()
        for p in parts:
            yield p
    if buf:
        yield buf


def _bump(counters, stage, name, n=1):
    if not n:
        return
    counters.setdefault(stage, {})
    counters[stage][name] = counters[stage].get(name, 0) + n


def _filter_stage(pred, stage, point, counters):
    _bump(counters, stage, 'ninputs')
    try:
        ok = krill_eval(pred, point)
    except EvalError:
        _bump(counters, stage, 'nfailedeval')
        return False
    if ok:
        _bump(counters, stage, 'noutputs')
        return True
    _bump(counters, stage, 'nfilteredout')
    return False


def scan(plan, chunks):
    counters = {}
    fmt = plan.get('format', 'json')
    breakdowns = plan['breakdowns']
    synthetic = plan.get('synthetic') or []
    bounds = plan.get('time_bounds')
    table = {}
    total = 0
    for line in iter_lines(chunks):
        _bump(counters, 'json parser', 'ninputs')
        try:
            obj = json_parse_line(line)
        except InvalidJson:
            _bump(counters, 'json parser', 'invalid json')
            continue
        _bump(counters, 'json parser', 'noutputs')
        if fmt == 'json':
            _bump(counters, 'SkinnerAdapterStream', 'ninputs')
            _bump(counters, 'SkinnerAdapterStream', 'noutputs')
            point = {'fields': obj, 'value': 1, 'synth': {}}
        else:
            # json-skinner (lib/format-json.js:55-73): the line IS the point.
            f = obj.get(b'fields') if isinstance(obj, dict) else None
            w = obj.get(b'value') if isinstance(obj, dict) else None
            if not isinstance(f, dict) or not isinstance(w, float) or \
                    isinstance(w, bool) or math.isinf(w) or w != int(w) or \
                    w < 0 or w > 2.0 ** 53:
                # the reference asserts here; we count and drop (unpinned)
                _bump(counters, 'json parser', 'invalid point')
                continue
            point = {'fields': f, 'value': int(w), 'synth': {}}

        if plan.get('ds_filter'):
            if not _filter_stage(plan['ds_filter'], 'Datasource filter',
                                 point, counters):
                continue
        if plan.get('filter'):
            if not _filter_stage(plan['filter'], 'User filter', point,
                                 counters):
                continue
        if synthetic:
            _bump(counters, 'Datetime parser', 'ninputs')
            nerrors = 0
            for sc in synthetic:
                val = pluck_point(point, sc['field'].encode('utf-8'))
                if val is UNDEF:
                    if nerrors == 0:
                        _bump(counters, 'Datetime parser', 'undef')
                    nerrors += 1
                    continue
                if isinstance(val, float):
                    parsed = val
                else:
                    text = js_to_string(val)
                    ms = date_parse_ms(text)
                    if ms is None and isinstance(val, (bytes, str)) and \
                            date_maybe_legacy(text):
                        # not in the ES5 format, but V8's legacy parser may
                        # well make a date of it: not restated, so not ours
                        # to call NaN (the CUDA path refuses such input too)
                        raise Unsupported('Date.parse(%r): outside the '
                                          'ISO format' % text)
                    if ms is None:
                        if nerrors == 0:
                            _bump(counters, 'Datetime parser', 'baddate')
                        nerrors += 1
                        continue
                    parsed = float(math.floor(ms / 1000.0))
                if isinstance(point['fields'], (dict, list)):
                    point['synth'][sc['name'].encode('utf-8')] = parsed
            if nerrors:
                continue
            _bump(counters, 'Datetime parser', 'noutputs')
        if bounds:
            pred = {'and': [{'ge': [bounds['field'], bounds['ge']]},
                            {'lt': [bounds['field'], bounds['lt']]}]}
            if not _filter_stage(pred, 'Time filter', point, counters):
                continue

        _bump(counters, 'Aggregator', 'ninputs')
        key = []
        for b in breakdowns:
            v = pluck_point(point, b['name'].encode('utf-8'))
            if b.get('aggr') == 'quantize':
                o = p2_ordinal(v)
                key.append(('n', 'nan' if o != o else o))
            elif b.get('aggr') == 'lquantize':
                o = linear_ordinal(v, b['step'])
                key.append(('n', 'nan' if o != o else o))
            else:
                key.append(('s', js_to_string(v)))
        key = tuple(key)
        table[key] = table.get(key, 0) + point['value']
        total += point['value']

    points = []
    if not breakdowns:
        points.append(([], total))
    else:
        for key, value in table.items():
            fields = []
            for b, (kind, k) in zip(breakdowns, key):
                if kind == 's':
                    fields.append((b['name'], k))
                else:
                    o = float('nan') if k == 'nan' else k
                    if b['aggr'] == 'quantize':
                        fields.append((b['name'], p2_bucket_min(o)))
                    else:
                        fields.append((b['name'],
                                       linear_bucket_min(o, b['step'])))
            points.append((fields, value))
    _bump(counters, 'Aggregator', 'noutputs', len(points))
    return points, counters


# --------------------------------------------------------------------------
# rendering helpers shared by the tests
# --------------------------------------------------------------------------

def js_json_string(b):
    """JSON.stringify of a JS string given as UTF-8 bytes -> str."""
    try:
        s = b.decode('utf-8', 'surrogatepass')
    except UnicodeDecodeError:
        s = b.decode('latin-1')
    out = ['"']
    for ch in s:
        o = ord(ch)
        if ch == '"':
            out.append('\\"')
        elif ch == '\\':
            out.append('\\\\')
        elif o < 0x20:
            out.append({8: '\\b', 9: '\\t', 10: '\\n', 12: '\\f',
                        13: '\\r'}.get(o, '\\u%04x' % o))
        elif 0xD800 <= o <= 0xDFFF:
            out.append('\\u%04x' % o)
        else:
            out.append(ch)
    out.append('"')
    return ''.join(out)


def point_to_json(point):
    """JSON.stringify({fields:{...}, value:N}) as `dn scan --points` prints it
    (bin/dn:972-975)."""
    fields, value = point
    parts = []
    for name, v in fields:
        if isinstance(v, bytes):
            vs = js_json_string(v)
        elif v != v or math.isinf(v):
            vs = 'null'
        else:
            vs = js_number_to_string(v)
        parts.append('%s:%s' % (js_json_string(name.encode('utf-8')), vs))
    return '{"fields":{%s},"value":%d}' % (','.join(parts), value)
