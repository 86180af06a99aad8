"""Edge-case inputs and query plans shared by the CPU (hostcheck) and GPU
parity tests.  Everything here is checked oracle-vs-implementation, so the
cases may go beyond what the reference's own tests exercise (those are in
tests/golden)."""

import json
import random

from hostmirror import dn as mod_dn
from dragnet_b200 import query as mod_query

# ---------------------------------------------------------------------------
# handcrafted lines: JSON grammar corners, duplicates, nesting, types
# ---------------------------------------------------------------------------
EDGE_LINES = [
    b'{"a":1}', b'{"a":1} ', b' {"a":1}', b'\t{"a" : 1 }\r', b'{"a":1}x',
    b'{"a":1,}', b'{,"a":1}', b'{"a" 1}', b'{"a":}', b'{"a"}', b'{a:1}',
    b"{'a':1}", b'', b' ', b'{}', b'[]', b'[1,2', b'{"a":[1,2}', b'{"a":{]}',
    b'123', b'"str"', b'null', b'true', b'false', b'nul', b'tru', b'True',
    b'-', b'-0', b'01', b'1.', b'.5', b'1e', b'1e+', b'1.5e3', b'1E5', b'0e0',
    b'-0.0', b'1.0', b'{"a":01}', b'{"a":1.}', b'{"a":+1}', b'{"a":0x10}',
    b'{"a":NaN}', b'{"a":Infinity}', b'{"a":-Infinity}', b'{"a":1e999}',
    b'{"a":-1e999}', b'{"a":1e-999}',
    b'{"a":"\\u0041"}', b'{"a":"\\u004"}', b'{"a":"\\uZZZZ"}', b'{"a":"\\x41"}',
    b'{"a":"\\"}', b'{"a":"\\\\"}', b'{"a":"\\/"}', b'{"a":"a\\nb"}',
    b'{"a":"tab\there"}', b'{"a":"nul\x00"}', b'{"a":"del\x7f"}',
    b'{"a":"\xc3\xa9"}', b'{"a":"\\u00e9"}', b'{"a":"\xe2\x82\xac"}',
    b'{"a":"\\ud83d\\ude00"}', b'{"a":"\xf0\x9f\x98\x80"}',
    b'{"\\u0061":5}', b'{"\\u0061":5,"a":6}', b'{"a":5,"\\u0061":6}',
    b'{"a":1,"a":2}', b'{"a":{"b":1},"a":{"c":2}}', b'{"a":{"b":1},"a":3}',
    b'{"a.b":1,"a":{"b":2}}', b'{"a":{"b":2},"a.b":1}', b'{"a":{"b.c":3}}',
    b'{"a":{"b":{"c":4}}}', b'{"a":{"b.c":3,"b":{"c":4}}}',
    b'{"a.b.c":5,"a":{"b.c":3,"b":{"c":4}}}', b'{"a":{"b":5}}',
    b'{"a":{"b":null}}', b'{"a":{"b":true}}', b'{"a":{"b":false}}',
    b'{"a":{"b":"x"}}', b'{"a":{"b":[1,2]}}', b'{"a":{"b":{}}}',
    b'{"a":{"b":[]}}', b'{"a":{"b":[[1,2],[3]]}}', b'{"a":{"b":[null,1]}}',
    b'{"a":{"b":[{"x":1},"s",true,false,null,1.50,"q\\"r"]}}',
    b'{"a":{"b":[" 7 "]}}', b'{"a":[{"b":1}]}', b'{"a":"str","b":2}',
    b'{"a":5,"b":"5"}', b'{"a":"5","b":5}', b'{"a":"05"}', b'{"a":" 5 "}',
    b'{"a":"5e0"}', b'{"a":"0x5"}', b'{"a":""}', b'{"a":"abc"}', b'{"a":"ABC"}',
    b'{"a":200}', b'{"a":200.0}', b'{"a":2e2}', b'{"a":2.0e2}', b'{"a":"200"}',
    b'{"a":1e21}', b'{"a":1e-7}', b'{"a":123456789012345678901234567890}',
    b'{"a":0.1}', b'{"a":0.30000000000000004}', b'{"a":-5}', b'{"a":-0}',
    b'{"a":4.35}', b'{"a":9007199254740993}', b'{"a":1.7976931348623157e308}',
    b'{"a":5e-324}', b'{"a":0.000001}', b'{"a":1234.5678}', b'{"a":100000000000000000000}',
    b'{"b":1}', b'{"x":{"a":1}}', b'[{"a":1}]', b'{"a":1,"b":{"a":2}}',
    b'{"t":"2014-05-01T00:00:00.000Z","a":1}', b'{"t":"2014-05-01","a":1}',
    b'{"t":"2014-05-01T10:20","a":1}', b'{"t":"2014-05-01T10:20:30+02:00","a":1}',
    b'{"t":1398902400,"a":1}', b'{"t":1398902400.5,"a":1}', b'{"t":"nope","a":1}',
    b'{"t":null,"a":1}', b'{"t":true,"a":1}', b'{"t":{},"a":1}',
    b'{"t":["2014-05-01"],"a":1}', b'{"t":"\\u0032014-05-01","a":1}',
    b'{"t":"2014-02-30","a":1}', b'{"a":1,"t":"1969-12-31T23:59:59.500Z"}',
    b'{"fields":{"a":1},"value":3}', b'{"fields":{"a":1},"value":0}',
    b'{"fields":{"a":"x","b":2},"value":7}', b'{"fields":{},"value":1}',
    b'{"fields":{"a":1},"value":-1}', b'{"fields":{"a":1},"value":1.5}',
    b'{"fields":{"a":1},"value":"3"}', b'{"fields":[1],"value":3}',
    b'{"fields":{"a":1}}', b'{"value":3}', b'{"fields":{"a.b":9,"a":{"b":8}},"value":2}',
    b'{"a":' + b'[' * 40 + b']' * 40 + b'}',
    b'{"a":"' + b'x' * 300 + b'"}', b'{"k' + b'y' * 100 + b'":1,"a":2}',
    b'{"a":1,"pad":"' + b'p' * 5000 + b'"}',
    b'{"pad":"' + b'p' * 9000 + b'","a":"late"}',
    b'{"a":"\xff\xfe"}', b'{"a":"\xc3"}',
    b'{"a": [ 1 , 2 , { "b" : [ ] } ] , "b" : { } }',
    b'{"a":"x"}{"a":"y"}', b'{"a":"x"} {"a":"y"}', b'[1,2,3]', b'[[[]]]',
    b'{"":1}', b'{"":{"":2}}', b'{"a":{"":3}}', b'{"a.":4}', b'{".a":5}',
    b'{"a..b":6}', b'{"a":{"":{"b":7}}}',
]

SKINNER_LINES = [l for l in EDGE_LINES if b'"value"' in l or b'fields' in l] + [
    b'{"fields":{"req.method":"GET","x":1},"value":5}',
    b'{"fields":{"req":{"method":"PUT"}},"value":2}',
    b'not json', b'', b'{"fields":{"req.method":"GET"},"value":1}',
]


def rand_json(rng, depth=0):
    r = rng.random()
    if depth > 3 or r < 0.35:
        c = rng.randrange(9)
        if c == 0:
            return None
        if c == 1:
            return rng.random() < 0.5
        if c == 2:
            return rng.randrange(-50, 5000)
        if c == 3:
            return rng.choice([0.5, 1.25, 1e21, 1e-7, 123.456, -0.0, 2e2,
                               1.5e300, 4.35, 0.1])
        if c == 4:
            return rng.choice(['GET', 'PUT', 'a b', '', '200', ' 12 ', '1e3',
                               'x"y', 'tab\t', 'é', '€uro', '\U0001F600',
                               'null', 'true', '2014-05-01T00:00:00Z', '0x1f',
                               'Infinity', '-5', '5.0'])
        return rng.choice(['GET', 'HEAD', 'PUT', 'DELETE', 'x', 'y'])
    if r < 0.7:
        n = rng.randrange(0, 5)
        keys = ['a', 'b', 'c', 'a.b', 'b.c', 'req', 'method', 'x', '', 'length',
                '0', 't']
        return {rng.choice(keys): rand_json(rng, depth + 1) for _ in range(n)}
    return [rand_json(rng, depth + 1) for _ in range(rng.randrange(0, 4))]


def _dumps(rng, v):
    """JSON text with random (legal) formatting: whitespace, escapes,
    exponent spellings, duplicate keys."""
    if isinstance(v, dict):
        parts = []
        for k, x in v.items():
            ks = json.dumps(k)
            if k and rng.random() < 0.15:
                ks = '"' + ''.join('\\u%04x' % ord(ch) for ch in k) + '"'
            parts.append(ks + rng.choice([':', ' : ', ': ']) + _dumps(rng, x))
            if rng.random() < 0.08:       # duplicate key, later wins
                parts.append(ks + ':' + _dumps(rng, rand_json(rng, 3)))
        return '{' + rng.choice([',', ' , ', ', ']).join(parts) + '}'
    if isinstance(v, list):
        return '[' + rng.choice([',', ' ,']).join(_dumps(rng, x) for x in v) + ']'
    if isinstance(v, float) and rng.random() < 0.3 and v == int(v) and abs(v) < 1e15:
        return rng.choice(['%d.0', '%de0', '%d.00E+0']) % int(v)
    s = json.dumps(v, ensure_ascii=rng.random() < 0.5)
    return s


def random_lines(seed, n):
    rng = random.Random(seed)
    out = []
    for _ in range(n):
        v = rand_json(rng)
        if rng.random() < 0.8 and not isinstance(v, dict):
            v = {'a': v, 'b': rand_json(rng), 'req': {'method': rand_json(rng, 3)}}
        text = _dumps(rng, v)
        r = rng.random()
        if r < 0.04:
            text = text[:rng.randrange(0, max(1, len(text)))]   # truncated
        elif r < 0.06:
            text += rng.choice(['x', ',', '}', ' 1'])
        out.append(text.encode('utf-8'))
    return out


# ---------------------------------------------------------------------------
# query plans
# ---------------------------------------------------------------------------

def make_plan(argv, ds=None):
    """`dn scan ARGV` against datasource properties ds -> plan dict."""
    ds = ds or {}
    options = mod_dn.dnParseArgs(list(argv))
    q = mod_dn.dnQueryConfig(options)['query']
    return mod_query.scan_plan(q, ds_filter=ds.get('filter'),
                               time_field=ds.get('timeField'),
                               data_format=ds.get('dataFormat', 'json'))


EDGE_QUERIES = [
    ([], None),
    (['-b', 'a'], None),
    (['-b', 'a.b'], None),
    (['-b', 'a.b.c'], None),
    (['-b', 'a,b'], None),
    (['-b', 'a[aggr=quantize]'], None),
    (['-b', 'a[aggr=lquantize,step=100]'], None),
    (['-b', 'b[aggr=lquantize,step=7],a'], None),
    (['-b', 'a.b[aggr=quantize],req.method'], None),
    (['-f', '{"eq":["a",200]}'], None),
    (['-f', '{"eq":["a","200"]}', '-b', 'a'], None),
    (['-f', '{"ne":["a",5]}', '-b', 'a'], None),
    (['-f', '{"lt":["a",100]}', '-b', 'a'], None),
    (['-f', '{"ge":["a","abc"]}', '-b', 'a'], None),
    (['-f', '{"le":["a","5"]}', '-b', 'a'], None),
    (['-f', '{"gt":["a.b",1]}', '-b', 'a.b'], None),
    (['-f', '{"eq":["a",true]}', '-b', 'a'], None),
    (['-f', '{"eq":["a.b","1,2"]}', '-b', 'a.b'], None),
    (['-f', '{"eq":["a","[object Object]"]}'], None),
    (['-f', '{"and":[{"ge":["a",1]},{"lt":["a",300]}]}', '-b', 'a'], None),
    (['-f', '{"or":[{"eq":["a","x"]},{"eq":["b",2]}]}', '-b', 'a,b'], None),
    (['-f', '{"or":[{"and":[{"eq":["a",1]},{"eq":["b",1]}]},{"gt":["a",100]}]}'],
     None),
    (['-f', '{"eq":["req.method","GET"]}', '-b', 'req.method'],
     {'filter': {'ne': ['a', 1]}}),
    (['-b', 'ts[date,field=t,aggr=lquantize,step=86400]'], None),
    (['-b', 'ts[date,field=t],a'], None),
    (['-b', 'a', '--after', '2014-04-30', '--before', '2014-05-02'],
     {'timeField': 't'}),
    (['-b', 't[date],a'], None),
    (['-b', 'a[date,field=t]'], None),
    (['-b', '[aggr=quantize]x'], None) if False else (['-b', 'x'], None),
    (['-b', ''], None) if False else (['-b', 'length'], None),
    (['-b', 'a.0'], None),
    (['-b', 'a.b.1,a.b.length'], None),
    (['-b', '0.a,a.length'], None),
    (['-b', 'a.b.0.1'], None),
    (['-f', '{"eq":["a.b.length",2]}', '-b', 'a.b'], None),
    (['-b', 'a.b.0.x'], None),
]

SKINNER_QUERIES = [
    ([], {'dataFormat': 'json-skinner'}),
    (['-b', 'a'], {'dataFormat': 'json-skinner'}),
    (['-b', 'req.method'], {'dataFormat': 'json-skinner'}),
    (['-b', 'a.b,b'], {'dataFormat': 'json-skinner'}),
    (['-f', '{"eq":["a",1]}', '-b', 'a'], {'dataFormat': 'json-skinner'}),
]

BASELINE_QUERIES = {
    'C2': (['-b', 'req.method'], None),
    'C3': (['-b', 'req.method,res.statusCode', '-f',
            '{"eq":["req.method","GET"]}'], None),
    'C4': (['-b', 'latency[aggr=quantize]'], None),
    'C5': (['-b', 'operation,req.method,host'], None),
    'count': ([], None),
    'date': (['-b', 'ts[date,field=time,aggr=lquantize,step=3600],'
              'req.method'], None),
    'bounds': (['-b', 'host', '--after', '2014-05-31T22:00:00Z', '--before',
                '2014-05-31T23:00:00Z'], {'timeField': 'time'}),
    'caller': (['-b', 'req.caller,res.statusCode', '-f',
                '{"ne":["req.caller","admin"]}'], None),
    'lq': (['-b', 'dataLatency[aggr=lquantize,step=100],host'], None),
    'url': (['-b', 'req.url'], None),
    'urlhost': (['-b', 'req.url,host'], None),
}


# scalar forms behind one skeleton ({"a":<bare>,"s":"<string>"}): what the
# template matcher's wildcard scans must accept, reject or hand to the automaton
BARE_FORMS = [b'0', b'-0', b'1', b'12', b'123', b'1234', b'12345', b'123456',
              b'1234567', b'12345678', b'123456789', b'123456789012345',
              b'1234567890123456', b'12345678901234567890', b'-1', b'-12',
              b'-123', b'-1234', b'-12345', b'-123456789012345',
              b'-1234567890123456', b'0.5', b'-0.5', b'1e5', b'1E+5', b'1e-5',
              b'1.5e3', b'1.0', b'100.000', b'0e0', b'0.0', b'-0.0', b'1e400',
              b'-1e400', b'1e-400', b'01', b'-01', b'00', b'1.', b'.5', b'1e',
              b'1e+', b'-', b'+1', b'0x10', b'1a', b'1 ', b' 1', b'true',
              b'false', b'null', b'tru', b'nul', b'falsy', b'truee', b'nulll',
              b'True', b'NaN', b'Infinity', b'-Infinity', b'1,', b'1}',
              b'"x"', b'{}', b'[]', b'[1]', b'{"b":1}', b'']
STR_FORMS = [b'', b'a', b'ab', b'abc', b'abcd', b'abcde', b'abcdefgh',
             b'abcdefghi', b'x' * 63, b'x' * 64, b'x' * 65, b'\\n', b'a\\"b',
             b'\\\\', b'a\\u0041', b'\x01', b'a\tb', b'\x7f', b'\xc3\xa9',
             b'\xf0\x9f\x98\x80', b'\xff\xfe', b'a"', b'"', b'a\\']


def scalar_lines():
    out = []
    for i, f in enumerate(BARE_FORMS):
        out.append(b'{"a":' + f + b',"s":"k%d"}' % (i % 5))
    for i, f in enumerate(STR_FORMS):
        out.append(b'{"a":%d,"s":"' % (i % 7) + f + b'"}')
    return out


# ---------------------------------------------------------------------------
# template fuzz: a few record shapes, every line one of them with re-rolled
# scalar values (nasty strings, every number form, bare-scalar type flips) and
# a sprinkle of damage -- what the template matcher sees in the field
# ---------------------------------------------------------------------------

def _reroll(rng, v):
    if isinstance(v, dict):
        return {k: _reroll(rng, x) for k, x in v.items()}
    if isinstance(v, list):
        return [_reroll(rng, x) for x in v]
    if isinstance(v, str):
        return rng.choice(['GET', 'PUT', 'x', '', 'a b', '\u00e9', 'x"y', 'tab\t',
                           '\U0001F600', '2014-05-01T00:00:00Z', '200', ' 12 ',
                           '\\', '/', '\u20acuro' * 3, 'long' * 9])
    return rng.choice([None, True, False, 0, -0.0, 7, -12, 123456789012345,
                       1234567890123456, 0.5, 1e21, 1e-7, 1.5e300, 2e2, 4.35,
                       100])


def _dumps_shape(v, fmt, rng):
    """Formatting (separators) is a property of the shape, not of the line."""
    if isinstance(v, dict):
        sep, col = fmt.choice([',', ', ']), fmt.choice([':', ': '])
        return '{' + sep.join(json.dumps(k) + col + _dumps_shape(x, fmt, rng)
                              for k, x in v.items()) + '}'
    if isinstance(v, list):
        return '[' + ','.join(_dumps_shape(x, fmt, rng) for x in v) + ']'
    if isinstance(v, float) and rng.random() < 0.3 and v == int(v) and \
            abs(v) < 1e15:
        return rng.choice(['%d.0', '%de0', '%d.00E+0']) % int(v)
    return json.dumps(v, ensure_ascii=rng.random() < 0.5)


def template_fuzz_lines(seed, n, nshapes=6):
    rng = random.Random(seed)
    shapes = []
    while len(shapes) < nshapes:
        v = rand_json(rng)
        if isinstance(v, dict) and len(v) >= 1:
            shapes.append((v, rng.randrange(1 << 30)))
    out = []
    for _ in range(n):
        v, fmt = rng.choice(shapes)
        text = _dumps_shape(_reroll(rng, v), random.Random(fmt), rng)
        r = rng.random()
        if r < 0.03:
            text = text[:rng.randrange(0, max(1, len(text)))]
        elif r < 0.05:
            text += rng.choice(['x', ',', '}', ' 1', ' '])
        elif r < 0.07:
            text = text.replace(':', ' :', 1)
        out.append(text.encode('utf-8'))
    return out
