"""Exact JS number/date semantics of the device code (csrc/jsnum.cuh,
jsdate.cuh), exercised on the host through tests/hostcheck/jsnum_check (a
test-only build) against Python's correctly rounded float()/repr() and the
oracle's restatements: decimal -> binary64 (Clinger + big-integer slow path),
Number::toString shortest round-trip digits and layout, StringToNumber,
Date.parse."""

import os
import random
import struct
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'oracle'))
from dn_oracle import (date_parse_ms, js_number_to_string,  # noqa: E402
                       js_string_to_number)

HC = os.path.join(ROOT, 'tests', 'hostcheck')


@pytest.fixture(scope='module')
def exe():
    out = os.path.join(HC, 'jsnum_check')
    src = os.path.join(HC, 'jsnum_check.cpp')
    deps = [src] + [os.path.join(ROOT, 'dragnet_b200', 'csrc', n)
                    for n in ('jsnum.cuh', 'jsdate.cuh')]
    if not os.path.exists(out) or os.path.getmtime(out) < max(
            os.path.getmtime(d) for d in deps):
        subprocess.check_call(['g++', '-std=c++17', '-O2', '-o', out, src])
    return out


def run(exe, lines):
    return subprocess.run([exe], input=''.join(l + '\n' for l in lines),
                          capture_output=True, text=True,
                          check=True).stdout.split('\n')


def bits(d):
    return '%016x' % struct.unpack('<Q', struct.pack('<d', d))[0]


def test_decimal_to_double_is_correctly_rounded(exe):
    rng = random.Random(1)
    dec = ['0', '-0', '1', '123', '1.5', '0.1', '1e21', '1e-7',
           '123456789012345678', '1234567890123456789012345', '5e-324',
           '2.4703282292062327e-324', '2.4703282292062328e-324',
           '1.7976931348623157e308', '1.7976931348623158e308',
           '1.7976931348623159e308', '1e309', '9007199254740993',
           '9007199254740992.5', '4.9e-324', '1e-400', '0.1e1', '100e-2',
           '1E5', '1e+5', '8.5', '0.30000000000000004',
           '2.2250738585072011e-308', '2.2250738585072014e-308',
           '9007199254740993.' + '0' * 40 + '1', '0.' + '0' * 400 + '1',
           '1' + '0' * 400, '3.' + '141592653589793' * 60]
    for _ in range(4000):
        k = rng.choice([1, 5, 10, 17, 18, 19, 20, 25, 40, 120])
        s = ''.join(rng.choice('0123456789') for _ in range(k)).lstrip('0') \
            or '0'
        if rng.random() < 0.6:
            p = rng.randint(0, len(s))
            s = (s[:p] or '0') + '.' + (s[p:] or '0')
        if rng.random() < 0.6:
            s += 'e%d' % rng.randint(-340, 320)
        dec.append(s)
    for _ in range(3000):
        d = struct.unpack('<d', struct.pack('<Q', rng.getrandbits(64)))[0]
        if d == d and abs(d) != float('inf'):
            dec.append(repr(d))
    out = run(exe, ['p ' + s for s in dec])
    bad = [(s, bits(float(s)), o) for s, o in zip(dec, out)
           if bits(float(s)) != o]
    assert not bad, bad[:5]


def test_number_to_string_matches_js(exe):
    rng = random.Random(2)
    vals = [0.0, 1.0, -1.5, 1e21, 1e-7, 123.456, 5e-324,
            1.7976931348623157e308, 0.1, 0.30000000000000004, 1e20, 2.0 ** 53,
            2.0 ** 53 + 2, 1 / 3, 1e-6, 1.5e-7, 123456789012345680000.0, 4.35,
            9.5e-7, float('inf'), float('-inf'), float('nan'),
            2.2250738585072014e-308, 1e23, 9.999999999999999e22,
            -1378331202551613.2]
    for _ in range(6000):
        d = struct.unpack('<d', struct.pack('<Q', rng.getrandbits(64)))[0]
        if d == d:
            vals.append(d)
    for _ in range(3000):
        vals.append(rng.uniform(-1e6, 1e6))
        vals.append(round(rng.uniform(0, 1000), rng.randint(0, 6)))
        vals.append(float(rng.randint(0, 2 ** 60)))
    out = run(exe, ['s ' + bits(v) for v in vals])
    bad = [(v, js_number_to_string(v), o) for v, o in zip(vals, out)
           if js_number_to_string(v) != o]
    assert not bad, bad[:5]


def test_string_to_number(exe):
    strs = ['', '  ', ' 12 ', '0x1f', '0X1F', '0b101', '0o17', '0x', '1e3',
            '.5', '5.', '.', '+.5e-3', '-Infinity', 'Infinity', '+Infinity',
            'infinity', '12px', '0x1g', '1e', '1e+', '--1', '+-1',
            '\xa0 7 ﻿', ' 7', '1_000', '0xffffffffffffffffffff',
            '0x1fffffffffffff8', '0x1fffffffffffff9', '26', '007', '-0',
            '1,2', 'null', 'true']
    out = run(exe, ['n ' + s.encode('utf-8').hex() for s in strs])
    for s, o in zip(strs, out):
        e = bits(js_string_to_number(s.encode('utf-8')))
        assert e == o or (e[:4] == '7ff8' and o[:4] in ('7ff8', 'fff8')), s


def test_date_parse(exe):
    rng = random.Random(3)
    dates = ['2014-05-01T00:00:00.000Z', '2014-05-01', '2014', '2014-05',
             '2014-05-02T04:05:06.123', '2014-05-02T04:05', '2014-02-29',
             '2016-02-29T23:59:59Z', 'invalid', '2014-05-01T00:00:00+01:00',
             '2014-05-01T00:00:00.5Z', '2014-05-01T24:00:00Z',
             '2014-05-01T24:00:01Z', '+002014-05-01', '-000001-01-01',
             '-000000-01-01', '2014-13-01', '2014-05-01T00:00:00.123456Z',
             '2014-05-01 00:00:00', '1969-12-31T23:59:59.999Z',
             '2014-05-01T00:00Z', '2014-5-1', '0000-01-01',
             '+275760-09-13T00:00:00.000Z', '+275760-09-13T00:00:00.001Z',
             '-271821-04-20T00:00:00Z', '-271821-04-19T00:00:00Z',
             '1900-02-29', '2000-02-29', '2100-02-29']
    for _ in range(3000):
        y = rng.randint(-3000, 9000)
        ys = '%04d' % y if 0 <= y <= 9999 else '%+07d' % y
        dates.append('%s-%02d-%02dT%02d:%02d:%02d.%03dZ' % (
            ys, rng.randint(1, 12), rng.randint(1, 31), rng.randint(0, 23),
            rng.randint(0, 59), rng.randint(0, 59), rng.randint(0, 999)))
    out = run(exe, ['d ' + s for s in dates])
    for s, o in zip(dates, out):
        e = date_parse_ms(s.encode())
        assert ('NaN' if e is None else str(e)) == o, s
