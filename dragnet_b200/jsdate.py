"""ECMAScript date-time string format (ES5 15.9.1.15), host side.

Used for ``--after/--before`` (reference: dashdash 'date' option type feeding
``new Date(x)`` in lib/dragnet.js:166-176).  The same grammar is implemented on
the device for record ``date`` fields (lib/stream-synthetic.js:65).

Accepted: ``YYYY``, ``YYYY-MM``, ``YYYY-MM-DD`` (also ``+YYYYYY``/``-YYYYYY``),
optionally followed by ``THH:mm``, ``THH:mm:ss`` or ``THH:mm:ss.s+`` and an
optional ``Z`` or ``+HH:mm``/``-HH:mm`` offset.  A missing offset means UTC
(the ES5 reading the reference's goldens were produced under; deliberately
independent of the host TZ).  Anything else is NaN -> ``None``.  V8's legacy
free-form date fallback parser is not restated (unpinned by any reference test).
"""

import re

_ISO = re.compile(
    r'^([+-]\d{6}|\d{4})(?:-(\d{2})(?:-(\d{2}))?)?'
    r'(?:T(\d{2}):(\d{2})(?::(\d{2})(?:\.(\d+))?)?(Z|[+-]\d{2}:\d{2})?)?$')


def days_from_civil(y, m, d):
    """Days since 1970-01-01 of a proleptic Gregorian date."""
    y -= m <= 2
    era = y // 400
    yoe = y - era * 400
    doy = (153 * (m + (-3 if m > 2 else 9)) + 2) // 5 + d - 1
    doe = yoe * 365 + yoe // 4 - yoe // 100 + doy
    return era * 146097 + doe - 719468


def _dim(y, m):
    if m == 2:
        return 29 if (y % 4 == 0 and (y % 100 != 0 or y % 400 == 0)) else 28
    return 30 if m in (4, 6, 9, 11) else 31


def date_parse_ms(s):
    """Date.parse(s) for the ISO format; returns integer ms or None (NaN)."""
    m = _ISO.match(s)
    if not m:
        return None
    ys, mo, dd, hh, mi, ss, frac, tz = m.groups()
    if ys == '-000000':
        return None
    y = int(ys)
    mo = int(mo) if mo else 1
    dd = int(dd) if dd else 1
    # (V8 accepts day 1..31 in any month and lets it carry over)
    if not (1 <= mo <= 12) or not (1 <= dd <= 31):
        return None
    h = int(hh) if hh else 0
    mi_ = int(mi) if mi else 0
    s_ = int(ss) if ss else 0
    ms = int((frac + '00')[:3]) if frac else 0
    if h > 24 or mi_ > 59 or s_ > 59:
        return None
    if h == 24 and (mi_ or s_ or ms):
        return None
    t = days_from_civil(y, mo, dd) * 86400000 + \
        ((h * 60 + mi_) * 60 + s_) * 1000 + ms
    if tz and tz != 'Z':
        oh, om = int(tz[1:3]), int(tz[4:6])
        if oh > 23 or om > 59:
            return None
        off = (oh * 60 + om) * 60000
        t = t - off if tz[0] == '+' else t + off
    if abs(t) > 8.64e15:
        return None
    return t


def to_iso_string(ms):
    """new Date(ms).toISOString() for in-range integral ms."""
    ms = int(ms)
    days, rem = divmod(ms, 86400000)
    z = days + 719468
    era = z // 146097
    doe = z - era * 146097
    yoe = (doe - doe // 1460 + doe // 36524 - doe // 146096) // 365
    y = yoe + era * 400
    doy = doe - (365 * yoe + yoe // 4 - yoe // 100)
    mp = (5 * doy + 2) // 153
    d = doy - (153 * mp + 2) // 5 + 1
    mth = mp + (3 if mp < 10 else -9)
    y += mth <= 2
    h, rem = divmod(rem, 3600000)
    mi, rem = divmod(rem, 60000)
    s, msec = divmod(rem, 1000)
    if 0 <= y <= 9999:
        ys = '%04d' % y
    else:
        ys = ('+' if y > 0 else '-') + '%06d' % abs(y)
    return '%s-%02d-%02dT%02d:%02d:%02d.%03dZ' % (ys, mth, d, h, mi, s, msec)
