"""Input enumeration for the file-backed GPU datasource (test infrastructure:
the real integration keeps the reference's own scanInit enumeration and hands
the file list to the addon, integration/datasource-gpu.js).

Host-side mirror of ``DatasourceFile.findStream`` (lib/datasource-file.js:218-246):
a recursive, name-sorted file walk (lib/fs-find.js) optionally pruned by a
strftime-like ``timeFormat`` (%Y %m %d %H) expanded over [after, before)
(lib/path-enum.js:64-265: align the start down to the smallest unit used in the
pattern, then step by that unit while < end).  SURVEY.md section 8(f) rank 4.
"""

import os
import stat

from dragnet_b200 import jsdate


def path_enumerate(pattern, start_ms, end_ms):
    """All distinct expansions of ``pattern`` for times in [start, end)."""
    units = {'Y': 4, 'm': 3, 'd': 2, 'H': 1}
    parts = []
    i = 0
    minunit = None
    while i < len(pattern):
        if pattern[i] == '%' and i + 1 < len(pattern):
            k = pattern[i + 1]
            if k not in units:
                raise ValueError('unsupported conversion specifier: "%%%s"'
                                 % k)
            parts.append(('conv', k))
            if minunit is None or units[k] < units[minunit]:
                minunit = k
            i += 2
        else:
            j = pattern.find('%', i)
            j = len(pattern) if j == -1 else j
            if j == i:
                j = i + 1
            parts.append(('str', pattern[i:j]))
            i = j

    def fields(ms):
        iso = jsdate.to_iso_string(ms)
        return int(iso[0:4]), int(iso[5:7]), int(iso[8:10]), int(iso[11:13])

    def mk(y, mo, d, h):
        return (jsdate.days_from_civil(y, mo, d) * 24 + h) * 3600000

    y, mo, d, h = fields(start_ms)
    if minunit == 'Y':
        mo, d, h = 1, 1, 0
    elif minunit == 'm':
        d, h = 1, 0
    elif minunit == 'd':
        h = 0
    out = []
    cur = mk(y, mo, d, h)
    while True:
        y, mo, d, h = fields(cur)
        s = ''.join(v if kind == 'str' else
                    ('%d' % y if v == 'Y' else
                     '%02d' % {'m': mo, 'd': d, 'H': h}[v])
                    for kind, v in parts)
        out.append(s)
        if minunit is None:
            break
        if minunit == 'Y':
            cur = mk(y + 1, mo, d, h)
        elif minunit == 'm':
            cur = mk(y + (mo == 12), 1 if mo == 12 else mo + 1, d, h)
        elif minunit == 'd':
            cur += 86400000
        else:
            cur += 3600000
        if cur >= end_ms:
            break
    return out


def _walk(path, out):
    try:
        st = os.stat(path)
    except OSError:
        return
    if stat.S_ISDIR(st.st_mode):
        for name in sorted(os.listdir(path)):
            _walk(os.path.join(path, name), out)
    elif stat.S_ISREG(st.st_mode) or stat.S_ISCHR(st.st_mode):
        out.append(path)


def find_files(root, time_format=None, after_ms=None, before_ms=None):
    """Files a scan of this datasource reads, in traversal order."""
    out = []
    if before_ms is None or time_format is None:
        _walk(root, out)
        return out
    for sub in path_enumerate(os.path.join(root, time_format),
                              after_ms, before_ms):
        _walk(sub, out)
    return out
