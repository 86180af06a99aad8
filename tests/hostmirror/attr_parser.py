"""Breakdown attribute mini-language: ``name[attr=value,attr2],name2``.

Host-side mirror of the reference's ``attrsParse`` (lib/attr-parser.js:17-77),
including its quirk that a trailing bare name of exactly one character is
dropped (attr-parser.js:73 tests ``j < str.length - 1``).

Errors are *returned* (not raised), as the reference does: the caller checks
``isinstance(rv, Exception)``.
"""


def attrsParse(s):
    rv = []
    propname = None
    props = None
    j = 0
    n = len(s)
    for i in range(n):
        ch = s[i]
        if propname is None:
            if ch == ',':
                if i - j > 0:
                    rv.append({'name': s[j:i]})
                j = i + 1
            elif ch == '[':
                if i - j == 0:
                    return ValueError('missing field name')
                propname = s[j:i]
                props = {'name': propname}
                j = i + 1
            continue

        if ch == ',' or ch == ']':
            if i - j > 0:
                propdef = s[j:i]
                eq = propdef.find('=')
                if eq == -1:
                    props[propdef] = ''
                elif eq == 0:
                    return ValueError('missing attribute name')
                else:
                    props[propdef[:eq]] = propdef[eq + 1:]
            if ch == ']':
                rv.append(props)
                propname = None
                props = None
            j = i + 1

    if propname is not None:
        return ValueError('unexpected end of string')

    if j < n - 1:
        rv.append({'name': s[j:]})
    return rv
