"""Host-side mirror of the krill predicate language as dragnet uses it.

The reference validates filters with ``krill.createPredicate`` (third-party
npm module ``krill@^1.0.0``, not vendored; call sites lib/dragnet.js:113-121,
lib/stream-scan.js:57,76, lib/datasource-file.js:155).  Only validation and
field enumeration live on the host; *evaluation* happens on the GPU
(csrc/scan_kernel.cuh) and, for tests, in oracle/.

Syntax: ``{}`` (always true), ``{op: [field, constant]}`` for
op in eq ne lt le gt ge, ``{and: [p, ...]}``, ``{or: [p, ...]}``.
"""

REL_OPS = ('eq', 'ne', 'lt', 'le', 'gt', 'ge')
LOGICAL_OPS = ('and', 'or')


class KrillError(ValueError):
    pass


def _inspect(v, depth=0):
    """Small restatement of node's util.inspect for plain JSON values (used
    only to reproduce krill's error text, tests/dn/local/tst.badargs.sh.out:9)."""
    if isinstance(v, dict):
        if not v:
            return '{}'
        if depth > 2:
            return '[Object]'
        parts = []
        for k, x in v.items():
            ks = k if k.isidentifier() else "'%s'" % k
            parts.append('%s: %s' % (ks, _inspect(x, depth + 1)))
        return '{ ' + ', '.join(parts) + ' }'
    if isinstance(v, list):
        if not v:
            return '[]'
        if depth > 2:
            return '[Object]'
        return '[ ' + ', '.join(_inspect(x, depth + 1) for x in v) + ' ]'
    if isinstance(v, str):
        return "'" + v.replace("'", "\\'") + "'"
    if v is None:
        return 'null'
    if v is True:
        return 'true'
    if v is False:
        return 'false'
    if isinstance(v, float) and v == int(v):
        return str(int(v))
    return str(v)


def _validate(pred):
    if not isinstance(pred, dict):
        raise KrillError('predicate %s: not an object' % _inspect(pred))
    keys = list(pred.keys())
    if len(keys) == 0:
        return
    if len(keys) > 1:
        raise KrillError('predicate %s: expected exactly one key' %
                         _inspect(pred))
    key = keys[0]
    if key in LOGICAL_OPS:
        args = pred[key]
        if not isinstance(args, list):
            raise KrillError('predicate %s: "%s" must be an array' %
                             (_inspect(pred), key))
        if len(args) == 0:
            raise KrillError('predicate %s: "%s" requires at least one '
                             'subpredicate' % (_inspect(pred), key))
        for sub in args:
            _validate(sub)
        return
    if key in REL_OPS:
        args = pred[key]
        if not isinstance(args, list) or len(args) != 2:
            raise KrillError('predicate %s: "%s" requires exactly two '
                             'arguments' % (_inspect(pred), key))
        if not isinstance(args[0], str):
            raise KrillError('predicate %s: field name must be a string' %
                             _inspect(pred))
        c = args[1]
        if not (isinstance(c, (str, int, float)) or isinstance(c, bool)):
            raise KrillError('predicate %s: constant must be a string, '
                             'number or boolean' % _inspect(pred))
        return
    raise KrillError('predicate %s: unknown operator "%s"' %
                     (_inspect(pred), key))


class Predicate(object):
    def __init__(self, pred):
        _validate(pred)
        self.p_pred = pred

    def trivial(self):
        return len(self.p_pred) == 0

    def fields(self):
        out = []

        def walk(p):
            if not p:
                return
            key = next(iter(p))
            if key in LOGICAL_OPS:
                for s in p[key]:
                    walk(s)
            elif p[key][0] not in out:
                out.append(p[key][0])
        walk(self.p_pred)
        return out


def createPredicate(pred):
    """Validate ``pred``; raises KrillError like krill throws."""
    return Predicate(pred)
