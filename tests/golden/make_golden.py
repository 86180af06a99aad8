#!/usr/bin/env python
"""Generate the committed golden fixtures from the reference tree.

Run HERE (the container that has /root/reference); the outputs travel with the
repo because /root/reference does not exist on the GPU box:

  tests/golden/data.tar.gz      the reference's test inputs
                                (tests/data/2014/05-0{1..5}/*.log, incl. the
                                hand-edited edge-case lines)
  tests/golden/scan_goldens.json  every "# dn scan ..." section of the
                                reference's golden outputs, split into
                                {suite, argv, text}:
                                  tests/dn/local/tst.scan_file.sh.out
                                  tests/dn/local/tst.scan_fileset.sh.out
                                  tests/dn/local/tst.empty.sh.out
                                  tests/dn/manta/tst.scan_manta.sh.out
                                plus tst.format_skinner.sh.out, tst.badargs.sh.out
                                and tst.scan_250k.sh.out verbatim.

Nothing here is reference *source*: these are its test vectors, which pin the
oracle (oracle/) and, through it, the CUDA path.
"""

import io
import json
import os
import re
import sys
import tarfile

REF = '/root/reference'
HERE = os.path.dirname(os.path.abspath(__file__))


def split_sections(text):
    """Sections start with '# dn scan ...' and run to the next such header."""
    sections = []
    cur = None
    for line in text.split('\n'):
        if line.startswith('# dn scan') or line.startswith('# dn query'):
            if cur:
                sections.append(cur)
            cur = {'header': line, 'lines': []}
        elif cur is not None:
            cur['lines'].append(line)
    if cur:
        sections.append(cur)
    return sections


def parse_argv(header):
    """Recover argv from the echoed (unquoted) command line."""
    rest = header[len('# dn '):]
    cmd, _, rest = rest.partition(' ')
    argv = []
    i = 0
    rest = rest.strip()
    dec = json.JSONDecoder()
    while i < len(rest):
        if rest[i] == ' ':
            i += 1
            continue
        j = rest.find(' ', i)
        tok = rest[i:] if j == -1 else rest[i:j]
        i = len(rest) if j == -1 else j + 1
        argv.append(tok)
        if tok in ('-f', '--filter'):
            while rest[i] == ' ':
                i += 1
            _, end = dec.raw_decode(rest, i)
            argv.append(rest[i:end])
            i = end
    return cmd, argv


def main():
    out = {'suites': {}}
    for suite, path in [
            ('scan_file', 'tests/dn/local/tst.scan_file.sh.out'),
            ('scan_fileset', 'tests/dn/local/tst.scan_fileset.sh.out'),
            ('empty', 'tests/dn/local/tst.empty.sh.out'),
            ('scan_manta', 'tests/dn/manta/tst.scan_manta.sh.out')]:
        text = open(os.path.join(REF, path)).read()
        secs = []
        for s in split_sections(text):
            cmd, argv = parse_argv(s['header'])
            body = '\n'.join(s['lines'])
            secs.append({'cmd': cmd, 'argv': argv, 'header': s['header'],
                         'text': body})
        out['suites'][suite] = secs
    for name, path in [
            ('format_skinner', 'tests/dn/local/tst.format_skinner.sh.out'),
            ('badargs', 'tests/dn/local/tst.badargs.sh.out'),
            ('scan_250k', 'tests/dn/local/tst.scan_250k.sh.out')]:
        out[name] = open(os.path.join(REF, path)).read()
    with open(os.path.join(HERE, 'scan_goldens.json'), 'w') as f:
        json.dump(out, f, indent=0, sort_keys=True)

    buf = io.BytesIO()
    with tarfile.open(fileobj=buf, mode='w:gz') as tf:
        root = os.path.join(REF, 'tests/data')
        for dp, dn, fn in sorted(os.walk(root)):
            dn.sort()
            for n in sorted(fn):
                full = os.path.join(dp, n)
                ti = tf.gettarinfo(full, arcname=os.path.relpath(full, root))
                ti.mtime = 0
                ti.uid = ti.gid = 0
                ti.uname = ti.gname = ''
                with open(full, 'rb') as fh:
                    tf.addfile(ti, fh)
    with open(os.path.join(HERE, 'data.tar.gz'), 'wb') as f:
        f.write(buf.getvalue())
    print('wrote', len(buf.getvalue()), 'bytes of data;',
          sum(len(v) for v in out['suites'].values()), 'sections')


if __name__ == '__main__':
    sys.exit(main())
