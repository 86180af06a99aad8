#!/usr/bin/env python
"""SASS of one kernel's own body (callees excluded) with source lines; no GPU.
  python tools/sass_body.py KERNEL_SUBSTRING [LIB]  > listing
  python tools/sass_body.py KERNEL_SUBSTRING --summary    per-line counts"""
import glob, os, re, subprocess, sys, tempfile, collections
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
want = sys.argv[1]
summary = '--summary' in sys.argv
args = [a for a in sys.argv[2:] if not a.startswith('--')]
lib = args[0] if args else os.path.join(root, 'dragnet_b200', 'libdragnet_gpu.so')
d = tempfile.mkdtemp()
subprocess.run(['cuobjdump', '-xelf', 'all', lib], cwd=d, capture_output=True)
rows = []
for cub in glob.glob(os.path.join(d, 'api*.cubin')):
    dis = subprocess.run(['nvdisasm', '-g', '-c', cub], capture_output=True,
                         text=True, errors='replace').stdout
    on = False
    fl = ln = None
    for l in dis.splitlines():
        m = re.match(r'^\s*\.section\s+\.text\.(\S+?),', l)
        if m:
            on = want in m.group(1)
            continue
        if on and re.match(r'^\s*\.type\s+\$', l):
            on = False          # first callee: the body is over
        if not on:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', l)
        if m:
            fl, ln = m.group(1).split('/')[-1], int(m.group(2))
        m = re.match(r'^\s*/\*([0-9a-f]{4,})\*/\s+(\S.*?);', l)
        if m:
            rows.append((m.group(1), fl, ln, m.group(2)))
        elif re.match(r'^\.L_x_\d+:', l) and not summary:
            rows.append((None, None, None, l))
if summary:
    c = collections.Counter((r[1], r[2]) for r in rows)
    f = collections.Counter(r[1] for r in rows)
    print('total', len(rows), dict(f))
    spill = collections.Counter((r[1], r[2]) for r in rows if re.search(r'\b(LDL|STL)', r[3]))
    print('LDL/STL', sum(spill.values()))
    for k, v in sorted(c.items(), key=lambda kv: (kv[0][0] or '', kv[0][1] or 0)):
        print('%-28s:%5s %5d %s' % (k[0], k[1], v, ('spill %d' % spill[k]) if spill[k] else ''))
else:
    for a, fl, ln, ins in rows:
        if a is None:
            print(ins)
        else:
            print('%s %-16s:%4d  %s' % (a, fl, ln or 0, ins))
