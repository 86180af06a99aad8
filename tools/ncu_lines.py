#!/usr/bin/env python
"""Aggregate an ncu report's per-SASS-instruction counters by CUDA source line
(needs -lineinfo builds and the matching libdragnet_gpu.so).

  python tools/ncu_lines.py gpurun_out/prof.ncu-rep [N]
"""
import collections
import csv
import glob
import os
import re
import subprocess
import sys
import tempfile

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(root, 'dragnet_b200', 'libdragnet_gpu.so')
d = tempfile.mkdtemp()
subprocess.run(['cuobjdump', '-xelf', 'all', lib], cwd=d, capture_output=True)
addr2line = {}
for cub in glob.glob(os.path.join(d, 'api*.cubin')):
    dis = subprocess.run(['nvdisasm', '-g', '-c', cub], capture_output=True,
                         text=True, errors='replace').stdout
    fn = fl = ln = None
    for l in dis.splitlines():
        m = re.match(r'^//-+ \.text\.(\S+)', l) or \
            re.match(r'^\s*\.section\s+\.text\.(\S+?),', l)
        if m:
            fn = m.group(1)
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', l)
        if m:
            fl, ln = m.group(1).split('/')[-1], int(m.group(2))
        m = re.match(r'^\s*/\*([0-9a-f]{4,})\*/\s+(\S.*?);', l)
        want = os.environ.get('NCU_FN', 'scan_kernel_w')
        if m and fn and ((want + 'ENS_8ScanArgs') in fn or
                         ('fILi' in want and want in fn)):
            addr2line[int(m.group(1), 16)] = (fl, ln)
skip = os.environ.get('NCU_SKIP', '0')
raw = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv',
                      '--launch-skip', skip, '--launch-count', '1'],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
h = next(i for i, r in enumerate(rows[:5]) if 'Instructions Executed' in r)
hdr = rows[h]
ci, cs, ca = hdr.index('Instructions Executed'), hdr.index('# Samples'), \
    hdr.index('Address')
cw, cwi = hdr.index('L1 Wavefronts Shared'), \
    hdr.index('L1 Wavefronts Shared Ideal')
base = None
ct = hdr.index('Thread Instructions Executed')
agg = collections.defaultdict(lambda: [0, 0, 0, 0, 0])
tot = tots = 0
for r in rows[h + 1:]:
    try:
        a = int(r[ca], 16) if r[ca].startswith('0x') else int(r[ca])
    except ValueError:
        continue
    if base is None:
        base = a
    k = addr2line.get(a - base, ('?', 0))
    v = agg[k]
    v[0] += float(r[ci] or 0)
    v[1] += float(r[cs] or 0)
    v[2] += float(r[cw] or 0)
    v[3] += float(r[cwi] or 0)
    v[4] += float(r[ct] or 0)
    tot += float(r[ci] or 0)
    tots += float(r[cs] or 0)
print('total warp-instructions %.0f, samples %.0f' % (tot, tots))
byfile = collections.defaultdict(lambda: [0, 0])
for k, v in agg.items():
    byfile[k[0]][0] += v[0]
    byfile[k[0]][1] += v[1]
for f, v in sorted(byfile.items(), key=lambda kv: -kv[1][1]):
    print('  %-22s inst %5.1f%%  samples %5.1f%%' % (f, 100 * v[0] / tot, 100 * v[1] / tots))
key = (lambda kv: -kv[1][0]) if os.environ.get('NCU_BY') == 'inst' else \
    (lambda kv: -kv[1][1])
for k, v in sorted(agg.items(), key=key)[:top]:
    print('%-22s:%4d  inst %5.1f%%  thr/inst %4.1f  samples %5.1f%%  smem '
          'wavefronts %10d (ideal %9d)' % (k[0], k[1], 100 * v[0] / tot,
                                           v[4] / max(v[0], 1),
                                           100 * v[1] / tots, v[2], v[3]))
