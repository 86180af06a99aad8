#!/usr/bin/env python
"""Top SASS instructions of scan_kernel by stall samples, with source line and
the dominant stall reasons.   python tools/ncu_sass.py REPORT [N]"""
import csv, glob, os, re, subprocess, sys, tempfile
rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.environ.get('DNG_LIB') or os.path.join(root, 'dragnet_b200', 'libdragnet_gpu.so')
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
        if m and fn and (os.environ.get('NCU_FN', 'scan_kernel_w') + 'ENS_8ScanArgs') in fn:
            addr2line[int(m.group(1), 16)] = (fl, ln)
skip = os.environ.get('NCU_SKIP', '0')
raw = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv',
                      '--launch-skip', skip, '--launch-count', '1'],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
h = next(i for i, r in enumerate(rows[:5]) if 'Instructions Executed' in r)
hdr = rows[h]
ca, cs, csrc, ci = hdr.index('Address'), hdr.index('# Samples'), hdr.index('Source'), hdr.index('Instructions Executed')
stalls = [(i, n) for i, n in enumerate(hdr) if n.startswith('stall_') and 'Not Issued' not in n]
out = []
base = None
tot = 0
for r in rows[h + 1:]:
    try:
        a = int(r[ca], 16)
    except ValueError:
        continue
    if base is None:
        base = a
    s = float(r[cs] or 0)
    tot += s
    st = sorted(((float(r[i] or 0), n[6:]) for i, n in stalls), reverse=True)[:2]
    out.append((s, a - base, addr2line.get(a - base, ('?', 0)), r[csrc].strip()[:48], r[ci], st))
for s, a, (f, l), src, n, st in sorted(out, reverse=True)[:top]:
    print('%5.1f%% %05x %-16s:%4d x%-9s %-48s %s' % (100 * s / tot, a, f, l, n, src,
          ' '.join('%s=%d' % (n2, v) for v, n2 in st if v)))
