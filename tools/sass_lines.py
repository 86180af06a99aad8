#!/usr/bin/env python
"""Static SASS listing of scan_kernel with source lines (no GPU needed).
  python tools/sass_lines.py [LIB] > listing.txt"""
import glob, os, re, subprocess, sys, tempfile
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(root, 'dragnet_b200', 'libdragnet_gpu.so')
d = tempfile.mkdtemp()
subprocess.run(['cuobjdump', '-xelf', 'all', lib], cwd=d, capture_output=True)
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
        want = os.environ.get('NCU_FN', 'scan_kernel_w')
        if not (fn and ((want + 'ENS_8ScanArgs') in fn or
                        ('fILi' in want and want in fn))):
            continue
        m = re.match(r'^\s*/\*([0-9a-f]{4,})\*/\s+(\S.*?);', l)
        if m:
            print('%s %-16s:%4d  %s' % (m.group(1), fl, ln, m.group(2)))
        elif re.match(r'^\.L_x_\d+:', l):
            print(l)
