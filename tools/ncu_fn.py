#!/usr/bin/env python
"""Per-function and per-source-line executed-instruction summary of an ncu
report whose kernel calls other functions (the run-time linked F kernel).
  python tools/ncu_fn.py REPORT NRECORDS [TOPN]"""
import collections, csv, subprocess, sys
rep, nrec = sys.argv[1], float(sys.argv[2])
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
f = lambda x: float(x) if x not in ('', '-') else 0.0
raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units, r = rows[0], rows[1], rows[2]
want = ['gpu__time_duration.sum', 'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'smsp__thread_inst_executed_per_inst_executed.ratio',
        'launch__registers_per_thread']
for i, h in enumerate(hdr):
    if h in want:
        print(h, units[i], r[i])
    elif 'smsp__average_warps_issue_stalled' in h and h.endswith('_per_issue_active.ratio') and f(r[i]) >= 0.3:
        print(' stall', h.split('stalled_')[1].split('_per_issue')[0], r[i])
src = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
secs = []; cur = None
for r in csv.reader(src.splitlines()):
    if r and r[0] == 'Kernel Name':
        cur = {'name': r[1], 'hdr': None, 'data': []}; secs.append(cur); continue
    if cur is not None and cur['hdr'] is None:
        cur['hdr'] = r; continue
    if cur is not None and len(r) == len(cur['hdr']):
        cur['data'].append(r)
for s in secs:
    h = s['hdr']; ci = h.index('Instructions Executed'); cs = h.index('# Samples'); ct = h.index('Thread Instructions Executed')
    tot = sum(f(r[ci]) for r in s['data']); st = sum(f(r[ct]) for r in s['data']); ss = sum(f(r[cs]) for r in s['data'])
    print('%-28s sass %5d  %6.1f warp-instr/rec  thr/inst %4.1f  samples %d' % (s['name'], len(s['data']), tot / nrec, st / max(1, tot), ss))
cs_ = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'cuda,sass'], capture_output=True, text=True).stdout
agg = collections.Counter(); thr = collections.Counter(); smp = collections.Counter()
fpath = fn = hdr = None
for r in csv.reader(cs_.splitlines()):
    if not r: continue
    if r[0] == 'File Path': fpath = r[1].split('/')[-1]; continue
    if r[0] == 'Function Name': fn = r[1]; continue
    if r[0] == 'Line No':
        hdr = r; ci = hdr.index('Instructions Executed'); ct = hdr.index('Thread Instructions Executed'); cs = hdr.index('# Samples'); continue
    if hdr and len(r) == len(hdr) and r[0]:
        try: ln = int(r[0])
        except ValueError: continue
        k = (fn, fpath, ln); agg[k] += f(r[ci]); thr[k] += f(r[ct]); smp[k] += f(r[cs])
for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:top]:
    print('%-22s %-18s:%4d %5.2f/rec thr/inst %4.1f samples %5.0f' % (k[0][:22], k[1], k[2], v / nrec, thr[k] / max(v, 1), smp[k]))
