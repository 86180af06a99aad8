#!/usr/bin/env python
"""Round-2 evidence: raw outputs of tools/gpu_evidence.sh (gpurun_out/) -> the
tracked summaries under profiles/r2_*.  The run, on the GPU box:

  python bench.py                         > gpurun_out/bench_r2_n1.json
  python bench.py --impl reference        > gpurun_out/bench_r2_ref.json
  ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv \
      --log-file gpurun_out/r2_launches.csv python bench.py $FLAGS
  ncu --set full --clock-control none --import-source on \
      -k regex:dng_scan_kernel_j -s 3 -c 1 -o gpurun_out/prof_r2_c3_100M python bench.py $FLAGS
"""
import collections
import csv
import json
import os
import shutil
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = os.path.join(ROOT, 'gpurun_out')
P = os.path.join(ROOT, 'profiles')
NREC = 100000000


def last_json(path):
    return json.loads(open(path).read().strip().splitlines()[-1])


b = last_json(os.path.join(G, 'bench_r2_n1.json'))
shutil.copy(os.path.join(G, 'bench_r2_n1.json'), os.path.join(P, 'r2_bench_n1_C3.json'))
shutil.copy(os.path.join(G, 'bench_r2_ref.json'), os.path.join(P, 'r2_bench_reference_arm.json'))
shutil.copy(os.path.join(G, 'r2_launches.csv'), os.path.join(P, 'r2_launches_bench_C3.csv'))

rows = list(csv.reader(l for l in open(os.path.join(G, 'r2_launches.csv')) if l.startswith('"')))
hdr = rows[0]
ki, vi, mi = hdr.index('Kernel Name'), hdr.index('Metric Value'), hdr.index('Metric Name')
unit = rows[1][hdr.index('Metric Unit')]
agg = collections.defaultdict(lambda: [0, 0.0])
for r in rows[1:]:
    if r[mi] == 'gpu__time_duration.sum':
        agg[r[ki]][0] += 1
        agg[r[ki]][1] += float(r[vi].replace(',', ''))
tot = sum(v[1] for v in agg.values())
with open(os.path.join(P, 'r2_launch_summary.csv'), 'w') as f:
    f.write('kernel,launches,total_%s,share\n' % unit)
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        f.write('"%s",%d,%.1f,%.4f\n' % (k, v[0], v[1], v[1] / tot))

rep = os.path.join(G, 'prof_r2_c3_100M.ncu-rep')
raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr = rows[0]
d, u = dict(zip(hdr, rows[2])), dict(zip(hdr, rows[1]))
want = ['gpu__time_duration.sum', 'smsp__inst_executed.sum',
        'smsp__thread_inst_executed_per_inst_executed.ratio',
        'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'launch__registers_per_thread',
        'launch__block_size', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'sm__icc_request_hit_rate.pct',
        'gcc__cache_requests_type_instruction.sum.pct_of_peak_sustained_elapsed']
metrics = {k: {'value': float(d[k].replace(',', '')), 'unit': u[k]} for k in want if k in d}
stalls = {}
for k in hdr:
    if 'smsp__average_warps_issue_stalled' in k and k.endswith('_per_issue_active.ratio'):
        stalls[k.split('stalled_')[1].split('_per_issue')[0]] = float(d[k])


def to_bytes(m):
    v, un = m['value'], m['unit'].lower()
    return v * {'byte': 1, 'kbyte': 1e3, 'mbyte': 1e6, 'gbyte': 1e9, 'tbyte': 1e12}[un]


alg = float(b['roofline']['bytes_per_launch'])
rd, wr = to_bytes(metrics['dram__bytes_read.sum']), to_bytes(metrics['dram__bytes_write.sum'])
summ = json.load(open(os.path.join(P, 'r2_ncu_summary.json')))
summ['r2_final_C3_100M_rows'] = {
    'command': 'ncu --set full --clock-control none --import-source on -k regex:dng_scan_kernel_j -s 3 -c 1 python bench.py ... (tools/gpu_evidence.sh)',
    'kernel': 'dng_scan_kernel_j (scan_kernel_f, 28 warps of 72 registers, linked at run time with the matcher generated for the 3 learned templates; plan constant-folded)',
    'query': 'C3: ' + b['config']['query'],
    'records': NREC, 'algorithmic_bytes': alg, 'metrics': metrics,
    'stalls_per_issue': stalls,
    'warp_instructions_per_record': metrics['smsp__inst_executed.sum']['value'] / NREC,
    'dram_traffic_over_algorithmic': (rd + wr) / alg,
}
json.dump(summ, open(os.path.join(P, 'r2_ncu_summary.json'), 'w'), indent=1)
import sys
sys.path.insert(0, ROOT)
import bench
json.dump({'kernel': 'dng_scan_kernel_j', 'query': 'C3', 'rows': NREC, 'algorithmic_bytes': alg,
           'sources_sha16': bench.kernel_sources_sha16(),
           'dram_bytes_read': rd, 'dram_bytes_write': wr,
           'traffic_over_algorithmic': (rd + wr) / alg,
           'source': 'ncu --set full capture of the bench launch (profiles/r2_ncu_summary.json r2_final_C3_100M_rows)'},
          open(os.path.join(P, 'r2_traffic.json'), 'w'), indent=1)
print('value', b['value'], 'frac', b['roofline']['frac'], 'instr/rec',
      summ['r2_final_C3_100M_rows']['warp_instructions_per_record'],
      'traffic x', (rd + wr) / alg)
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:8]:
    print('%-60s %4d %12.1f %.4f' % (k[:60], v[0], v[1], v[1] / tot))
