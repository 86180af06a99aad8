cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python tools/icache_probe.py 2000000 2>&1 | tail -5
timeout 900 ncu --metrics gcc__cache_requests_type_instruction.sum,sm__icc_requests.sum,sm__icc_request_hit_rate.pct,smsp__inst_executed.sum,gpu__time_duration.sum,smsp__issue_active.avg.pct_of_peak_sustained_active -k regex:dng_scan_kernel_j --csv --log-file gpurun_out/icache_probe.csv python tools/icache_probe.py 2000000 > gpurun_out/icache_probe.log 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/icache_probe.csv')) if len(r)>10]
h=rows[0]
im=h.index('Metric Name'); iv=h.index('Metric Value'); iid=h.index('ID')
d={}
for r in rows[1:]:
    d.setdefault(r[iid],{})[r[im]]=r[iv]
for k,v in d.items(): print(k, {a.split('.')[0][-28:]:b for a,b in v.items()})
PY
