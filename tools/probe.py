"""Quick device-side throughput probe (not the bench): scan N synthetic records
resident in HBM with a few query shapes and print records/s + GB/s."""
import ctypes, json, sys, time
import os; R = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, 'tests'))
import torch
from dragnet_b200 import native, datasource_gpu
import corpus

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4000000
params = native.gen_params(total_records=n)
cap = n * 232
buf = torch.empty(cap, dtype=torch.uint8, device='cuda:0')
ln = ctypes.c_size_t()
t0 = time.time()
rc = native.lib().dng_gen_device(ctypes.byref(params), 0, 0, n, buf.data_ptr(), cap, ctypes.byref(ln))
torch.cuda.synchronize()
print('gen rc', rc, 'bytes', ln.value, 'sec', time.time() - t0, flush=True)
import os
if os.environ.get('PROBE_FANOUT'):
    from dragnet_b200 import query as mod_query
    mets = [{'filter': None, 'breakdowns': []},
            {'filter': {'eq': ['req.method', 'GET']}, 'breakdowns': [
                {'name': 'operation', 'field': 'operation'},
                {'name': 'res.statusCode', 'field': 'res.statusCode'}]},
            {'filter': None, 'breakdowns': [
                {'name': 'latency', 'field': 'latency', 'aggr': 'quantize'},
                {'name': 'host', 'field': 'host'}]},
            {'filter': None, 'breakdowns': [
                {'name': 'req.caller', 'field': 'req.caller'}]}]
    if os.environ.get('PROBE_MATRIX'):
        for interval in ('all', 'hour'):
            for sub in ([0], [1], [2], [3], [0, 1], [0, 1, 2], [0, 1, 2, 3], [2, 2, 2, 2], [0, 0, 0, 0]):
                qs = [mod_query.metricQuery(mets[i], None, None, interval, 'time') for i in sub]
                multi = mod_query.scan_plan_multi(qs, time_field='time')
                for rep in range(3):
                    r = datasource_gpu.run_plan(multi, device_buffers=[(buf.data_ptr(), ln.value)])
                print('interval=%-4s metrics=%-14s kernel %.2f ms points %d' % (interval, sub, r.stats['kernel_ms'], len(r.points)), flush=True)
        sys.exit(0)
    qs = [mod_query.metricQuery(m, None, None, os.environ.get('PROBE_INTERVAL', 'hour'), 'time') for m in mets]
    multi = mod_query.scan_plan_multi(qs, time_field='time')
    for rep in range(int(os.environ.get('PROBE_REPS', 3))):
        r = datasource_gpu.run_plan(multi, device_buffers=[(buf.data_ptr(), ln.value)])
    ms = r.stats['kernel_ms']
    if os.environ.get('PROBE_FANOUT') == 'only':
        sys.exit(0)
    print('fanout4 (one pass, 4 metrics, hourly __dn_ts): kernel %.2f ms  %.1f Mrec/s  points %d' % (ms, n / ms / 1e3, len(r.points)), flush=True)
    tot = 0
    for q in qs:
        plan = mod_query.scan_plan(q, time_field='time')
        for rep in range(3):
            r = datasource_gpu.run_plan(plan, device_buffers=[(buf.data_ptr(), ln.value)])
        tot += r.stats['kernel_ms']
    print('same 4 metrics as 4 separate scans: kernel %.2f ms total' % tot, flush=True)
for name in os.environ.get('PROBE_Q', 'count,C2,C3,C4,C5,date').split(','):
    argv, ds = corpus.BASELINE_QUERIES[name]
    plan = corpus.make_plan(argv, ds)
    for tm in (True, False):
        for rep in range(3):
            t1 = time.time()
            r = datasource_gpu.run_plan(plan, device_buffers=[(buf.data_ptr(), ln.value)], templates=tm)
            wall = (time.time() - t1) * 1e3
        ms = r.stats['kernel_ms']
        print('%-6s templates=%d kernel %.2f ms  %.1f Mrec/s  %.1f GB/s  wall %.2f ms  points %d  lines %d slow %d  tmpl %d/%d' % (
            name, tm, ms, n / ms / 1e3, ln.value / ms / 1e6, wall, len(r.points), r.flat_counters['lines'], r.flat_counters['slowpath_records'],
            r.stats['templates'], r.stats['templated_records']), flush=True)
