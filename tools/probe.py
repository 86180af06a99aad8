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
for name in os.environ.get('PROBE_Q', 'count,C2,C3,C4,C5,date').split(','):
    argv, ds = corpus.BASELINE_QUERIES[name]
    plan = corpus.make_plan(argv, ds)
    for rep in range(3):
        r = datasource_gpu.run_plan(plan, device_buffers=[(buf.data_ptr(), ln.value)])
    ms = r.stats['kernel_ms']
    print('%-6s kernel %.2f ms  %.1f Mrec/s  %.1f GB/s  points %d  lines %d slow %d' % (
        name, ms, n / ms / 1e3, ln.value / ms / 1e6, len(r.points), r.flat_counters['lines'], r.flat_counters['slowpath_records']), flush=True)
