"""How the F kernel's instruction fetches depend on the size of its hot loop:
records of K string fields (K = 2, 4, 8 -> matchers of K blocks) of about the
same length as the bench's, scanned from HBM.  Run under
  ncu --metrics gcc__cache_requests_type_instruction.sum,sm__icc_requests.sum,sm__icc_request_hit_rate.pct,smsp__inst_executed.sum,gpu__time_duration.sum -k regex:dng_scan_kernel_j
"""
import os, sys, json, random
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, 'tests'))
import numpy as np, torch
from dragnet_b200 import datasource_gpu
import corpus

os.environ.setdefault('DNG_JIT', 'sync')
n = int(sys.argv[1]) if len(sys.argv) > 1 else 2000000
rnd = random.Random(7)
for K in (2, 4, 8):
    width = max(4, (200 - 8 * K) // K)
    pool = []
    for i in range(20000):
        rec = {}
        for f in range(K):
            ln = rnd.randint(width - 3, width + 3) if f else rnd.randint(2, 5)
            rec['f%d' % f] = ''.join(rnd.choice('abcdefghijklmnopqrstuvwxyz0123456789/') for _ in range(ln)) if f else rnd.choice(['GET', 'PUT', 'HEAD', 'DELETE'])
        pool.append(json.dumps(rec, separators=(',', ':')))
    blob = ('\n'.join(pool) + '\n').encode()
    reps = max(1, n // len(pool))
    host = np.frombuffer(blob * reps, dtype=np.uint8)
    buf = torch.from_numpy(host.copy()).cuda()
    plan = corpus.make_plan(['-b', 'f0'], None)
    for rep in range(3):
        r = datasource_gpu.run_plan(plan, device_buffers=[(buf.data_ptr(), buf.numel())])
    nrec = reps * len(pool)
    ms = r.stats['kernel_ms']
    print('K=%d fields: %d records, %.1f B/line, kernel %.3f ms, %.2f Grec/s, %.0f GB/s, templated %s' % (
        K, nrec, buf.numel() / nrec, ms, nrec / ms / 1e6, buf.numel() / ms / 1e6, r.stats.get('templated_records')), flush=True)
