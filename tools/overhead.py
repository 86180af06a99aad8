"""Where one scan's host-side time goes (open / feed / sync / finish / close)."""
import ctypes, json, os, sys, time
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, 'tests'))
import torch
from dragnet_b200 import native
import corpus

n = int(sys.argv[1]) if len(sys.argv) > 1 else 8000000
params = native.gen_params(total_records=n)
cap = n * 232
buf = torch.empty(cap, dtype=torch.uint8, device='cuda:0')
ln = ctypes.c_size_t()
native.lib().dng_gen_device(ctypes.byref(params), 0, 0, n, buf.data_ptr(), cap, ctypes.byref(ln))
torch.cuda.synchronize()
plan_json = json.dumps(corpus.make_plan(['-b', 'req.method']), separators=(',', ':'))
for rep in range(4):
    t = [time.perf_counter()]
    p = native.Plan(plan_json); t.append(time.perf_counter())
    s = native.Scan(p, 0); t.append(time.perf_counter())
    s.feed_device(buf.data_ptr(), ln.value); t.append(time.perf_counter())
    s.sync(); t.append(time.perf_counter())
    r = s.finish(); t.append(time.perf_counter())
    pts = list(r.points()); t.append(time.perf_counter())
    st = s.kernel_stats(); r.close(); s.close(); p.close(); t.append(time.perf_counter())
    names = ['plan', 'open', 'feed_device', 'sync', 'finish', 'points', 'close']
    print(' '.join('%s %.2f' % (k, (b - a) * 1e3) for k, a, b in zip(names, t, t[1:])),
          '| total %.2f ms kernel %.2f ms' % ((t[-1] - t[0]) * 1e3, st['kernel_ms']), flush=True)
