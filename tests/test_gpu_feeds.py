"""The feeds the headline numbers rely on, against the oracle (`-m gpu`):

  dng_scan_feed_pinned   caller-pinned host buffers (bench.py's e2e leg): odd
                         piece sizes, cuts in the middle of lines, buffers that
                         stay untouched until dng_scan_sync and are reused after
  dng_merge_nccl         shards on two GPUs, tallies merged with one reduce,
                         against ONE oracle scan of all the data (the property
                         the reference's tests/dn/manta/tst.scan_manta.sh.out
                         pins: shard-then-merge == single scan)
"""

import ctypes
import json
import os
import subprocess
import sys

import pytest

sys.path.insert(0, os.path.dirname(__file__))
import corpus  # noqa: E402
from engines import canon_points, py_engine  # noqa: E402

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _decode(plan, s, res, counters_of):
    from dragnet_b200 import datasource_gpu
    names = [b['name'] for b in plan['breakdowns']]
    pts = [(list(zip(names, cols)), v) for cols, v in res.points()]
    return pts, datasource_gpu.stage_counters(plan, counters_of, len(pts))


@pytest.mark.parametrize('kernel', ['auto', 'fast', 'warp', 'tile'])
@pytest.mark.parametrize('piece', [1 << 20, 65536 + 13, 4099, 777])
def test_feed_pinned_matches_oracle(kernel, piece, tmp_path, monkeypatch):
    import torch
    from dragnet_b200 import native
    monkeypatch.delenv('DNG_KERNEL', raising=False)
    if kernel != 'auto':
        monkeypatch.setenv('DNG_KERNEL', kernel)
    n = 30000
    data = native.gen_host(native.gen_params(total_records=n), 0, n)
    data += b'{"req":{"method":"TAIL"},"res":{"statusCode":1}}'  # no newline
    p = tmp_path / 'in.log'
    p.write_bytes(data)
    plan = corpus.make_plan(['-b', 'req.method,res.statusCode', '-f',
                             '{"ne":["req.method","HEAD"]}'])
    exp_p, exp_c = py_engine(plan, [str(p)])

    host = torch.frombuffer(bytearray(data), dtype=torch.uint8).pin_memory()
    s = native.Scan(native.Plan(json.dumps(plan)), 0)
    # pieces of the caller's pinned buffer, cut anywhere; it must only stay
    # valid until dng_scan_sync
    for off in range(0, len(data), piece):
        s.feed_pinned(host.data_ptr() + off, min(piece, len(data) - off))
    s.sync()
    host.fill_(0x58)            # the buffer is the caller's again
    res = s.finish()
    act_p, act_c = _decode(plan, s, res, s.counters())
    s.close()
    assert canon_points(act_p) == canon_points(exp_p)
    assert act_c == exp_c


def test_feed_pinned_buffer_reuse_between_syncs(tmp_path):
    """One pinned buffer refilled for every piece, as a streaming caller would
    (fill, feed, sync, refill): nothing may still be read after sync."""
    import torch
    from dragnet_b200 import native
    n = 20000
    data = native.gen_host(native.gen_params(total_records=n), 0, n)
    p = tmp_path / 'in.log'
    p.write_bytes(data)
    plan = corpus.make_plan(['-b', 'host,operation'])
    exp_p, exp_c = py_engine(plan, [str(p)])
    piece = 300007
    stage = torch.empty(piece, dtype=torch.uint8).pin_memory()
    s = native.Scan(native.Plan(json.dumps(plan)), 0)
    for off in range(0, len(data), piece):
        chunk = data[off:off + piece]
        stage[:len(chunk)] = torch.frombuffer(bytearray(chunk),
                                              dtype=torch.uint8)
        s.feed_pinned(stage.data_ptr(), len(chunk))
        s.sync()
    res = s.finish()
    act_p, act_c = _decode(plan, s, res, s.counters())
    s.close()
    assert canon_points(act_p) == canon_points(exp_p)
    assert act_c == exp_c


WORKER = r'''
import ctypes, json, os, sys
sys.path.insert(0, %(root)r)
sys.path.insert(0, os.path.join(%(root)r, 'tests'))
import torch
import torch.distributed as dist
from dragnet_b200 import native
rank = int(os.environ['RANK']); world = int(os.environ['WORLD_SIZE'])
torch.cuda.set_device(rank)
dist.init_process_group('nccl', device_id=torch.device('cuda', rank))
L = native.lib()
idbuf = torch.zeros(128, dtype=torch.uint8)
if rank == 0:
    raw = ctypes.create_string_buffer(128)
    assert L.dng_comm_unique_id(raw) == 0
    idbuf = torch.frombuffer(bytearray(raw.raw), dtype=torch.uint8)
idg = idbuf.cuda(rank)
dist.broadcast(idg, 0)
comm = ctypes.c_void_p()
err = ctypes.create_string_buffer(256)
assert L.dng_comm_init(ctypes.byref(comm), world, rank, bytes(idg.cpu().numpy().tobytes()), rank, err, 256) == 0, err.value
plan = json.load(open(sys.argv[1]))
s = native.Scan(native.Plan(json.dumps(plan)), rank)
s.feed_file(sys.argv[2 + rank])
out = ctypes.c_void_p()
ctr = native.DngCounters()
rc = L.dng_merge_nccl(s.handle, comm, 0, ctypes.byref(out), ctypes.byref(ctr))
assert rc == 0, rc
if rank == 0:
    pts = native.Result(out).points()
    json.dump({'points': [[[c.hex() if isinstance(c, bytes) else c for c in cols], v] for cols, v in pts],
               'counters': ctr.as_dict()}, open(sys.argv[2 + world], 'w'))
s.close()
L.dng_comm_destroy(comm)
dist.destroy_process_group()
'''


def test_merge_nccl_two_gpus_matches_single_oracle_scan(tmp_path):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip('needs two GPUs')
    from dragnet_b200 import native
    n = 40000
    data = native.gen_host(native.gen_params(total_records=n), 0, n)
    cut = data.index(b'\n', len(data) // 3) + 1     # uneven shards
    shards = [tmp_path / 'a.log', tmp_path / 'b.log']
    shards[0].write_bytes(data[:cut])
    shards[1].write_bytes(data[cut:] + b'not json\n')
    plan = corpus.make_plan(['-b', 'operation,req.method,host'])
    exp_p, exp_c = py_engine(plan, [str(shards[0]), str(shards[1])])
    pf = tmp_path / 'plan.json'
    pf.write_text(json.dumps(plan))
    wf = tmp_path / 'worker.py'
    wf.write_text(WORKER % {'root': ROOT})
    outf = tmp_path / 'out.json'
    subprocess.run([sys.executable, '-m', 'torch.distributed.run',
                    '--nnodes=1', '--nproc-per-node', '2', '--master-addr',
                    '127.0.0.1', '--master-port', '29517', str(wf), str(pf),
                    str(shards[0]), str(shards[1]), str(outf)],
                   check=True, timeout=600)
    got = json.load(open(outf))
    names = [b['name'] for b in plan['breakdowns']]
    pts = [(list(zip(names, [bytes.fromhex(c) if isinstance(c, str) else c
                             for c in cols])), v)
           for cols, v in got['points']]
    assert canon_points(pts) == canon_points(exp_p)
    assert got['counters']['lines'] == n + 1
    assert got['counters']['invalid_json'] == 1
