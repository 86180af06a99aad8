"""The N>1 path on CPU: shard -> per-rank tallies -> dictionary all-gather ->
dense vector -> ONE sum-reduce, over gloo with world_size 2 (the product runs
the same steps over NCCL in dng_merge_nccl).  Per-rank tallies come from the
C++ oracle on each shard (no GPU here); the merged result must equal a single
scan of the whole input -- the property the reference pins with
tests/dn/manta/tst.scan_manta.sh.out == tst.scan_fileset.sh.out."""

import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

sys.path.insert(0, os.path.dirname(__file__))


def _worker(rank, world, port, shard_paths, argv, q):
    import json
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    import corpus
    from engines import cpp_engine
    from dragnet_b200 import native
    plan = corpus.make_plan(argv)
    pts, _ = cpp_engine(plan, [shard_paths[rank]])
    p = native.Plan(json.dumps(plan))
    local = native.result_from_points(
        p, [([v for _, v in f], val) for f, val in pts])
    dicts = [None] * world
    dist.all_gather_object(dicts, local.dict_bytes())
    gdict = native.dict_union(dicts)
    vec = torch.tensor(local.dense(gdict), dtype=torch.int64)
    dist.reduce(vec, 0, op=dist.ReduceOp.SUM)
    if rank == 0:
        merged = local.from_dense(gdict, [int(x) for x in vec])
        q.put(merged.points())
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize('argv', [
    ['-b', 'operation,req.method,host'],
    ['-b', 'latency[aggr=quantize],req.method'],
    ['-b', 'dataLatency[aggr=lquantize,step=100]'],
    [],
])
def test_two_rank_merge_equals_single_scan(argv, tmp_path):
    import corpus
    from engines import canon_points, cpp_engine
    from dragnet_b200 import native
    n = 6000
    shards = []
    for r in range(2):
        data = native.gen_host(native.gen_params(seed=0xD5A60000 + r,
                                                 total_records=n), 0, n)
        if r == 1:
            data = data[:len(data) // 2]       # uneven shards
            data = data[:data.rfind(b'\n') + 1]
        path = tmp_path / ('shard%d.log' % r)
        path.write_bytes(data)
        shards.append(str(path))
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, shards, argv, q))
             for r in range(2)]
    for p in procs:
        p.start()
    merged = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    plan = corpus.make_plan(argv)
    exp, _ = cpp_engine(plan, shards)
    names = [b['name'] for b in plan['breakdowns']]
    got = [(list(zip(names, cols)), v) for cols, v in merged]
    assert canon_points(got) == canon_points(exp)
