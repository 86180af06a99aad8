cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python tools/cardinality_probe.py 8000000 2>&1 | tail -12
FLAGS="--steps 5 --warmup 3 --stream-rows 0 --cpu-rows 200000 --cpu-seconds 0.5 --e2e-steps 1 --file-steps 0 --cfg-steps 3 --pool-rows 2000000"
timeout 600 python bench.py $FLAGS > gpurun_out/bench15.json 2> gpurun_out/bench15.err; echo "bench rc=$?"; tail -2 gpurun_out/bench15.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench15.json').read().strip().splitlines()[-1])
print('value %.4g'%d['value'],'frac %.4f'%d['roofline']['frac'],'parity',d.get('parity'), d['roofline'].get('kernel'), 'launches', d.get('gpu_launches'))
for c in d.get('configs',[]):
    if c.get('query'): print('   ',c.get('query'),'%.4g'%c.get('value'),c.get('roofline_frac'))
PY
timeout 900 python -m pytest tests/test_gpu_feeds.py tests/test_fanout.py tests/test_reference_goldens_extra.py -x -q -m gpu 2>&1 | tail -2
