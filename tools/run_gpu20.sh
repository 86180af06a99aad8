cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "synthetic or baseline or kernel_choice or merge" 2>&1 | tail -2
timeout 600 python tools/cardinality_probe.py 8000000 2>&1 | tail -9
PROBE_Q=url,C5,C3 timeout 300 python tools/probe.py 8000000 2>&1 | grep "templates=1"
FLAGS="--steps 5 --warmup 3 --stream-rows 0 --cpu-rows 200000 --cpu-seconds 0.5 --e2e-steps 1 --file-steps 0 --cfg-steps 3 --pool-rows 2000000"
timeout 600 python bench.py $FLAGS > gpurun_out/bench20.json 2> gpurun_out/bench20.err; echo "bench rc=$?"; tail -2 gpurun_out/bench20.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench20.json').read().strip().splitlines()[-1])
print('value %.4g'%d['value'],'frac %.4f'%d['roofline']['frac'],'parity',d.get('parity'))
for c in d.get('configs',[]):
    if c.get('query'): print('   ',c.get('query'),'%.4g'%c.get('value'),c.get('roofline_frac'))
PY
