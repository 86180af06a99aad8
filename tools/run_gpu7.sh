set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "jit" > gpurun_out/t_fast8.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/t_fast8.log
FLAGS="--steps 5 --warmup 3 --stream-rows 0 --cpu-rows 200000 --cpu-seconds 0.5 --e2e-steps 1 --file-steps 0 --cfg-steps 2 --pool-rows 2000000"
timeout 900 python bench.py $FLAGS > gpurun_out/bench8.json 2> gpurun_out/bench8.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench8.json').read().strip().splitlines()[-1])
print('value',d['value'],'roofline',d['roofline'],'parity',d.get('parity'))
for c in d.get('configs',[]): print(c.get('query'),c.get('value'),c.get('roofline',{}).get('frac'),c.get('parity'))
PY
timeout 900 ncu --set full --clock-control none --import-source on -k regex:dng_scan_kernel_j -s 3 -c 1 -o gpurun_out/prof8 python bench.py $FLAGS > gpurun_out/ncu8.log 2>&1
tail -2 gpurun_out/ncu8.log
