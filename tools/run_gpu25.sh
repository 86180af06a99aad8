cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
FLAGS="--steps 2 --warmup 3 --stream-rows 0 --cfg-steps 0 --file-steps 3 --e2e-steps 1 --cpu-seconds 1"
for omp in default 8 default; do
if [ $omp = default ]; then timeout 600 python bench.py $FLAGS > gpurun_out/b25.json 2> gpurun_out/b25.err
else OMP_NUM_THREADS=$omp timeout 600 python bench.py $FLAGS > gpurun_out/b25.json 2> gpurun_out/b25.err; fi
python - $omp <<'PY'
import json,sys
d=json.loads(open('gpurun_out/b25.json').read().strip().splitlines()[-1])
print(sys.argv[1], 'file', d.get('e2e_file',{}).get('gbs'), 'e2e', d['e2e']['h2d_gbs'], 'cpu', d['cpu_baseline']['value'])
PY
grep -c . /proc/self/status > /dev/null
done
cat /sys/fs/cgroup/cpu.stat | head -6
