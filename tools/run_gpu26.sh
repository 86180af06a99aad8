cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
DNG_BENCH_TRACE=1 timeout 600 python bench.py --steps 2 --stream-rows 0 --cfg-steps 1 > gpurun_out/b26.json 2> gpurun_out/b26.err &
BP=$!
sleep 5
while kill -0 $BP 2>/dev/null; do
  echo "T $(date +%s.%N)" >> gpurun_out/top26.log
  ps -eLo pid,tid,pcpu,comm --sort=-pcpu | head -12 >> gpurun_out/top26.log
  cat /sys/fs/cgroup/cpu.stat | grep -E "nr_throttled|throttled_usec" >> gpurun_out/top26.log
  sleep 0.5
done
grep TRACE gpurun_out/b26.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/b26.json').read().strip().splitlines()[-1])
print('file', d.get('e2e_file',{}).get('gbs'))
PY
