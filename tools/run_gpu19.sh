cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 3 --warmup 3 --stream-rows 0 --cfg-steps 1 --cpu-seconds 0.5 --cpu-rows 200000 --file-steps 0 > gpurun_out/bench_n8.json 2> gpurun_out/bench_n8.err; echo "bench n8 rc=$?"
tail -2 gpurun_out/bench_n8.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_n8.json').read().strip().splitlines()[-1])
print({k:d.get(k) for k in ('value','n_gpus','ms_per_step','parity','merge_parity','scaling','finish_merge_ms_per_step')}, d['roofline']['frac'], d['e2e'].get('value'), d['e2e'].get('parity'))
for c in d.get('configs',[]): print('  ',c.get('query'), c.get('value'), c.get('roofline_frac'))
PY
