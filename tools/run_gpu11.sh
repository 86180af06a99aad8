cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
FLAGS="--steps 5 --warmup 3 --stream-rows 0 --cpu-rows 200000 --cpu-seconds 0.5 --e2e-steps 1 --file-steps 0 --cfg-steps 3 --pool-rows 2000000"
for v in 896 base; do
  if [ $v = base ]; then DNG_KERNEL=fast timeout 600 python bench.py $FLAGS > gpurun_out/bench11_$v.json 2> gpurun_out/bench11_$v.err
  else DNG_KERNEL=fast timeout 600 python tools/exp/run_variant.py tools/exp/lib$v.so $FLAGS > gpurun_out/bench11_$v.json 2> gpurun_out/bench11_$v.err; fi
  echo "$v rc=$?"; tail -2 gpurun_out/bench11_$v.err
  python - $v <<'PY'
import json,sys
d=json.loads(open('gpurun_out/bench11_%s.json'%sys.argv[1]).read().strip().splitlines()[-1])
print(sys.argv[1],'value %.4g'%d['value'],'frac %.4f'%d['roofline']['frac'],'parity',d.get('parity'), d['roofline'].get('kernel'))
for c in d.get('configs',[]):
    if c.get('query'): print('   ',c.get('query'),'%.4g'%c.get('value'),c.get('roofline_frac'))
PY
done
