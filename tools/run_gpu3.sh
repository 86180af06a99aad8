set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/t_all.log 2>&1; echo "rc=$?" >> gpurun_out/t_all.log
tail -15 gpurun_out/t_all.log
timeout 900 python bench.py --rows 20000000 --pool-rows 5000000 --stream-rows 50000000 --steps 3 --warmup 3 > gpurun_out/bench_small.json 2> gpurun_out/bench_small.err; echo "bench rc=$?"
tail -5 gpurun_out/bench_small.err
cat gpurun_out/bench_small.json | cut -c1-3000
