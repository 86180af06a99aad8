set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
cat /sys/fs/cgroup/cpu.max 2>/dev/null; nproc; python -c "import os; print(len(os.sched_getaffinity(0)), os.cpu_count())"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 --rows 20000000 --pool-rows 5000000 --stream-rows 20000000 > gpurun_out/bench_r2_n2_small.json 2> gpurun_out/bench_r2_n2_small.err; echo "n2 rc=$?"
tail -5 gpurun_out/bench_r2_n2_small.err
cut -c1-600 gpurun_out/bench_r2_n2_small.json
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 1 --warmup 1 > gpurun_out/bench_r2_n2_ref.json 2> gpurun_out/bench_r2_n2_ref.err; echo "ref n2 rc=$?"
timeout 600 python -m pytest tests/test_gpu_feeds.py -m gpu -q -x -k merge_nccl > gpurun_out/t_merge.log 2>&1; tail -3 gpurun_out/t_merge.log
