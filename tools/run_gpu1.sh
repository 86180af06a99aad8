set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -k "fast or jit" > gpurun_out/t_fast.log 2>&1; echo "rc=$?" >> gpurun_out/t_fast.log
tail -25 gpurun_out/t_fast.log
DNG_KERNEL=fast DNG_JIT=sync timeout 300 python tools/probe.py 8000000 > gpurun_out/probe_jit.txt 2>&1
cat gpurun_out/probe_jit.txt
DNG_KERNEL=fast DNG_JIT=sync PROBE_Q=C2 timeout 600 ncu --set full --clock-control none --import-source on -k regex:dng_scan_kernel_j -s 1 -c 1 -o gpurun_out/prof_j1 python tools/probe.py 8000000 > gpurun_out/ncu_j1.log 2>&1
tail -5 gpurun_out/ncu_j1.log
