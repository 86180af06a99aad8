set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x -k "jit or fast" > gpurun_out/t_jit.log 2>&1; echo "rc=$?" >> gpurun_out/t_jit.log
tail -8 gpurun_out/t_jit.log
DNG_KERNEL=fast DNG_JIT=sync PROBE_Q=count,C2,C3,C4,C5,date timeout 300 python tools/probe.py 8000000 > gpurun_out/probe_jit.txt 2>&1
grep "templates=1" gpurun_out/probe_jit.txt
DNG_KERNEL=fast DNG_JIT=sync PROBE_Q=C3 timeout 600 ncu --set full --clock-control none --import-source on -k regex:dng_scan_kernel_j -s 1 -c 1 -o gpurun_out/prof_j3 python tools/probe.py 8000000 > gpurun_out/ncu_j3.log 2>&1
tail -3 gpurun_out/ncu_j3.log
