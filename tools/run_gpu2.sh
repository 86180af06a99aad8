set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
DNG_KERNEL=fast DNG_JIT=sync PROBE_Q=count,C2,C3,C5 timeout 300 python tools/probe.py 8000000 > gpurun_out/probe_jit.txt 2>&1
cat gpurun_out/probe_jit.txt
DNG_KERNEL=fast DNG_JIT=sync PROBE_Q=C2 timeout 600 ncu --set full --clock-control none --import-source on -k regex:dng_scan_kernel_j -s 1 -c 1 -o gpurun_out/prof_j2 python tools/probe.py 8000000 > gpurun_out/ncu_j2.log 2>&1
tail -3 gpurun_out/ncu_j2.log
