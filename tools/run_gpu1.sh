set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -k "fast" > gpurun_out/t_fast.log 2>&1; echo "rc=$?" >> gpurun_out/t_fast.log
tail -25 gpurun_out/t_fast.log
DNG_KERNEL=fast PROBE_Q=C2 timeout 600 ncu --set full --clock-control none --import-source on -k regex:scan_kernel_f -s 1 -c 1 -o gpurun_out/prof_f2 python tools/probe.py 8000000 > gpurun_out/ncu_f2.log 2>&1
tail -5 gpurun_out/ncu_f2.log
