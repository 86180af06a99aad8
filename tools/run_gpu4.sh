set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -k "jit and (synthetic or goldens or feed_pinned)" > gpurun_out/t_jit2.log 2>&1; echo "rc=$?" >> gpurun_out/t_jit2.log
tail -4 gpurun_out/t_jit2.log
DNG_KERNEL=fast DNG_JIT=sync PROBE_Q=count,C2,C3,C4,C5,date timeout 300 python tools/probe.py 8000000 > gpurun_out/probe_jit.txt 2>&1
grep "templates=1" gpurun_out/probe_jit.txt
