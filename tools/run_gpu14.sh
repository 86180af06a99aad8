cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 cuda-gdb -batch -ex "set pagination off" -ex run -ex bt -ex "info threads" --args python -m pytest tests/test_gpu_feeds.py -x -q -m gpu -k "reuse or (auto and 777)" > gpurun_out/gdb14.log 2>&1; echo rc=$?
grep -v "^\[New Thread\|^\[Thread\|^\[Detaching\|warning:" gpurun_out/gdb14.log | tail -45
