cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/t_all13.log 2>&1; echo "pytest all rc=$?"; tail -2 gpurun_out/t_all13.log
python -c "import __graft_entry__ as g; g.smoke()"; echo "smoke rc=$?"
timeout 600 python -m pytest tests/test_gpu_feeds.py -x -q -m gpu > gpurun_out/t13_1.log 2>&1; echo "feeds rc=$?"; tail -1 gpurun_out/t13_1.log
