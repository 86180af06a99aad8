cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -x -q -m gpu > gpurun_out/t_final.log 2>&1; echo "pytest all rc=$?"; tail -2 gpurun_out/t_final.log
python -c "import __graft_entry__ as g; g.smoke()"; echo "smoke rc=$?"
