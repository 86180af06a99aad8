cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/t_all16.log 2>&1; echo "pytest all rc=$?"; tail -2 gpurun_out/t_all16.log
python -c "import __graft_entry__ as g; g.smoke()"; echo "smoke rc=$?"
timeout 1200 python bench.py > gpurun_out/bench_r2_n1.json 2> gpurun_out/bench_r2_n1.err; echo "bench rc=$?"
tail -3 gpurun_out/bench_r2_n1.err
timeout 600 python bench.py --impl reference > gpurun_out/bench_r2_ref.json 2> gpurun_out/bench_r2_ref.err; echo "ref rc=$?"
FLAGS="--steps 2 --warmup 3 --stream-rows 0 --cpu-rows 200000 --cpu-seconds 0.5 --e2e-steps 1 --file-steps 0 --cfg-steps 1 --pool-rows 2000000"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2_launches.csv python bench.py $FLAGS > gpurun_out/bench_under_ncu.json 2> gpurun_out/bench_under_ncu.err
timeout 900 ncu --set full --clock-control none --import-source on -k regex:dng_scan_kernel_j -s 4 -c 1 -o gpurun_out/prof_r2_c3_100M python bench.py $FLAGS > gpurun_out/ncu_r2.log 2>&1
tail -3 gpurun_out/ncu_r2.log
cut -c1-1200 gpurun_out/bench_r2_n1.json
