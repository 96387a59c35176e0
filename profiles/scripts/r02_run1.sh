set -x
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
(timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/r2a_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2a_pytest.log; tail -5 gpurun_out/r2a_pytest.log)
timeout 600 python bench.py --workload L --steps 5 --warmup 3 > gpurun_out/r2a_bench_L.json 2> gpurun_out/r2a_bench_L.err; tail -3 gpurun_out/r2a_bench_L.err; cat gpurun_out/r2a_bench_L.json | cut -c1-1500
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 120 -c 80 --csv --log-file gpurun_out/r2a_launches_L.csv python bench.py --workload L --steps 2 --warmup 3 --no-cpu-baseline --check 0 > gpurun_out/r2a_ncu_bench.log 2>&1
tail -50 gpurun_out/r2a_launches_L.csv | cut -c1-200
