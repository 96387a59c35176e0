set -x
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
(timeout 1200 python -m pytest tests -x -q -m gpu > gpurun_out/r2e_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2e_pytest.log; tail -6 gpurun_out/r2e_pytest.log)
for T in 0 3; do
  ABB_BLOCK_TIERS=$T timeout 400 python bench.py --workload L --steps 5 --warmup 3 > gpurun_out/r2e_bench_L_t$T.json 2> gpurun_out/r2e_bench_L_t$T.err
  tail -4 gpurun_out/r2e_bench_L_t$T.err | cut -c1-300
  python -c "
import json; d=json.load(open('gpurun_out/r2e_bench_L_t$T.json')); print('RESULT TIERS=$T', d['ms_per_step'], d['walk_ms_per_step'], d['paths_ms_per_step'], d['e2e']['value']/1e6, d['e2e']['first_call_ms'], d['e2e']['python_zero_copy'], d['roofline']['frac'], d['cpu_baseline']['value'])"
done
ABB_BLOCK_TIERS=0 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 120 -c 80 --csv --log-file gpurun_out/r2e_launches_L.csv python bench.py --workload L --steps 2 --warmup 3 --no-cpu-baseline --check 0 > gpurun_out/r2e_ncu_bench.log 2>&1
grep -E "walk_smem|walk_global" gpurun_out/r2e_launches_L.csv | awk -F'","' '{gsub(/"/,"",$NF); print substr($5,1,35), $NF/1e6 " ms"}' | head -8
