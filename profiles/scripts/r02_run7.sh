set -x
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
(timeout 1800 python -m pytest tests -x -q -m gpu > gpurun_out/r2g_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2g_pytest.log; tail -6 gpurun_out/r2g_pytest.log)
for CFG in "0 5" "0 4" "2 5" "3 5"; do
  set -- $CFG
  ABB_BLOCK_TIERS=$1 ABB_S1_MINB=$2 timeout 300 python bench.py --workload L --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2g_bench_L_t$1_m$2.json 2> gpurun_out/r2g_bench_L_t$1_m$2.err
  python -c "
import json; d=json.load(open('gpurun_out/r2g_bench_L_t$1_m$2.json')); print('RESULT TIERS=$1 MINB=$2', d['ms_per_step'], d['walk_ms_per_step'], d['paths_ms_per_step'], d['e2e']['value']/1e6, d['tier_handoffs']['first'])"
done
ABB_BLOCK_TIERS=0 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 120 -c 80 --csv --log-file gpurun_out/r2g_launches_L.csv python bench.py --workload L --steps 2 --warmup 3 --no-cpu-baseline --check 0 > gpurun_out/r2g_ncu_bench.log 2>&1
grep -E "walk_smem|walk_global|walk_block" gpurun_out/r2g_launches_L.csv | awk -F'","' '{gsub(/"/,"",$NF); print substr($5,1,35), $NF/1e6 " ms"}' | head -8
