set -x
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
(timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r2f_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2f_pytest.log; tail -6 gpurun_out/r2f_pytest.log)
for CFG in "0 8" "2 8" "3 8" "3 4" "1 4"; do
  set -- $CFG
  ABB_BLOCK_TIERS=$1 ABB_MID_WARPS=$2 timeout 300 python bench.py --workload L --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2f_bench_L_t$1_w$2.json 2> gpurun_out/r2f_bench_L_t$1_w$2.err
  python -c "
import json; d=json.load(open('gpurun_out/r2f_bench_L_t$1_w$2.json')); print('RESULT TIERS=$1 MIDW=$2', d['ms_per_step'], d['walk_ms_per_step'], d['paths_ms_per_step'], d['e2e']['value']/1e6, d['tier_handoffs'])"
done
ABB_BLOCK_TIERS=0 timeout 900 ncu --set full --clock-control none --import-source on -k regex:"walk_smem_kernel|walk_global_kernel" -c 2 -o gpurun_out/r2f_walk_t0 python bench.py --workload L --steps 1 --warmup 3 --no-cpu-baseline --check 0 > gpurun_out/r2f_ncu.log 2>&1
ls -la gpurun_out/*.ncu-rep
