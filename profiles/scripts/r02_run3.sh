set -x
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_scale_properties.py -x -q -m gpu > gpurun_out/r2c_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2c_pytest.log; tail -4 gpurun_out/r2c_pytest.log)
for CFG in "0 5" "0 4" "3 5"; do
  set -- $CFG
  ABB_BLOCK_TIERS=$1 ABB_S1_MINB=$2 timeout 300 python bench.py --workload L --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2c_bench_L_t$1_m$2.json 2> gpurun_out/r2c_bench_L_t$1_m$2.err
  python -c "
import json; d=json.load(open('gpurun_out/r2c_bench_L_t$1_m$2.json')); print('TIERS=$1 MINB=$2', d['value']/1e6, d['ms_per_step'], d['walk_ms_per_step'], d['paths_ms_per_step'], d['e2e']['value']/1e6)"
done
ABB_BLOCK_TIERS=0 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 120 -c 80 --csv --log-file gpurun_out/r2c_launches_L.csv python bench.py --workload L --steps 2 --warmup 3 --no-cpu-baseline --check 0 > gpurun_out/r2c_ncu_bench.log 2>&1
