set -x
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_gpu_tiers.py -x -q -m gpu > gpurun_out/r2b_pytest_tiers.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2b_pytest_tiers.log; tail -15 gpurun_out/r2b_pytest_tiers.log)
for T in 3 0 1 2; do
  ABB_BLOCK_TIERS=$T timeout 300 python bench.py --workload L --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2b_bench_L_t$T.json 2> gpurun_out/r2b_bench_L_t$T.err
  tail -2 gpurun_out/r2b_bench_L_t$T.err | cut -c1-400; python -c "
import json; d=json.load(open('gpurun_out/r2b_bench_L_t$T.json')); print('TIERS=$T', d['value']/1e6, d['ms_per_step'], d['e2e']['value']/1e6)"
done
