set -x
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
(timeout 1500 python -m pytest tests/test_gpu_tiers.py tests/test_gpu_parity.py -x -q -m gpu > gpurun_out/r2h_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2h_pytest.log; tail -4 gpurun_out/r2h_pytest.log)
timeout 400 python bench.py --workload L --steps 5 --warmup 3 > gpurun_out/r2h_bench_L.json 2> gpurun_out/r2h_bench_L.err
tail -4 gpurun_out/r2h_bench_L.err | cut -c1-300
python -c "
import json; d=json.load(open('gpurun_out/r2h_bench_L.json')); print('RESULT default', d['value']/1e6, d['ms_per_step'], d['walk_ms_per_step'], d['paths_ms_per_step'], d['e2e']['value']/1e6, d['roofline']['frac'], d['cpu_baseline']['value'], d['tier_handoffs']['first'])"
bash profiles/scripts/r02_configs.sh
