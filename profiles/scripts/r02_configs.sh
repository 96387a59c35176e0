# BASELINE configs 2, 3 (M1, M2 on one GPU) and config 5 (L, enumerate mode) with the final build; sanitizer passes on the smoke estate
set -x
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
for W in M1 M2; do
  timeout 300 python bench.py --workload $W --steps 10 --warmup 3 > gpurun_out/r02_bench_$W.json 2> gpurun_out/r02_bench_$W.err
  python -c "
import json; d=json.load(open('gpurun_out/r02_bench_$W.json')); print('RESULT $W', d['value']/1e6, d['ms_per_step'], d['walk_ms_per_step'], d['e2e']['value']/1e6, d['roofline']['frac'], d['cpu_baseline']['value'])"
done
timeout 600 python bench.py --workload L --mode enumerate --steps 3 --warmup 3 > gpurun_out/r02_bench_L_enumerate.json 2> gpurun_out/r02_bench_L_enumerate.err
tail -3 gpurun_out/r02_bench_L_enumerate.err | cut -c1-300
python -c "
import json; d=json.load(open('gpurun_out/r02_bench_L_enumerate.json')); print('RESULT L-enumerate', d['value']/1e6, d['ms_per_step'], d['enumerate'], d['e2e']['value']/1e6, d['cpu_baseline'])"
(timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_sanitizer_memcheck_smoke.log 2>&1; echo "memcheck rc=$?" >> gpurun_out/r02_sanitizer_memcheck_smoke.log; tail -4 gpurun_out/r02_sanitizer_memcheck_smoke.log)
(timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_sanitizer_racecheck_smoke.log 2>&1; echo "racecheck rc=$?" >> gpurun_out/r02_sanitizer_racecheck_smoke.log; tail -4 gpurun_out/r02_sanitizer_racecheck_smoke.log)
(timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_tiers.py -x -q -m gpu -k "impact and shrunk and 2000" > gpurun_out/r02_sanitizer_racecheck_tiers.log 2>&1; echo "racecheck rc=$?" >> gpurun_out/r02_sanitizer_racecheck_tiers.log; tail -4 gpurun_out/r02_sanitizer_racecheck_tiers.log)
(timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_tiers.py -x -q -m gpu -k "impact and shrunk and 2000" > gpurun_out/r02_sanitizer_memcheck_tiers.log 2>&1; echo "memcheck rc=$?" >> gpurun_out/r02_sanitizer_memcheck_tiers.log; tail -4 gpurun_out/r02_sanitizer_memcheck_tiers.log)
