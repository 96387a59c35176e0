set -x
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
(ABB_LIB=libabb200_xrow.so timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "impact or distances" > gpurun_out/r2i_pytest_xrow.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2i_pytest_xrow.log; tail -3 gpurun_out/r2i_pytest_xrow.log)
run() { # name, env...
  name=$1; shift
  env "$@" timeout 300 python bench.py --workload L --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2i_bench_$name.json 2> gpurun_out/r2i_bench_$name.err
  python -c "
import json; d=json.load(open('gpurun_out/r2i_bench_$name.json')); print('RESULT $name', d['ms_per_step'], d['walk_ms_per_step'], d['paths_ms_per_step'], d['e2e']['value']/1e6)"
}
run default A=1
run xrow ABB_LIB=libabb200_xrow.so
run g1w48 ABB_G1_WARPS_PER_SM=48
run xrow_g1w48 ABB_LIB=libabb200_xrow.so ABB_G1_WARPS_PER_SM=48
run big_always ABB_BIG_LIMIT=1000000
(timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_tiers.py -x -q -m gpu -k "test_impact_every_tier and shrunk and key0" > gpurun_out/r02_sanitizer_racecheck_tiers.log 2>&1; echo "racecheck rc=$?" >> gpurun_out/r02_sanitizer_racecheck_tiers.log; tail -4 gpurun_out/r02_sanitizer_racecheck_tiers.log)
(timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_tiers.py -x -q -m gpu -k "test_impact_every_tier and shrunk and key0" > gpurun_out/r02_sanitizer_memcheck_tiers.log 2>&1; echo "memcheck rc=$?" >> gpurun_out/r02_sanitizer_memcheck_tiers.log; tail -4 gpurun_out/r02_sanitizer_memcheck_tiers.log)
