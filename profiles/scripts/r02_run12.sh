set -x
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
(ABB_LIB=libabb200_e2e.so timeout 900 python -m pytest tests/test_gpu_api.py tests/test_gpu_scale_properties.py -x -q -m gpu -k "not replay" > gpurun_out/r2l_pytest_e2e.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2l_pytest_e2e.log; tail -3 gpurun_out/r2l_pytest_e2e.log)
(ABB_LIB=libabb200_e2e.so timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1)
run() { name=$1; shift
  env "$@" timeout 300 python bench.py --workload L --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2l_bench_$name.json 2> gpurun_out/r2l_bench_$name.err
  grep "exposure_host" gpurun_out/r2l_bench_$name.err | tail -2
  python -c "
import json; d=json.load(open('gpurun_out/r2l_bench_$name.json')); print('RESULT $name', d['ms_per_step'], d['walk_ms_per_step'], 'e2e', d['e2e']['value']/1e6, d['e2e']['ms_per_step'], d['e2e']['python_zero_copy'])"
}
run default ABB_TRACE=1
run e2e ABB_LIB=libabb200_e2e.so ABB_TRACE=1
