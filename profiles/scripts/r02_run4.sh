set -x
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
(ABB_LIB=libabb200_winpf.so timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "impact or distances or traverse or bfs" > gpurun_out/r2d_pytest_winpf.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2d_pytest_winpf.log; tail -3 gpurun_out/r2d_pytest_winpf.log)
for CFG in "libabb200.so 5" "libabb200_g4.so 5" "libabb200_winpf.so 5" "libabb200_winpf.so 4" "libabb200.so 4"; do
  set -- $CFG
  ABB_LIB=$1 ABB_BLOCK_TIERS=0 ABB_S1_MINB=$2 timeout 300 python bench.py --workload L --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2d_bench_$1_m$2.json 2> gpurun_out/r2d_bench_$1_m$2.err
  python -c "
import json; d=json.load(open('gpurun_out/r2d_bench_$1_m$2.json')); print('RESULT LIB=$1 MINB=$2', d['ms_per_step'], d['walk_ms_per_step'], d['paths_ms_per_step'], d['e2e']['value']/1e6)"
done
ABB_BLOCK_TIERS=0 timeout 900 ncu --set full --clock-control none --import-source on -k regex:walk_smem_kernel -c 1 -o gpurun_out/r2d_s1_default python bench.py --workload L --steps 1 --warmup 3 --no-cpu-baseline --check 0 > gpurun_out/r2d_ncu.log 2>&1
ls -la gpurun_out/
