set -x
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_scale_properties.py -x -q -m gpu -k "impact or distances or dedup or monoton or replay" > gpurun_out/r2k_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2k_pytest.log; tail -3 gpurun_out/r2k_pytest.log)
run() { name=$1; shift
  env "$@" timeout 300 python bench.py --workload L --steps 5 --warmup 3 --no-cpu-baseline --no-overlap > gpurun_out/r2k_bench_$name.json 2> gpurun_out/r2k_bench_$name.err
  python -c "
import json; d=json.load(open('gpurun_out/r2k_bench_$name.json')); print('RESULT $name', d['ms_per_step'], d['walk_ms_per_step'], d['paths_ms_per_step'], d['e2e']['value']/1e6)"
}
run locality1 ABB_LOCALITY=1
run locality0 ABB_LOCALITY=0
run locality1_g48 ABB_LOCALITY=1 ABB_G1_WARPS_PER_SM=48
ABB_LOCALITY=1 timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct --clock-control none -k regex:"walk_smem_kernel|walk_global_kernel" -c 3 --csv --log-file gpurun_out/r2k_walk_metrics.csv python bench.py --workload L --steps 1 --warmup 3 --no-cpu-baseline --check 0 --no-overlap > gpurun_out/r2k_ncu.log 2>&1
grep -E "walk_" gpurun_out/r2k_walk_metrics.csv | awk -F'","' '{print substr($5,1,30), $(NF-2), $(NF-1), $NF}' | head -16
