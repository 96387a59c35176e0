set -x
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_gpu_hist_pack.py tests/test_gpu_parity.py tests/test_gpu_api.py tests/test_gpu_scale_properties.py -x -q -m gpu -k "not replay" > gpurun_out/r2m_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2m_pytest.log; tail -4 gpurun_out/r2m_pytest.log)
(timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1)
ABB_TRACE=1 timeout 500 python bench.py --workload L --steps 5 --warmup 3 > gpurun_out/r02_bench_L_1gpu.json 2> gpurun_out/r02_bench_L_1gpu.err
grep "exposure_host" gpurun_out/r02_bench_L_1gpu.err | tail -2
tail -2 gpurun_out/r02_bench_L_1gpu.err | cut -c1-300
python -c "
import json; d=json.load(open('gpurun_out/r02_bench_L_1gpu.json')); print('RESULT L', d['value']/1e6, d['ms_per_step'], d['walk_ms_per_step'], d['paths_ms_per_step'], 'e2e', d['e2e']['value']/1e6, d['e2e']['ms_per_step'], d['e2e']['d2h_bytes_per_step'], d['e2e']['python_zero_copy'], d['roofline']['frac'])"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"walk_smem_kernel|walk_global_kernel|walk_block_kernel" -c 3 -o gpurun_out/r02_walk_final python bench.py --workload L --steps 1 --warmup 3 --no-cpu-baseline --check 0 --no-overlap > gpurun_out/r2m_ncu.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 130 -c 100 --csv --log-file gpurun_out/r02_launches_L_step.csv python bench.py --workload L --steps 2 --warmup 3 --no-cpu-baseline --check 0 --no-overlap > gpurun_out/r2m_ncu2.log 2>&1
ls -la gpurun_out/r02_walk_final.ncu-rep gpurun_out/r02_launches_L_step.csv
