set -x
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
export ABB_LIB=libabb200_ck.so
(timeout 240 python -m pytest tests/test_gpu_chunked.py tests/test_gpu_hist_pack.py -x -q -m gpu > gpurun_out/r2n_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2n_pytest.log; tail -15 gpurun_out/r2n_pytest.log | cut -c1-300)
(timeout 100 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1)
timeout 200 python profiles/chunk_check.py > gpurun_out/r02_chunk_check.json 2> gpurun_out/r02_chunk_check.err
grep chunk_check gpurun_out/r02_chunk_check.err | cut -c1-400; tail -3 gpurun_out/r02_chunk_check.err | cut -c1-300
ABB_TRACE=1 timeout 300 python bench.py --workload L --steps 5 --warmup 3 > gpurun_out/r2n_bench_L.json 2> gpurun_out/r2n_bench_L.err
grep "exposure_host" gpurun_out/r2n_bench_L.err | tail -2
tail -2 gpurun_out/r2n_bench_L.err | cut -c1-300
python -c "
import json; d=json.load(open('gpurun_out/r2n_bench_L.json')); print('RESULT L', d['value']/1e6, d['ms_per_step'], 'e2e', d['e2e']['value']/1e6, d['e2e']['ms_per_step'], d['e2e']['d2h_bytes_per_step'], d['e2e']['first_call_ms'], d['e2e']['python_zero_copy'], d['roofline']['frac'])"
