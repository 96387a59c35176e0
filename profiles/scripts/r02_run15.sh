set -x
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
export ABB_LIB=libabb200_ck.so
(timeout 150 python -m pytest tests/test_gpu_chunked.py tests/test_gpu_hist_pack.py -x -q -m gpu > gpurun_out/r2o_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2o_pytest.log; tail -12 gpurun_out/r2o_pytest.log | cut -c1-300)
timeout 170 python profiles/chunk_check.py > gpurun_out/r02_chunk_check.json 2> gpurun_out/r02_chunk_check.err
grep chunk_check gpurun_out/r02_chunk_check.err | cut -c1-420; tail -3 gpurun_out/r02_chunk_check.err | cut -c1-300
