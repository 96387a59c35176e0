cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 70 python -m pytest tests/test_gpu_chunked.py -x -q -m gpu -k "exposure or other_walks or estate_dense" > gpurun_out/r2p_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2p_pytest.log; tail -5 gpurun_out/r2p_pytest.log | cut -c1-300
