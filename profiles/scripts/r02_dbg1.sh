cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
for T in 0 1 2 3; do
  (ABB_BLOCK_TIERS=$T timeout 400 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "test_impact_many and key3" > gpurun_out/dbg1_t$T.log 2>&1; echo "TIERS=$T rc=$?"; grep -E "passed|failed|Mismatched" gpurun_out/dbg1_t$T.log | head -3)
done
