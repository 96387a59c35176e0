set -x
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
(timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_tiers.py -x -q -m gpu -k "test_impact_every_tier and shrunk and key0" > gpurun_out/r02_sanitizer_racecheck_tiers.log 2>&1; echo "racecheck rc=$?" >> gpurun_out/r02_sanitizer_racecheck_tiers.log; tail -4 gpurun_out/r02_sanitizer_racecheck_tiers.log)
(timeout 600 python -m pytest tests/test_gpu_tiers.py -x -q -m gpu -k "impact" > gpurun_out/r2j_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2j_pytest.log; tail -3 gpurun_out/r2j_pytest.log)
timeout 400 python bench.py --workload L --steps 5 --warmup 3 > gpurun_out/r2j_bench_L_overlap.json 2> gpurun_out/r2j_bench_L_overlap.err
tail -3 gpurun_out/r2j_bench_L_overlap.err | cut -c1-300
python -c "
import json; d=json.load(open('gpurun_out/r2j_bench_L_overlap.json')); print('RESULT overlap', d['value']/1e6, d['ms_per_step'], d['sequential_ms_per_step'], d['walk_ms_per_step'], d['paths_ms_per_step'], d['e2e']['value']/1e6, d['roofline']['frac'])"
timeout 400 python bench.py --workload L --steps 5 --warmup 3 --no-overlap --no-cpu-baseline > gpurun_out/r2j_bench_L_seq.json 2> gpurun_out/r2j_bench_L_seq.err
python -c "
import json; d=json.load(open('gpurun_out/r2j_bench_L_seq.json')); print('RESULT sequential', d['value']/1e6, d['ms_per_step'], d['sequential_ms_per_step'], d['walk_ms_per_step'], d['paths_ms_per_step'], d['e2e']['value']/1e6)"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"walk_smem_kernel|walk_global_kernel|walk_block_kernel" -c 3 -o gpurun_out/r02_walk_final python bench.py --workload L --steps 1 --warmup 3 --no-cpu-baseline --check 0 --no-overlap > gpurun_out/r2j_ncu.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 130 -c 90 --csv --log-file gpurun_out/r02_launches_L_step.csv python bench.py --workload L --steps 2 --warmup 3 --no-cpu-baseline --check 0 --no-overlap > gpurun_out/r2j_ncu2.log 2>&1
ls -la gpurun_out/r02_walk_final.ncu-rep gpurun_out/r02_launches_L_step.csv
