# N-GPU legs (run with gpurun --gpus N): exposure bench, enumerate bench, dependency reach across ranks
set -x
N=${1:-2}
LEGS=${2:-exposure,enumerate,reach}
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533"
if [[ $LEGS == *exposure* ]]; then
timeout 600 $TR bench.py --gpus $N --workload L --steps 5 --warmup 3 > gpurun_out/r02_bench_L_${N}gpu.json 2> gpurun_out/r02_bench_L_${N}gpu.err
tail -3 gpurun_out/r02_bench_L_${N}gpu.err | cut -c1-300
python -c "
import json; d=json.load(open('gpurun_out/r02_bench_L_${N}gpu.json')); print('RESULT N=$N exposure', d['value']/1e6, d['ms_per_step'], d['per_rank'], d['e2e']['value']/1e6, d['config']['csr_broadcast'])"
fi
if [[ $LEGS == *enumerate* ]]; then
timeout 600 $TR bench.py --gpus $N --workload L --mode enumerate --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_L_enumerate_${N}gpu.json 2> gpurun_out/r02_bench_L_enumerate_${N}gpu.err
tail -3 gpurun_out/r02_bench_L_enumerate_${N}gpu.err | cut -c1-300
python -c "
import json; d=json.load(open('gpurun_out/r02_bench_L_enumerate_${N}gpu.json')); print('RESULT N=$N enumerate', d['value']/1e6, d['ms_per_step'], d['enumerate']['per_walk_ms'], d['e2e']['value']/1e6)"
fi
if [[ $LEGS == *reach* ]]; then
timeout 600 $TR profiles/reach_bench.py --workload L > gpurun_out/r02_reach_L_${N}gpu.json 2> gpurun_out/r02_reach_L_${N}gpu.err
tail -3 gpurun_out/r02_reach_L_${N}gpu.err | cut -c1-300; cat gpurun_out/r02_reach_L_${N}gpu.json | cut -c1-600

fi
