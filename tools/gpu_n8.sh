#!/bin/bash
# 8-GPU call: the whole-brain configuration (BASELINE.json configs[3]) and the weak-scaling point N=8.
mkdir -p gpurun_out
tag=${1:-n8}
N=${2:-8}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
nvidia-smi --query-gpu=index,name,memory.total --format=csv > gpurun_out/gpus_${tag}.txt 2>&1
timeout 900 $TR --master-port 29541 bench.py --gpus $N --workload cfg4 --steps 1 --warmup 1 > gpurun_out/bench_cfg4_${tag}.json 2> gpurun_out/bench_cfg4_${tag}.err; echo "cfg4 exit $?"; cat gpurun_out/bench_cfg4_${tag}.json; tail -4 gpurun_out/bench_cfg4_${tag}.err | cut -c1-400
timeout 600 $TR --master-port 29542 bench.py --gpus $N --steps 2 --warmup 3 > gpurun_out/bench_${tag}.json 2> gpurun_out/bench_${tag}.err; echo "bench exit $?"; cat gpurun_out/bench_${tag}.json
timeout 300 $TR --master-port 29543 tools/nccl_parity.py > gpurun_out/nccl_parity_${tag}.log 2>&1; echo "nccl parity exit $?"; grep "nccl parity" gpurun_out/nccl_parity_${tag}.log
