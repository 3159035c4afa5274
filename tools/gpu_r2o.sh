#!/bin/bash
# round 2, call o (gpurun --gpus 8): NCCL parity at 8 ranks, cfg5 window/overlap sweep on 8 GPUs
mkdir -p gpurun_out
tag=${1:-r2o}; N=${2:-8}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
nvidia-smi --query-gpu=index,name,memory.total --format=csv > gpurun_out/gpus_${tag}.txt 2>&1
timeout 300 $TR --master-port 29543 tools/nccl_parity.py > gpurun_out/nccl_parity_${tag}.log 2>&1; echo "nccl parity exit $?"; grep "nccl parity" gpurun_out/nccl_parity_${tag}.log
timeout 800 $TR --master-port 29546 bench.py --gpus $N --workload cfg5 --steps 1 --warmup 1 > gpurun_out/bench_cfg5_${tag}.json 2> gpurun_out/bench_cfg5_${tag}.err; echo "cfg5 exit $?"; head -c 1200 gpurun_out/bench_cfg5_${tag}.json; tail -3 gpurun_out/bench_cfg5_${tag}.err
