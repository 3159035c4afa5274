#!/bin/bash
# round 2, call k (gpurun --gpus N): NCCL parity of the slab driver, the driver's weak-scaling line with its cfg4 leg, one
# traced cfg4 step (per-phase times of every rank), the cfg5 sweep when asked for
# usage: bash tools/gpu_r2k.sh tag N [cfg5]
mkdir -p gpurun_out
tag=${1:-r2k}; N=${2:-2}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
nvidia-smi --query-gpu=index,name,memory.total --format=csv > gpurun_out/gpus_${tag}.txt 2>&1
timeout 300 $TR --master-port 29543 tools/nccl_parity.py > gpurun_out/nccl_parity_${tag}.log 2>&1; echo "nccl parity exit $?"; grep "nccl parity" gpurun_out/nccl_parity_${tag}.log
timeout 900 $TR --master-port 29542 bench.py --gpus $N --steps 2 --warmup 1 > gpurun_out/bench_${tag}.json 2> gpurun_out/bench_${tag}.err; echo "bench exit $?"; cat gpurun_out/bench_${tag}.json; tail -3 gpurun_out/bench_${tag}.err
if [ "$3" = "cfg5" ]; then
  timeout 1200 $TR --master-port 29546 bench.py --gpus $N --workload cfg5 --steps 2 --warmup 1 > gpurun_out/bench_cfg5_${tag}.json 2> gpurun_out/bench_cfg5_${tag}.err; echo "cfg5 exit $?"; head -c 1500 gpurun_out/bench_cfg5_${tag}.json; tail -3 gpurun_out/bench_cfg5_${tag}.err
fi
