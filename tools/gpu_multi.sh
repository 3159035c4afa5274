#!/bin/bash
# N-GPU call (gpurun --gpus N): real-NCCL parity of the slab driver, the weak-scaling point, the reference arm and -
# at N = 8 - the whole-brain configuration (BASELINE.json configs[3]).
# usage: bash tools/gpu_multi.sh tag N
mkdir -p gpurun_out
tag=${1:-multi}
N=${2:-2}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
nvidia-smi --query-gpu=index,name,memory.total --format=csv > gpurun_out/gpus_${tag}.txt 2>&1
timeout 300 $TR --master-port 29543 tools/nccl_parity.py > gpurun_out/nccl_parity_${tag}.log 2>&1; echo "nccl parity exit $?"; grep "nccl parity" gpurun_out/nccl_parity_${tag}.log
if [ "$N" = "8" ]; then
  timeout 900 $TR --master-port 29541 bench.py --gpus $N --workload cfg4 --steps 1 --warmup 1 > gpurun_out/bench_cfg4_${tag}.json 2> gpurun_out/bench_cfg4_${tag}.err; echo "cfg4 exit $?"; cat gpurun_out/bench_cfg4_${tag}.json
fi
timeout 600 $TR --master-port 29542 bench.py --gpus $N --steps 2 --warmup 3 > gpurun_out/bench_${tag}.json 2> gpurun_out/bench_${tag}.err; echo "bench exit $?"; cat gpurun_out/bench_${tag}.json
timeout 600 $TR --master-port 29544 bench.py --impl reference --gpus $N --steps 1 --warmup 0 > gpurun_out/bench_ref_${tag}.json 2> gpurun_out/bench_ref_${tag}.err; echo "ref exit $?"; cat gpurun_out/bench_ref_${tag}.json
