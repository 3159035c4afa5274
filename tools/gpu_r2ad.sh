#!/bin/bash
# Round-2 last call (2 GPUs): real-NCCL parity of the slab driver at HEAD and the weak-scaling point without the cfg4 leg.
mkdir -p gpurun_out
tag=${1:-r2ad}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 200 $TR --master-port 29543 tools/nccl_parity.py > gpurun_out/nccl_parity_${tag}.log 2>&1; echo "nccl parity exit $?"; grep "nccl parity" gpurun_out/nccl_parity_${tag}.log
timeout 200 $TR --master-port 29542 bench.py --gpus 2 --steps 2 --warmup 3 --no-cfg4 > gpurun_out/bench_${tag}.json 2> gpurun_out/bench_${tag}.err; echo "bench exit $?"; cat gpurun_out/bench_${tag}.json; tail -2 gpurun_out/bench_${tag}.err
