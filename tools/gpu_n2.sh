#!/bin/bash
# 2-GPU call: real-NCCL parity of the slab driver, the N=2 bench (ours + reference arm).
mkdir -p gpurun_out
tag=${1:-n2}
N=${2:-2}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29533 tools/nccl_parity.py > gpurun_out/nccl_parity_${tag}.log 2>&1; echo "nccl parity exit $?"; grep "nccl parity" gpurun_out/nccl_parity_${tag}.log
timeout 900 $TR --master-port 29534 bench.py --gpus $N --steps 2 --warmup 3 > gpurun_out/bench_${tag}.json 2> gpurun_out/bench_${tag}.err; echo "bench exit $?"; cat gpurun_out/bench_${tag}.json; tail -3 gpurun_out/bench_${tag}.err
timeout 600 $TR --master-port 29535 bench.py --impl reference --gpus $N --steps 1 --warmup 0 > gpurun_out/bench_ref_${tag}.json 2> gpurun_out/bench_ref_${tag}.err; echo "ref exit $?"; cat gpurun_out/bench_ref_${tag}.json
