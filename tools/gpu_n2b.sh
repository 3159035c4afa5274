#!/bin/bash
mkdir -p gpurun_out
tag=${1:-n2b}
N=2
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29551 bench.py --gpus $N --workload cfg4s --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/bench_cfg4s_${tag}.json 2> gpurun_out/bench_cfg4s_${tag}.err; echo "cfg4s exit $?"; cat gpurun_out/bench_cfg4s_${tag}.json; tail -3 gpurun_out/bench_cfg4s_${tag}.err | cut -c1-300
timeout 900 $TR --master-port 29552 bench.py --gpus $N --steps 2 --warmup 3 > gpurun_out/bench_${tag}.json 2> gpurun_out/bench_${tag}.err; echo "bench exit $?"; cat gpurun_out/bench_${tag}.json
timeout 600 python -m pytest tests/test_gpu_e_segment.py tests/test_gpu_c_ccl.py -q -m gpu -p no:cacheprovider > gpurun_out/ci_${tag}.log 2>&1; echo "tests exit $?"; tail -3 gpurun_out/ci_${tag}.log
