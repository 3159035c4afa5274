#!/bin/bash
mkdir -p gpurun_out
tag=${1:-n2c}
N=2
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29561 tools/nccl_parity.py > gpurun_out/nccl_parity_${tag}.log 2>&1; echo "nccl parity exit $?"; grep "nccl parity" gpurun_out/nccl_parity_${tag}.log
timeout 900 $TR --master-port 29562 bench.py --gpus $N --steps 2 --warmup 3 > gpurun_out/bench_${tag}.json 2> gpurun_out/bench_${tag}.err; echo "bench exit $?"; cat gpurun_out/bench_${tag}.json; grep -i "error" gpurun_out/bench_${tag}.err | head -5
timeout 600 python -m pytest tests/test_gpu_c_ccl.py tests/test_gpu_f_slabs.py tests/test_gpu_e_segment.py -q -m gpu -p no:cacheprovider > gpurun_out/ci_${tag}.log 2>&1; echo "tests exit $?"; tail -4 gpurun_out/ci_${tag}.log
for v in 0 1; do DLV_CCL_SPARSE_INIT=$v timeout 600 python bench.py --workload cfg3 --steps 3 --warmup 2 > gpurun_out/bench_cfg3_sparse${v}_${tag}.json 2>/dev/null; python -c "
import json; d=json.load(open('gpurun_out/bench_cfg3_sparse${v}_${tag}.json')); print('sparse_init=$v', 'ms/step', round(d['ms_per_step'],2), 'kernels', round(d['roofline']['kernels_ms_per_step'],2), 'n', d['config']['components'])"; done
