#!/bin/bash
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_gpu_h_paint.py tests/test_gpu_f_slabs.py -q -m gpu -p no:cacheprovider > gpurun_out/ci_last.log 2>&1; echo "tests exit $?"; tail -2 gpurun_out/ci_last.log
timeout 200 python bench.py --workload cfg3 --steps 3 --warmup 2 > gpurun_out/bench_cfg3_last.json 2> gpurun_out/bench_cfg3_last.err; echo "cfg3 exit $?"; cat gpurun_out/bench_cfg3_last.json; tail -2 gpurun_out/bench_cfg3_last.err
