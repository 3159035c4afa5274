#!/bin/bash
# Round-2 final 1-GPU record at HEAD: the driver's pytest form, smoke, bench (ours + reference arm), cfg3.
mkdir -p gpurun_out
tag=${1:-r2ac}
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu_${tag}.txt 2>&1
timeout 600 python -m pytest tests/ -x -q -m gpu > gpurun_out/pytest_${tag}.log 2>&1; echo "pytest exit $?"; tail -4 gpurun_out/pytest_${tag}.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_${tag}.log 2>&1; echo "smoke exit $?"; tail -2 gpurun_out/smoke_${tag}.log
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_${tag}.json 2> gpurun_out/bench_${tag}.err; echo "bench exit $?"; cat gpurun_out/bench_${tag}.json; tail -3 gpurun_out/bench_${tag}.err
timeout 600 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/bench_ref_${tag}.json 2> gpurun_out/bench_ref_${tag}.err; echo "ref exit $?"; cat gpurun_out/bench_ref_${tag}.json
timeout 300 python bench.py --workload cfg3 --steps 3 --warmup 2 > gpurun_out/bench_cfg3_${tag}.json 2> gpurun_out/bench_cfg3_${tag}.err; echo "cfg3 exit $?"; cat gpurun_out/bench_cfg3_${tag}.json
