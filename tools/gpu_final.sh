#!/bin/bash
mkdir -p gpurun_out
tag=${1:-fin}
bash tools/gpu_ci.sh > gpurun_out/ci_${tag}.log 2>&1; echo "ci exit $?"; grep -E "^===|passed|failed|error" gpurun_out/ci_${tag}.log | tail -20
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_${tag}.log 2>&1; echo "smoke exit $?"; tail -2 gpurun_out/smoke_${tag}.log
timeout 600 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_${tag}.json 2> gpurun_out/bench_${tag}.err; echo "bench exit $?"; cat gpurun_out/bench_${tag}.json
timeout 300 python bench.py --workload cfg3 --steps 3 --warmup 2 > gpurun_out/bench_cfg3_${tag}.json 2> gpurun_out/bench_cfg3_${tag}.err; echo "cfg3 exit $?"; cat gpurun_out/bench_cfg3_${tag}.json; tail -2 gpurun_out/bench_cfg3_${tag}.err
timeout 300 python bench.py --workload cfg1 --tta --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_cfg1_tta_${tag}.json 2> gpurun_out/bench_cfg1_tta_${tag}.err; echo "cfg1 tta exit $?"; cat gpurun_out/bench_cfg1_tta_${tag}.json
