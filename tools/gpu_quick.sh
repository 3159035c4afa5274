#!/bin/bash
# Short 1-GPU call: parity tests (one process per file), smoke, a short bench; optional follow-up script when all green.
# usage (under gpurun): bash tools/gpu_quick.sh tag [follow-up script args...]
mkdir -p gpurun_out
tag=${1:-q}; shift
bash tools/gpu_ci.sh > gpurun_out/ci_${tag}.log 2>&1; ci=$?; echo "ci exit $ci"; grep -E "^===|passed|failed|error|timed out" gpurun_out/ci_${tag}.log | tail -24
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_${tag}.log 2>&1; echo "smoke exit $?"; tail -2 gpurun_out/smoke_${tag}.log
timeout 600 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_${tag}.json 2> gpurun_out/bench_${tag}.err; echo "bench exit $?"; cat gpurun_out/bench_${tag}.json; tail -3 gpurun_out/bench_${tag}.err
if [ $ci -eq 0 ] && [ -n "$1" ]; then bash "$@"; fi
