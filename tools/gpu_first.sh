#!/bin/bash
# first GPU call of a session: parity tests, smoke, bench, ncu launch list
mkdir -p gpurun_out
bash tools/gpu_ci.sh > gpurun_out/ci.log 2>&1; echo "ci exit $?"
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?"; tail -3 gpurun_out/smoke.log
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_cfg2.json 2> gpurun_out/bench_cfg2.err; echo "bench exit $?"; cat gpurun_out/bench_cfg2.json; tail -5 gpurun_out/bench_cfg2.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_small.csv python bench.py --workload small --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_small.log 2>&1; echo "ncu exit $?"
