#!/bin/bash
# Round-2 call ae (1 GPU): BASELINE.json configs[0] (cfg1, the reference's own CPU-runnable case) with BOTH arms on the whole
# workload - a same-configuration ratio (the cfg2 reference arm is a bounded crop).
mkdir -p gpurun_out
tag=${1:-r2ae}
timeout 200 python bench.py --workload cfg1 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_cfg1_${tag}.json 2> gpurun_out/bench_cfg1_${tag}.err; echo "cfg1 exit $?"; cat gpurun_out/bench_cfg1_${tag}.json; tail -2 gpurun_out/bench_cfg1_${tag}.err
timeout 200 python bench.py --workload cfg1 --tta --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/bench_cfg1_tta_${tag}.json 2> gpurun_out/bench_cfg1_tta_${tag}.err; echo "cfg1 tta exit $?"; cat gpurun_out/bench_cfg1_tta_${tag}.json
timeout 400 python bench.py --impl reference --workload cfg1 --steps 1 --warmup 0 > gpurun_out/bench_cfg1_ref_${tag}.json 2> gpurun_out/bench_cfg1_ref_${tag}.err; echo "cfg1 reference exit $?"; cat gpurun_out/bench_cfg1_ref_${tag}.json
