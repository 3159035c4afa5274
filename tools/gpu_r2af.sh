#!/bin/bash
# Round-2 call af (1 GPU box, CPU work): the reference arm with torch.cuda.is_available() masked while the reference runs
# (it had been moving its window batches to the GPU through DataParallel): cfg2 crop and the WHOLE cfg1 workload.
mkdir -p gpurun_out
tag=${1:-r2af}
timeout 200 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/bench_ref_${tag}.json 2> gpurun_out/bench_ref_${tag}.err; echo "reference (cfg2 crop) exit $?"; cat gpurun_out/bench_ref_${tag}.json; tail -2 gpurun_out/bench_ref_${tag}.err
timeout 400 python bench.py --impl reference --workload cfg1 --steps 1 --warmup 0 > gpurun_out/bench_cfg1_ref_${tag}.json 2> gpurun_out/bench_cfg1_ref_${tag}.err; echo "reference (whole cfg1) exit $?"; cat gpurun_out/bench_cfg1_ref_${tag}.json; tail -2 gpurun_out/bench_cfg1_ref_${tag}.err
