#!/bin/bash
mkdir -p gpurun_out
tag=${1:-e5}
bash tools/gpu_ci.sh > gpurun_out/ci_${tag}.log 2>&1; echo "ci exit $?"; grep -E "^===|passed|failed|error" gpurun_out/ci_${tag}.log | tail -20
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_${tag}.log 2>&1; echo "smoke exit $?"; tail -2 gpurun_out/smoke_${tag}.log
timeout 900 python bench.py --workload cfg3 --steps 3 --warmup 2 > gpurun_out/bench_cfg3_${tag}.json 2> gpurun_out/bench_cfg3_${tag}.err; echo "cfg3 exit $?"; cat gpurun_out/bench_cfg3_${tag}.json
timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k 'regex:ccl_|scan_|bbox_' -c 9 --csv --log-file gpurun_out/ccl_traffic_${tag}.csv python bench.py --workload cfg3 --steps 1 --warmup 0 > gpurun_out/ncu_ccl_${tag}.log 2>&1; echo "ncu ccl exit $?"
DLV_TRACE=1 timeout 600 python bench.py --steps 1 --warmup 1 --no-cpu-baseline > /dev/null 2> gpurun_out/trace_${tag}.txt; grep "dlv_segment" gpurun_out/trace_${tag}.txt | tail -7
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_${tag}.json 2> gpurun_out/bench_${tag}.err; echo "bench exit $?"; cat gpurun_out/bench_${tag}.json
