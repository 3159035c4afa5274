#!/bin/bash
mkdir -p gpurun_out
tag=${1:-e2}
bash tools/gpu_ci.sh > gpurun_out/ci_${tag}.log 2>&1; echo "ci exit $?"; grep -E "^===|passed|failed|error" gpurun_out/ci_${tag}.log | tail -20
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_${tag}.log 2>&1; echo "smoke exit $?"; tail -2 gpurun_out/smoke_${tag}.log
timeout 900 python bench.py --workload cfg3 --steps 3 --warmup 2 > gpurun_out/bench_cfg3_${tag}.json 2> gpurun_out/bench_cfg3_${tag}.err; echo "cfg3 exit $?"; cat gpurun_out/bench_cfg3_${tag}.json
timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k 'regex:ccl_|scan_|bbox_' -c 9 --csv --log-file gpurun_out/ccl_traffic_${tag}.csv python bench.py --workload cfg3 --steps 1 --warmup 0 > gpurun_out/ncu_ccl_${tag}.log 2>&1; echo "ncu ccl exit $?"
bash tools/gpu_bench.sh "DLV_X=0" "DLV_WINDOW_BATCH=64" "DLV_LIB=$PWD/delivr_cfos_b200/libdelivr_b200_xw12.so" "DLV_LIB=$PWD/delivr_cfos_b200/libdelivr_b200_np1.so" "DLV_LIB=$PWD/delivr_cfos_b200/libdelivr_b200_np2.so" 2>&1 | tee gpurun_out/benchsum_${tag}.txt
