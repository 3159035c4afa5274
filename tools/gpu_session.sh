#!/bin/bash
# One GPU call of the build -> measure loop (1 GPU):
#   parity tests, smoke, bench (+ variants), cfg3 CC bench, ncu launch list, conv DRAM traffic, ncu full of conv_is.
# usage: gpu_session.sh tag
mkdir -p gpurun_out
tag=${1:-cur}
bash tools/gpu_ci.sh > gpurun_out/ci_${tag}.log 2>&1; echo "ci exit $?"; grep -E "^===|passed|failed|error" gpurun_out/ci_${tag}.log | tail -16
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_${tag}.log 2>&1; echo "smoke exit $?"; tail -2 gpurun_out/smoke_${tag}.log
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_${tag}.json 2> gpurun_out/bench_${tag}.err; echo "bench exit $?"; cat gpurun_out/bench_${tag}.json
bash tools/gpu_bench.sh "DLV_WINDOW_BATCH=64" "DLV_WINDOW_BATCH=96" "DLV_LIB=$PWD/delivr_cfos_b200/libdelivr_b200_xw12.so" "DLV_LIB=$PWD/delivr_cfos_b200/libdelivr_b200_np1.so" "DLV_LIB=$PWD/delivr_cfos_b200/libdelivr_b200_np2.so" 2>&1 | tee gpurun_out/benchsum_${tag}.txt
timeout 900 python bench.py --workload cfg3 --steps 3 --warmup 2 > gpurun_out/bench_cfg3_${tag}.json 2> gpurun_out/bench_cfg3_${tag}.err; echo "cfg3 exit $?"; cat gpurun_out/bench_cfg3_${tag}.json
timeout 600 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/bench_ref_${tag}.json 2> gpurun_out/bench_ref_${tag}.err; echo "ref exit $?"; cat gpurun_out/bench_ref_${tag}.json
K='regex:conv_tc|conv_is|is_reduce|norm_mish|final_blend|gather_windows|window_active|average_kernel|erode_|ccl_|scan_|bbox_init|relabel|boundary'
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" -c 400 --csv --log-file gpurun_out/launches_${tag}.csv python bench.py --workload small --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_list_${tag}.log 2>&1; echo "ncu list exit $?"
# DRAM traffic of every conv launch of one full 32-window batch of cfg2 (second batch)
timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:conv_ --launch-skip 22 -c 22 --csv --log-file gpurun_out/conv_traffic_${tag}.csv python bench.py --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/ncu_traffic_${tag}.log 2>&1; echo "ncu traffic exit $?"
timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:ccl_\|scan_\|bbox_ -c 8 --csv --log-file gpurun_out/ccl_traffic_${tag}.csv python bench.py --workload cfg3 --steps 1 --warmup 0 > gpurun_out/ncu_ccl_${tag}.log 2>&1; echo "ncu ccl exit $?"
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:conv_is --launch-skip 7 -c 7 -o gpurun_out/prof_is_${tag} python bench.py --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/ncu_is_${tag}.log 2>&1; echo "ncu is exit $?"
