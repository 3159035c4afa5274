#!/bin/bash
# one GPU call of the build -> measure loop: all parity tests, smoke, bench (+ env variants), ncu launch list
# usage: gpu_round.sh tag ["ENV=..." variants...]
mkdir -p gpurun_out
tag=${1:-cur}; shift
bash tools/gpu_ci.sh > gpurun_out/ci_${tag}.log 2>&1; echo "ci exit $?"; grep -E "^===|passed|failed|error" gpurun_out/ci_${tag}.log | tail -14
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_${tag}.log 2>&1; echo "smoke exit $?"; tail -2 gpurun_out/smoke_${tag}.log
bash tools/gpu_bench.sh "DLV_X=0" "$@" 2>&1 | tee gpurun_out/benchsum_${tag}.txt
K='regex:conv_tc|conv_is|is_reduce|norm_mish|final_blend|gather_windows|window_active|average_kernel|erode_|ccl_|scan_|bbox_init|relabel|boundary'
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" -c 400 --csv --log-file gpurun_out/launches_${tag}.csv python bench.py --workload small --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_list_${tag}.log 2>&1; echo "ncu list exit $?"
