#!/bin/bash
# ncu evidence for the current build: launch list of our kernels + full captures of the dominant ones
# usage: gpu_prof.sh [tag] [env assignments...]
mkdir -p gpurun_out
tag=${1:-cur}; shift
K='regex:conv_tc|conv_is|is_reduce|norm_mish|final_blend|gather_windows|window_active|average_kernel|erode_|ccl_|scan_|bbox_init|relabel|boundary'
env "$@" timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" -c 400 --csv --log-file gpurun_out/launches_${tag}.csv python bench.py --workload small --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_list_${tag}.log 2>&1; echo "ncu list exit $?"
env "$@" timeout 1200 ncu --set full --clock-control none --import-source on -k regex:conv_is -c 8 -o gpurun_out/prof_is_${tag} python bench.py --workload small --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_is_${tag}.log 2>&1; echo "ncu is exit $?"
