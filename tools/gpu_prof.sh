#!/bin/bash
# ncu evidence for the current build: launch list of our kernels + full captures of the dominant ones
mkdir -p gpurun_out
K='regex:conv_tc|norm_mish|final_blend|gather_windows|window_active|average_kernel|erode_|ccl_|scan_|bbox_init|relabel|boundary'
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" -c 400 --csv --log-file gpurun_out/launches_small.csv python bench.py --workload small --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_small.log 2>&1; echo "ncu list exit $?"
# full capture: warm-up pass = 1 gather + 18 conv + 4 deconv ...; skip the first pass, take the L0 Cout=32 convs
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:conv_tc -s 22 -c 22 -o gpurun_out/prof_conv python bench.py --workload small --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_conv.log 2>&1; echo "ncu conv exit $?"
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:'norm_mish|final_blend|gather_windows' -s 18 -c 19 -o gpurun_out/prof_elem python bench.py --workload small --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_elem.log 2>&1; echo "ncu elem exit $?"
ls -la gpurun_out
