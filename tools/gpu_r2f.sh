#!/bin/bash
# round 2, call f: per-layer cycle counters of the fused conv, old MMA microbenchmark beside the clustered one, ncu launch
# list of one 128-window batch of cfg2
mkdir -p gpurun_out
tag=${1:-r2f}
DLV_IS_DEBUG=1 timeout 300 python bench.py --workload small --steps 1 --warmup 1 --no-cpu-baseline 2> gpurun_out/${tag}_isdbg.txt > /dev/null
grep "^\[is\]" gpurun_out/${tag}_isdbg.txt | head -8 | cut -c1-60,88-
timeout 120 tools/ubench/umma_bench > gpurun_out/${tag}_umma1.txt 2>&1; echo "umma1 exit $?"; grep -E "^ *(96|64|32|128) " gpurun_out/${tag}_umma1.txt | head -40
K='regex:conv_tc|conv_is|is_reduce|norm_mish|final_blend|gather_windows|window_active|average_|cell_table|erode_|ccl_|scan_|bbox_init|relabel|boundary|paint_|edt_'
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" --launch-skip 52 -c 104 --csv --log-file gpurun_out/launches_cfg2_${tag}.csv python bench.py --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/ncu_list_cfg2_${tag}.log 2>&1; echo "ncu list exit $?"
python tools/ncu_summary.py gpurun_out/launches_cfg2_${tag}.csv 2>/dev/null | head -30
