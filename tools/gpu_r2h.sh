#!/bin/bash
# round 2, call h: split wait counters of the fused conv, the round-1 MMA microbenchmarks (issue-stream ingredients), ncu
# launch list of one 128-window batch of cfg2
mkdir -p gpurun_out
tag=${1:-r2h}
DLV_IS_DEBUG=1 timeout 300 python bench.py --workload small --steps 1 --warmup 1 --no-cpu-baseline 2> gpurun_out/${tag}_isdbg.txt > /dev/null
grep "^\[is\]" gpurun_out/${tag}_isdbg.txt | head -8 | cut -c1-60,88-
for b in umma_bench2 umma_bench3; do timeout 120 tools/ubench/$b > gpurun_out/${tag}_$b.txt 2>&1; echo "$b exit $?"; cat gpurun_out/${tag}_$b.txt; done
K='regex:conv_tc|conv_is|is_reduce|norm_mish|final_blend|gather_windows|window_active|average_|cell_table|erode_|ccl_|scan_|bbox_init|relabel|boundary|paint_|edt_'
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" --launch-skip 52 -c 104 --csv --log-file gpurun_out/launches_cfg2_${tag}.csv python bench.py --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/ncu_list_cfg2_${tag}.log 2>&1; echo "ncu list exit $?"
python tools/ncu_summary.py launches gpurun_out/launches_cfg2_${tag}.csv 2>/dev/null | head -30
