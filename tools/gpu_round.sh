#!/bin/bash
# One 1-GPU call of the build -> measure loop: parity tests (one process per file AND the driver's single-process
# form), smoke, bench (ours + reference arm), cfg3 CC/painter bench, ncu launch list of one cfg2 batch, DRAM traffic of
# the conv and CC kernels, one ncu --set full capture of the dominant kernel.
# usage (under gpurun): bash tools/gpu_round.sh tag [quick]
mkdir -p gpurun_out
tag=${1:-cur}
bash tools/gpu_ci.sh > gpurun_out/ci_${tag}.log 2>&1; echo "ci exit $?"; grep -E "^===|passed|failed|error" gpurun_out/ci_${tag}.log | tail -20
timeout 300 python -m pytest tests/ -x -q -m gpu > gpurun_out/pytest_single_${tag}.log 2>&1; echo "single-process pytest exit $?"; tail -2 gpurun_out/pytest_single_${tag}.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_${tag}.log 2>&1; echo "smoke exit $?"; tail -2 gpurun_out/smoke_${tag}.log
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_${tag}.json 2> gpurun_out/bench_${tag}.err; echo "bench exit $?"; cat gpurun_out/bench_${tag}.json
timeout 600 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/bench_ref_${tag}.json 2> gpurun_out/bench_ref_${tag}.err; echo "ref exit $?"; cat gpurun_out/bench_ref_${tag}.json
timeout 900 python bench.py --workload cfg3 --steps 3 --warmup 2 > gpurun_out/bench_cfg3_${tag}.json 2> gpurun_out/bench_cfg3_${tag}.err; echo "cfg3 exit $?"; cat gpurun_out/bench_cfg3_${tag}.json
[ "$2" = "quick" ] && exit 0
K='regex:conv_tc|conv_is|is_reduce|norm_mish|final_blend|gather_windows|window_active|average_|cell_table|erode_|ccl_|scan_|bbox_init|relabel|boundary|paint_|edt_'
# every kernel of the second 128-window batch of cfg2 (skip the first batch's launches), time only
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" --launch-skip 60 -c 120 --csv --log-file gpurun_out/launches_cfg2_${tag}.csv python bench.py --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/ncu_list_cfg2_${tag}.log 2>&1; echo "ncu list exit $?"
timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:conv_ --launch-skip 22 -c 22 --csv --log-file gpurun_out/conv_traffic_${tag}.csv python bench.py --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/ncu_traffic_${tag}.log 2>&1; echo "ncu traffic exit $?"
timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k 'regex:ccl_|scan_|bbox_|paint_' -c 14 --csv --log-file gpurun_out/ccl_traffic_${tag}.csv python bench.py --workload cfg3 --steps 1 --warmup 0 > gpurun_out/ncu_ccl_${tag}.log 2>&1; echo "ncu ccl exit $?"
# one ncu --set full capture of the dominant kernel on the HEADLINE workload (cfg2): the 8 conv_is launches of the second batch
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:conv_is --launch-skip 8 -c 8 -o gpurun_out/prof_is_${tag} python bench.py --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/ncu_is_${tag}.log 2>&1; echo "ncu is exit $?"
