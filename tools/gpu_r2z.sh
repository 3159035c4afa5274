#!/bin/bash
# Round-2 call z (1 GPU): deconvs with a deep smem pipeline (8 stages) x {4, 8} epilogue warps, CC merge without the implied
# unions; parity, A/B timings, launch list, CC traffic.
# usage (under gpurun): bash tools/gpu_r2z.sh [tag]
mkdir -p gpurun_out
tag=${1:-r2z}
timeout 600 python -m pytest tests/ -q -m gpu -p no:cacheprovider > gpurun_out/pytest_${tag}.log 2>&1; t=$?; echo "pytest exit $t"; tail -15 gpurun_out/pytest_${tag}.log
if [ $t -ne 0 ]; then
  DLV_DECONV_EPI=1 DLV_DECONV_STAGES=2 DLV_CCL_PRUNE=0 timeout 600 python -m pytest tests/ -q -m gpu -p no:cacheprovider > gpurun_out/pytest_fallback_${tag}.log 2>&1; echo "pytest (previous kernels) exit $?"; tail -8 gpurun_out/pytest_fallback_${tag}.log
fi
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_${tag}.log 2>&1; echo "smoke exit $?"; tail -2 gpurun_out/smoke_${tag}.log
timeout 600 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_${tag}.json 2> gpurun_out/bench_${tag}.err; echo "bench exit $?"; cat gpurun_out/bench_${tag}.json; tail -3 gpurun_out/bench_${tag}.err
DLV_DECONV_EPI=1 timeout 300 python bench.py --steps 2 --warmup 2 --no-cpu-baseline > gpurun_out/bench_epi1_${tag}.json 2> gpurun_out/bench_epi1_${tag}.err; echo "bench EPI=1 (8 stages) exit $?"
python -c "import json,sys; j=json.loads(open('gpurun_out/bench_epi1_${tag}.json').read().strip().splitlines()[-1]); print('EPI=1 stages=8', j['value'], j['ms_per_step'], j['roofline']['conv_ms_per_step'], j['roofline']['frac'])"
timeout 300 python bench.py --workload cfg3 --steps 3 --warmup 2 > gpurun_out/bench_cfg3_${tag}.json 2> gpurun_out/bench_cfg3_${tag}.err; echo "cfg3 exit $?"; cat gpurun_out/bench_cfg3_${tag}.json
DLV_CCL_PRUNE=0 timeout 300 python bench.py --workload cfg3 --steps 3 --warmup 2 > gpurun_out/bench_cfg3_noprune_${tag}.json 2> gpurun_out/bench_cfg3_noprune_${tag}.err; echo "cfg3 (all unions) exit $?"; cat gpurun_out/bench_cfg3_noprune_${tag}.json
[ "$2" = "quick" ] && exit 0
K='regex:conv_tc|conv_is|is_reduce|norm_mish|final_blend|gather_windows|window_active|average_|cell_table|erode_|ccl_|scan_|bbox_init|relabel|boundary|paint_|edt_'
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" --launch-skip 60 -c 120 --csv --log-file gpurun_out/launches_cfg2_${tag}.csv python bench.py --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/ncu_list_cfg2_${tag}.log 2>&1; echo "ncu list exit $?"
python tools/ncu_summary.py launches gpurun_out/launches_cfg2_${tag}.csv > gpurun_out/launches_cfg2_${tag}.txt 2>&1; head -14 gpurun_out/launches_cfg2_${tag}.txt
DLV_BENCH_CFG3_MIN_MS=0 timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k 'regex:ccl_|scan_|bbox_|paint_' --launch-skip 18 -c 12 --csv --log-file gpurun_out/ccl_traffic_${tag}.csv python bench.py --workload cfg3 --steps 1 --warmup 0 > gpurun_out/ncu_ccl_${tag}.log 2>&1; echo "ncu ccl exit $?"
