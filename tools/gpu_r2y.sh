#!/bin/bash
# Round-2 call y (1 GPU): parity of the new deconv epilogue / scan kernels / box check / box-driven painter resolve,
# A/B timings against the kernels they replace, launch list + conv DRAM traffic of the headline workload.
# usage (under gpurun): bash tools/gpu_r2y.sh [tag]
mkdir -p gpurun_out
tag=${1:-r2y}
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu_${tag}.txt 2>&1
timeout 600 python -m pytest tests/ -q -m gpu -p no:cacheprovider > gpurun_out/pytest_${tag}.log 2>&1; t=$?; echo "pytest exit $t"; tail -15 gpurun_out/pytest_${tag}.log
if [ $t -ne 0 ]; then
  DLV_DECONV_EPI=1 DLV_CCL_BBOX_CHECK=0 DLV_PAINT_RESOLVE=1 timeout 600 python -m pytest tests/ -q -m gpu -p no:cacheprovider > gpurun_out/pytest_fallback_${tag}.log 2>&1; echo "pytest (previous kernels) exit $?"; tail -8 gpurun_out/pytest_fallback_${tag}.log
  DLV_DECONV_EPI=2 timeout 300 python -m pytest tests/test_gpu_a_conv.py tests/test_gpu_b_unet.py -q -m gpu -p no:cacheprovider > gpurun_out/pytest_epi2_${tag}.log 2>&1; echo "pytest (EPI=2) exit $?"; tail -4 gpurun_out/pytest_epi2_${tag}.log
fi
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_${tag}.log 2>&1; echo "smoke exit $?"; tail -2 gpurun_out/smoke_${tag}.log
timeout 600 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_${tag}.json 2> gpurun_out/bench_${tag}.err; echo "bench exit $?"; cat gpurun_out/bench_${tag}.json; tail -3 gpurun_out/bench_${tag}.err
for e in 1; do
  DLV_DECONV_EPI=$e timeout 300 python bench.py --steps 2 --warmup 2 --no-cpu-baseline > gpurun_out/bench_epi${e}_${tag}.json 2> gpurun_out/bench_epi${e}_${tag}.err; echo "bench EPI=$e exit $?"
  python -c "import json,sys; j=json.loads(open('gpurun_out/bench_epi${e}_${tag}.json').read().strip().splitlines()[-1]); print('EPI=$e', j['value'], j['ms_per_step'], j['roofline']['conv_ms_per_step'], j['roofline']['frac'])"
done
timeout 300 python bench.py --workload cfg3 --steps 3 --warmup 2 > gpurun_out/bench_cfg3_${tag}.json 2> gpurun_out/bench_cfg3_${tag}.err; echo "cfg3 exit $?"; cat gpurun_out/bench_cfg3_${tag}.json
DLV_CCL_BBOX_CHECK=0 DLV_PAINT_RESOLVE=1 timeout 300 python bench.py --workload cfg3 --steps 3 --warmup 2 > gpurun_out/bench_cfg3_prev_${tag}.json 2> gpurun_out/bench_cfg3_prev_${tag}.err; echo "cfg3 (previous relabel / painter) exit $?"; cat gpurun_out/bench_cfg3_prev_${tag}.json
[ "$2" = "quick" ] && exit 0
K='regex:conv_tc|conv_is|is_reduce|norm_mish|final_blend|gather_windows|window_active|average_|cell_table|erode_|ccl_|scan_|bbox_init|relabel|boundary|paint_|edt_'
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" --launch-skip 60 -c 120 --csv --log-file gpurun_out/launches_cfg2_${tag}.csv python bench.py --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/ncu_list_cfg2_${tag}.log 2>&1; echo "ncu list exit $?"
python tools/ncu_summary.py launches gpurun_out/launches_cfg2_${tag}.csv > gpurun_out/launches_cfg2_${tag}.txt 2>&1; head -14 gpurun_out/launches_cfg2_${tag}.txt
timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:conv_ --launch-skip 22 -c 22 --csv --log-file gpurun_out/conv_traffic_${tag}.csv python bench.py --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/ncu_traffic_${tag}.log 2>&1; echo "ncu traffic exit $?"
python tools/conv_traffic_summary.py gpurun_out/conv_traffic_${tag}.csv 128 > gpurun_out/conv_traffic_${tag}.txt 2>&1; tail -3 gpurun_out/conv_traffic_${tag}.txt
DLV_BENCH_CFG3_MIN_MS=0 timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k 'regex:ccl_|scan_|bbox_|paint_' --launch-skip 9 -c 13 --csv --log-file gpurun_out/ccl_traffic_${tag}.csv python bench.py --workload cfg3 --steps 1 --warmup 0 > gpurun_out/ncu_ccl_${tag}.log 2>&1; echo "ncu ccl exit $?"
