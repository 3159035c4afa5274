#!/bin/bash
# Round-2 call ab (1 GPU): deconv epilogue with shuffled (fully coalesced) stores; parity, A/B lines, launch list.
mkdir -p gpurun_out
tag=${1:-r2ab}
timeout 600 python -m pytest tests/ -q -m gpu -p no:cacheprovider > gpurun_out/pytest_${tag}.log 2>&1; t=$?; echo "pytest exit $t"; tail -12 gpurun_out/pytest_${tag}.log
line() { python -c "import json,sys; j=json.loads(open('$1').read().strip().splitlines()[-1]); print('$2', j['value'], j['ms_per_step'], j['roofline']['conv_ms_per_step'], j['roofline']['frac'], j['clocks']['sm_mhz'])"; }
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_${tag}.json 2> gpurun_out/bench_${tag}.err; echo "bench exit $?"; line gpurun_out/bench_${tag}.json "xstore, stages by size"
DLV_DECONV_STAGES=8 timeout 300 python bench.py --steps 2 --warmup 2 --no-cpu-baseline > gpurun_out/bench_s8_${tag}.json 2> gpurun_out/bench_s8_${tag}.err; echo "bench exit $?"; line gpurun_out/bench_s8_${tag}.json "xstore, 8 stages everywhere"
DLV_DECONV_XSTORE=0 timeout 300 python bench.py --steps 2 --warmup 2 --no-cpu-baseline > gpurun_out/bench_x0_${tag}.json 2> gpurun_out/bench_x0_${tag}.err; echo "bench exit $?"; line gpurun_out/bench_x0_${tag}.json "half-sector stores (previous)"
K='regex:conv_tc|conv_is|is_reduce|norm_mish|final_blend|gather_windows|window_active|average_|cell_table|erode_|ccl_|scan_|bbox_init|relabel|boundary|paint_|edt_'
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" --launch-skip 60 -c 120 --csv --log-file gpurun_out/launches_cfg2_${tag}.csv python bench.py --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/ncu_list_cfg2_${tag}.log 2>&1; echo "ncu list exit $?"
python tools/ncu_summary.py launches gpurun_out/launches_cfg2_${tag}.csv > gpurun_out/launches_cfg2_${tag}.txt 2>&1; head -14 gpurun_out/launches_cfg2_${tag}.txt
DLV_DECONV_STAGES=8 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:conv_tc --launch-skip 14 -c 14 --csv --log-file gpurun_out/launches_s8_${tag}.csv python bench.py --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/ncu_list_s8_${tag}.log 2>&1; echo "ncu list (8 stages) exit $?"
grep -c conv_tc gpurun_out/launches_s8_${tag}.csv
