#!/bin/bash
# development GPU call: parity tests of the touched kernels, traced + debug-counter runs, bench, ncu launch list
# usage: gpu_dev.sh tag [test files...]
mkdir -p gpurun_out
tag=${1:-dev}; shift
files="$@"
[ -z "$files" ] && files="tests/test_gpu_a_conv.py tests/test_gpu_b_unet.py tests/test_gpu_e_segment.py"
bash tools/gpu_ci.sh $files > gpurun_out/ci_${tag}.log 2>&1; echo "ci exit $?"; grep -E "passed|failed|error" gpurun_out/ci_${tag}.log | tail -8
DLV_IS_DEBUG=1 timeout 600 python bench.py --workload small --steps 1 --warmup 1 --no-cpu-baseline 2> gpurun_out/isdebug_${tag}.txt > /dev/null; echo "isdebug exit $?"; grep "^\[is\]" gpurun_out/isdebug_${tag}.txt | head -8
DLV_TRACE=1 timeout 900 python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/bench_${tag}.json 2> gpurun_out/bench_${tag}.err; echo "bench exit $?"
python - <<PY
import json
d=json.load(open("gpurun_out/bench_${tag}.json"))
r=d["roofline"]
print("value", round(d["value"],4), "ms", round(d["ms_per_step"],1), "e2e", round(d["e2e"]["value"],4), "conv_ms", round(r["conv_ms_per_step"],1), "TF", round(r["achieved"],1), "unet_ms", round(r["unet_ms_per_step"],1), "fin", round(r["finalise_ms_per_step"],1), "clk", d["clocks"]["sm_mhz"], d["clocks"]["reasons"])
PY
grep "dlv_segment" gpurun_out/bench_${tag}.err | tail -9
K='regex:conv_tc|conv_is|is_reduce|norm_mish|final_blend|gather_windows|window_active|average_kernel|erode_|ccl_|scan_|bbox_init|relabel|boundary'
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" -c 400 --csv --log-file gpurun_out/launches_${tag}.csv python bench.py --workload small --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_list_${tag}.log 2>&1; echo "ncu list exit $?"
