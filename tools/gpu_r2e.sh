#!/bin/bash
# round 2, call e: experiments that decide the next step of the fused conv (all timing-only unless stated).
#   1. umma_bench4: SS-mode MMA rate, cta_group::1 vs cta_group::2, with background shared-memory traffic
#   2. DLV_IS_MODE bit field on the production library (skip TMEM loads / zeroing / transform / output stores)
#   3. transform-role variants: copy only (xfcopy), arithmetic only (xfmath)
#   4. DLV_IS_TX=4 (four-tile columns for the 32 -> 32 layers that normalise while staging): parity tests + cfg2 bench
mkdir -p gpurun_out
tag=${1:-r2e}
timeout 120 tools/ubench/umma_bench4 > gpurun_out/${tag}_umma4.txt 2>&1; echo "umma4 exit $?"; cat gpurun_out/${tag}_umma4.txt
isdbg() {   # label, env...
  local label=$1; shift
  env "$@" DLV_IS_DEBUG=1 timeout 300 python bench.py --workload small --steps 1 --warmup 1 --no-cpu-baseline 2> gpurun_out/${tag}_isdbg_${label}.txt > /dev/null
  echo "=== $label ($*) exit $?"; grep "^\[is\]" gpurun_out/${tag}_isdbg_${label}.txt | head -8 | cut -c1-60,88-
}
isdbg prod DLV_X=0
for m in 1 2 3 4 8 15; do isdbg mode$m DLV_IS_MODE=$m; done
isdbg xfcopy DLV_LIB=$PWD/delivr_cfos_b200/libdelivr_b200_xfcopy.so
isdbg xfmath DLV_LIB=$PWD/delivr_cfos_b200/libdelivr_b200_xfmath.so
isdbg tx4 DLV_IS_TX=4
isdbg t4 DLV_IS_T=4
DLV_IS_TX=4 timeout 600 python -m pytest tests/test_gpu_a_conv.py tests/test_gpu_b_unet.py -q -m gpu -x -p no:cacheprovider > gpurun_out/${tag}_tx4_tests.log 2>&1; echo "tx4 tests exit $?"; tail -n 2 gpurun_out/${tag}_tx4_tests.log
for v in "DLV_X=0" "DLV_IS_TX=4"; do
  env $v timeout 600 python bench.py --steps 2 --warmup 2 --no-cpu-baseline > gpurun_out/${tag}_bench_${v%%=*}.json 2> gpurun_out/${tag}_bench_${v%%=*}.err || { echo "bench $v FAILED"; tail -3 gpurun_out/${tag}_bench_${v%%=*}.err; continue; }
  python - "$v" gpurun_out/${tag}_bench_${v%%=*}.json <<'PY'
import json, sys
d = json.load(open(sys.argv[2])); r = d["roofline"]
print(sys.argv[1], "value", round(d["value"], 4), "ms", round(d["ms_per_step"], 1), "e2e", round(d["e2e"]["value"], 4),
      "conv_ms", round(r["conv_ms_per_step"], 1), "TF", round(r["achieved"], 1), "unet_ms", round(r["unet_ms_per_step"], 1),
      "clk", d["clocks"]["sm_mhz"], d["clocks"]["reasons"])
PY
done
