#!/bin/bash
# A/B run of variant builds of the library (make -C delivr_cfos_b200/csrc VARIANT=x DEFS=...; selected with DLV_LIB):
# per variant the conv parity tests, the per-layer cycle counters of the fused conv (DLV_IS_DEBUG) on the small
# workload and a short cfg2 bench.   usage (under gpurun): bash tools/gpu_variants.sh tag variant [variant...]
# ("main" = the default library)
mkdir -p gpurun_out
tag=${1:-var}; shift
for v in "$@"; do
  lib=$PWD/delivr_cfos_b200/libdelivr_b200_$v.so
  [ "$v" = "main" ] && lib=$PWD/delivr_cfos_b200/libdelivr_b200.so
  echo "=== variant $v ($lib)"
  DLV_LIB=$lib timeout 600 python -m pytest tests/test_gpu_a_conv.py tests/test_gpu_b_unet.py -q -m gpu -x -p no:cacheprovider > gpurun_out/var_${tag}_${v}_tests.log 2>&1
  echo "tests exit $?"; tail -n 2 gpurun_out/var_${tag}_${v}_tests.log
  DLV_LIB=$lib DLV_IS_DEBUG=1 timeout 300 python bench.py --workload small --steps 1 --warmup 1 --no-cpu-baseline 2> gpurun_out/var_${tag}_${v}_isdebug.txt > /dev/null
  grep "^\[is\]" gpurun_out/var_${tag}_${v}_isdebug.txt | head -8 | cut -c1-60,88-
  DLV_LIB=$lib timeout 600 python bench.py --steps 2 --warmup 2 --no-cpu-baseline > gpurun_out/var_${tag}_${v}_bench.json 2> gpurun_out/var_${tag}_${v}_bench.err || { echo "bench FAILED"; tail -3 gpurun_out/var_${tag}_${v}_bench.err; continue; }
  python - "$v" gpurun_out/var_${tag}_${v}_bench.json <<'PY'
import json, sys
d = json.load(open(sys.argv[2])); r = d["roofline"]
print(sys.argv[1], "value", round(d["value"], 4), "ms", round(d["ms_per_step"], 1), "e2e", round(d["e2e"]["value"], 4),
      "conv_ms", round(r["conv_ms_per_step"], 1), "TF", round(r["achieved"], 1), "unet_ms", round(r["unet_ms_per_step"], 1),
      "clk", d["clocks"]["sm_mhz"], d["clocks"]["reasons"])
PY
done
