#!/bin/bash
# bench variants: $1.. = env assignments to try, e.g. "DLV_IS_T=2" "DLV_IS_T=4"
mkdir -p gpurun_out
for v in "$@"; do
  tag=$(echo "$v" | tr ' =' '__')
  echo "=== $v"
  env $v timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err
  python - <<PY
import json
d=json.load(open("gpurun_out/bench_$tag.json"))
r=d["roofline"]
print("$v", "value", round(d["value"],4), "ms", round(d["ms_per_step"],1), "e2e", round(d["e2e"]["value"],4), "conv_ms", round(r["conv_ms_per_step"],1), "TF", round(r["achieved"],1), "unet_ms", round(r["unet_ms_per_step"],1), "clk", d["clocks"]["sm_mhz"], d["clocks"]["reasons"])
PY
done
