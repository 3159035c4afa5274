#!/bin/bash
# usage: gpu_bench.sh "ENV=VAL ..." ...   one short cfg2 bench per environment setting (variant libraries via DLV_LIB)
for v in "$@"; do
  echo "=== $v"
  env $v timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bv.json 2> gpurun_out/bv.err || { echo "FAILED"; tail -3 gpurun_out/bv.err; continue; }
  python - "$v" <<'PY'
import json, sys
d = json.load(open("gpurun_out/bv.json")); r = d["roofline"]
print(sys.argv[1], "value", round(d["value"], 4), "ms", round(d["ms_per_step"], 1), "e2e", round(d["e2e"]["value"], 4),
      "conv_ms", round(r["conv_ms_per_step"], 1), "TF", round(r["achieved"], 1), "unet_ms", round(r["unet_ms_per_step"], 1),
      "clk", d["clocks"]["sm_mhz"], d["clocks"]["reasons"])
PY
done
