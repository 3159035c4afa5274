#!/bin/bash
# round 2, call g: where does the finalise stage of the resident-volume step go (DLV_TRACE), sub-step A/B, bench
mkdir -p gpurun_out
tag=${1:-r2g}
DLV_TRACE=1 timeout 300 python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/${tag}_trace.json 2> gpurun_out/${tag}_trace.txt; echo "trace exit $?"
grep "dlv_segment" gpurun_out/${tag}_trace.txt | head -60
for v in "DLV_IS_NSUB=1" "DLV_IS_NSUB=2"; do
  env $v DLV_IS_DEBUG=1 timeout 300 python bench.py --workload small --steps 1 --warmup 1 --no-cpu-baseline 2> gpurun_out/${tag}_isdbg_${v#*=}.txt > /dev/null
  echo "=== $v"; grep "^\[is\]" gpurun_out/${tag}_isdbg_${v#*=}.txt | head -8 | cut -c1-60,88-
done
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; echo "bench exit $?"
python - gpurun_out/${tag}_bench.json <<'PY'
import json, sys
d = json.load(open(sys.argv[1])); r = d["roofline"]
print("value", round(d["value"], 4), "ms", round(d["ms_per_step"], 1), "e2e", round(d["e2e"]["value"], 4),
      "conv_ms", round(r["conv_ms_per_step"], 1), "TF", round(r["achieved"], 1), "unet_ms", round(r["unet_ms_per_step"], 1), "fin", round(r["finalise_ms_per_step"], 1),
      "clk", d["clocks"]["sm_mhz"], d["clocks"]["reasons"])
PY
