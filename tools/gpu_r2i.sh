#!/bin/bash
# round 2, call i: first layer with four-tile columns (DLV_IS_TF=4 default) against two-tile
mkdir -p gpurun_out
tag=${1:-r2i}
timeout 600 python -m pytest tests/test_gpu_b_unet.py tests/test_gpu_e_segment.py tests/test_gpu_i_mirrors.py -q -m gpu -x -p no:cacheprovider > gpurun_out/${tag}_tests.log 2>&1; echo "tests exit $?"; tail -n 2 gpurun_out/${tag}_tests.log
for v in "DLV_IS_TF=4" "DLV_IS_TF=2"; do
  env $v DLV_IS_DEBUG=1 timeout 300 python bench.py --workload small --steps 1 --warmup 1 --no-cpu-baseline 2> gpurun_out/${tag}_isdbg_${v#*=}.txt > /dev/null
  echo "=== $v"; grep "^\[is\]" gpurun_out/${tag}_isdbg_${v#*=}.txt | head -2 | cut -c1-60,88-
  env $v timeout 600 python bench.py --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/${tag}_bench_${v#*=}.json 2> gpurun_out/${tag}_bench_${v#*=}.err; echo "bench exit $?"
  python - gpurun_out/${tag}_bench_${v#*=}.json <<'PY'
import json, sys
d = json.load(open(sys.argv[1])); r = d["roofline"]
print("value", round(d["value"], 4), "ms", round(d["ms_per_step"], 1), "e2e", round(d["e2e"]["value"], 4),
      "conv_ms", round(r["conv_ms_per_step"], 1), "TF", round(r["achieved"], 1), "unet_ms", round(r["unet_ms_per_step"], 1), "fin", round(r["finalise_ms_per_step"], 1),
      "clk", d["clocks"]["sm_mhz"], d["clocks"]["reasons"])
PY
done
