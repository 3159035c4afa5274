#!/bin/bash
# first layer: epilogue sets (2 = library default, 1 = ne1 variant) x tiles per column (DLV_IS_TF 4 / 2)
mkdir -p gpurun_out
tag=${1:-r2m}
for lib in main ne1; do for tf in 4 2; do
  L=$PWD/delivr_cfos_b200/libdelivr_b200.so; [ $lib = ne1 ] && L=$PWD/delivr_cfos_b200/libdelivr_b200_ne1.so
  DLV_LIB=$L DLV_IS_TF=$tf DLV_IS_DEBUG=1 timeout 300 python bench.py --workload small --steps 1 --warmup 1 --no-cpu-baseline 2> gpurun_out/${tag}_isdbg_${lib}_tf$tf.txt > /dev/null
  echo "=== $lib TF=$tf"; grep "^\[is\]" gpurun_out/${tag}_isdbg_${lib}_tf$tf.txt | head -1 | cut -c1-60,88-
done; done
timeout 600 python -m pytest tests/test_gpu_b_unet.py tests/test_gpu_e_segment.py -q -m gpu -x -p no:cacheprovider > gpurun_out/${tag}_tests.log 2>&1; echo "tests exit $?"; tail -n 2 gpurun_out/${tag}_tests.log
timeout 600 python bench.py --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; echo "bench exit $?"
python - gpurun_out/${tag}_bench.json <<'PY'
import json, sys
d = json.load(open(sys.argv[1])); r = d["roofline"]
print("value", round(d["value"], 4), "ms", round(d["ms_per_step"], 1), "e2e", round(d["e2e"]["value"], 4),
      "conv_ms", round(r["conv_ms_per_step"], 1), "TF", round(r["achieved"], 1), "unet_ms", round(r["unet_ms_per_step"], 1), "fin", round(r["finalise_ms_per_step"], 1),
      "gauss", d["config"].get("blend_gaussian", {}).get("value"), "clk", d["clocks"]["sm_mhz"], d["clocks"]["reasons"])
PY
