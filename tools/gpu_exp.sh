#!/bin/bash
# timing experiments on the fused conv: DLV_IS_MODE bit field (results invalid, timing only)
mkdir -p gpurun_out
tag=${1:-exp}; shift
for m in "$@"; do
  DLV_IS_MODE=$m DLV_IS_DEBUG=1 timeout 600 python bench.py --workload small --steps 1 --warmup 1 --no-cpu-baseline 2> gpurun_out/ismode_${tag}_$m.txt > /dev/null
  echo "=== DLV_IS_MODE=$m exit $?"; grep "^\[is\]" gpurun_out/ismode_${tag}_$m.txt | head -8 | cut -c1-60,88-
done
