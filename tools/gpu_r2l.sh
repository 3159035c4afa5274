#!/bin/bash
mkdir -p gpurun_out
tag=${1:-r2l}
DLV_IS_DEBUG=1 timeout 300 python bench.py --workload small --steps 1 --warmup 1 --no-cpu-baseline 2> gpurun_out/${tag}_isdbg.txt > /dev/null
grep "^\[is\]" gpurun_out/${tag}_isdbg.txt | head -8 | cut -c1-60,88-
DLV_IS_TF=2 DLV_IS_DEBUG=1 timeout 300 python bench.py --workload small --steps 1 --warmup 1 --no-cpu-baseline 2> gpurun_out/${tag}_isdbg_tf2.txt > /dev/null
grep "^\[is\]" gpurun_out/${tag}_isdbg_tf2.txt | head -1 | cut -c1-60,88-
