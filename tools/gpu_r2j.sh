#!/bin/bash
# round 2, call j: ncu --set full of the fused conv kernels on cfg2 (window batch 32 to keep the replay's memory save/restore
# short): launches 8,9 = first layer + 32->32 (four-tile), launches 14,15 = 64->32 + 32->32 of the second batch
mkdir -p gpurun_out
tag=${1:-r2j}
timeout 700 ncu --set full --clock-control none --import-source on -k regex:conv_is --launch-skip 8 -c 2 -o gpurun_out/prof_is_a_${tag} python bench.py --steps 1 --warmup 0 --no-cpu-baseline --window-batch 32 > gpurun_out/ncu_is_a_${tag}.log 2>&1; echo "ncu a exit $?"
timeout 700 ncu --set full --clock-control none --import-source on -k regex:conv_is --launch-skip 14 -c 2 -o gpurun_out/prof_is_b_${tag} python bench.py --steps 1 --warmup 0 --no-cpu-baseline --window-batch 32 > gpurun_out/ncu_is_b_${tag}.log 2>&1; echo "ncu b exit $?"
ls -la gpurun_out/prof_is_*_${tag}.ncu-rep
for r in a b; do
  python tools/ncu_summary.py report gpurun_out/prof_is_${r}_${tag}.ncu-rep > gpurun_out/${tag}_metrics_${r}.txt 2>&1
  for s in 0 1; do python tools/ncu_stalls.py gpurun_out/prof_is_${r}_${tag}.ncu-rep $s 60 > gpurun_out/${tag}_stalls_${r}${s}.txt 2>&1; done
done
head -50 gpurun_out/${tag}_metrics_a.txt
