#!/bin/bash
# Runs every GPU parity test file in its own process (a device trap must not poison the next file).
# Usage under gpurun:  bash tools/gpu_ci.sh [files...]
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
files="$@"
[ -z "$files" ] && files=$(ls tests/test_gpu_*.py)
rc=0
for f in $files; do
  b=$(basename $f .py)
  echo "=== $f" 
  timeout 900 python -m pytest $f -q -m gpu -s -p no:cacheprovider > gpurun_out/$b.log 2>&1
  r=$?
  echo "exit $r"; tail -n 25 gpurun_out/$b.log
  [ $r -ne 0 ] && rc=1
done
exit $rc
