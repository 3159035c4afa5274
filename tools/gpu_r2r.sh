#!/bin/bash
mkdir -p gpurun_out
tag=${1:-r2r}
timeout 600 ncu --set full --clock-control none --import-source on -k 'regex:ccl_init|ccl_merge' -c 2 -o gpurun_out/prof_ccl_${tag} python bench.py --workload cfg3 --steps 1 --warmup 0 > gpurun_out/ncu_cclfull_${tag}.log 2>&1; echo "ncu exit $?"
ncu -i gpurun_out/prof_ccl_${tag}.ncu-rep --page details --csv > gpurun_out/${tag}_ccl_details.csv 2>&1
python - gpurun_out/${tag}_ccl_details.csv <<'PY'
import csv, sys
rows = list(csv.DictReader(open(sys.argv[1])))
want = ("Duration", "DRAM Throughput", "Memory Throughput", "Compute (SM) Throughput", "Achieved Occupancy", "Registers Per Thread", "Theoretical Occupancy", "Issue Slots Busy", "Executed Ipc Active", "L2 Hit Rate", "Warp Cycles Per Issued Instruction", "Eligible Warps Per Scheduler", "No Eligible", "Avg. Active Threads Per Warp", "Mem Busy", "Max Bandwidth", "L1/TEX Hit Rate")
for r in rows:
    if r.get("Metric Name") in want:
        print(r["Kernel Name"][:24], "|", r["Section Name"][:28], "|", r["Metric Name"], "=", r["Metric Value"], r["Metric Unit"])
PY
for s in 0 1; do python tools/ncu_stalls.py gpurun_out/prof_ccl_${tag}.ncu-rep $s 24 | awk 'NR<=2 || NR%2==1' > gpurun_out/${tag}_ccl_stalls_$s.txt 2>&1; head -16 gpurun_out/${tag}_ccl_stalls_$s.txt | cut -c1-210; done
