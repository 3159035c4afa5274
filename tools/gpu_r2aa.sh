#!/bin/bash
# Round-2 call aa (1 GPU): painter resolve with two groups per round / four rounds per pass, deconv pipeline depth by output
# size; parity, cfg3 + cfg2 lines, window-batch A/B, ncu --set full of the large deconv and of the sparse CC kernels.
mkdir -p gpurun_out
tag=${1:-r2aa}
timeout 600 python -m pytest tests/ -q -m gpu -p no:cacheprovider > gpurun_out/pytest_${tag}.log 2>&1; t=$?; echo "pytest exit $t"; tail -12 gpurun_out/pytest_${tag}.log
timeout 300 python bench.py --workload cfg3 --steps 3 --warmup 2 > gpurun_out/bench_cfg3_${tag}.json 2> gpurun_out/bench_cfg3_${tag}.err; echo "cfg3 exit $?"; cat gpurun_out/bench_cfg3_${tag}.json
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_${tag}.json 2> gpurun_out/bench_${tag}.err; echo "bench exit $?"; cat gpurun_out/bench_${tag}.json; tail -3 gpurun_out/bench_${tag}.err
timeout 300 python bench.py --steps 2 --warmup 2 --no-cpu-baseline --window-batch 192 > gpurun_out/bench_wb192_${tag}.json 2> gpurun_out/bench_wb192_${tag}.err; echo "bench wb192 exit $?"
python -c "import json,sys; j=json.loads(open('gpurun_out/bench_wb192_${tag}.json').read().strip().splitlines()[-1]); print('wb192', j['value'], j['ms_per_step'], j['roofline']['conv_ms_per_step'], j['roofline']['frac'])"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:conv_tc --launch-skip 13 -c 1 -o gpurun_out/prof_deconv_${tag} python bench.py --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/ncu_deconv_${tag}.log 2>&1; echo "ncu deconv exit $?"
python tools/ncu_summary.py report gpurun_out/prof_deconv_${tag}.ncu-rep > gpurun_out/${tag}_deconv_metrics.txt 2>&1; head -5 gpurun_out/${tag}_deconv_metrics.txt
ncu -i gpurun_out/prof_deconv_${tag}.ncu-rep --page details --csv > gpurun_out/${tag}_deconv_details.csv 2>&1
DLV_BENCH_CFG3_MIN_MS=0 timeout 400 ncu --set full --clock-control none --import-source on -k 'regex:ccl_merge|ccl_compress|ccl_relabel|paint_resolve' --launch-skip 6 -c 4 -o gpurun_out/prof_ccl_${tag} python bench.py --workload cfg3 --steps 1 --warmup 0 > gpurun_out/ncu_cclfull_${tag}.log 2>&1; echo "ncu ccl exit $?"
ncu -i gpurun_out/prof_ccl_${tag}.ncu-rep --page details --csv > gpurun_out/${tag}_ccl_details.csv 2>&1
python - gpurun_out/${tag}_ccl_details.csv gpurun_out/${tag}_deconv_details.csv <<'PY'
import csv, sys
want = ("Duration", "DRAM Throughput", "Memory Throughput", "Compute (SM) Throughput", "Achieved Occupancy", "Registers Per Thread", "Theoretical Occupancy", "Issue Slots Busy", "Executed Ipc Active", "L2 Hit Rate", "Warp Cycles Per Issued Instruction", "Eligible Warps Per Scheduler", "No Eligible", "Avg. Active Threads Per Warp", "Mem Busy", "Max Bandwidth", "L1/TEX Hit Rate")
for f in sys.argv[1:]:
    try:
        rows = list(csv.DictReader(l for l in open(f) if l.startswith('"')))
    except Exception as e:
        print(f, e); continue
    for r in rows:
        if r.get("Metric Name") in want:
            print(r["Kernel Name"][:28], "|", r["Section Name"][:28], "|", r["Metric Name"], "=", r["Metric Value"], r["Metric Unit"])
PY
