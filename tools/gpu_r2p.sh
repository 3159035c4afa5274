#!/bin/bash
# CC changes: parity tests that label, cfg3 bench, CC kernel times + traffic
mkdir -p gpurun_out
tag=${1:-r2p}
timeout 600 python -m pytest tests/test_gpu_c_ccl.py tests/test_gpu_f_slabs.py tests/test_gpu_i_mirrors.py tests/test_gpu_h_paint.py -q -m gpu -x -p no:cacheprovider > gpurun_out/${tag}_tests.log 2>&1; echo "tests exit $?"; tail -n 2 gpurun_out/${tag}_tests.log
timeout 600 python bench.py --workload cfg3 --steps 5 --warmup 3 > gpurun_out/${tag}_bench_cfg3.json 2> gpurun_out/${tag}_bench_cfg3.err; echo "cfg3 exit $?"; cat gpurun_out/${tag}_bench_cfg3.json
timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k 'regex:ccl_|scan_|bbox_' -c 9 --csv --log-file gpurun_out/ccl_traffic_${tag}.csv python bench.py --workload cfg3 --steps 1 --warmup 0 > gpurun_out/ncu_ccl_${tag}.log 2>&1; echo "ncu ccl exit $?"
python - gpurun_out/ccl_traffic_${tag}.csv <<'PY'
import csv, sys, re, collections
rows = [r for r in csv.DictReader([l for l in open(sys.argv[1]) if not l.startswith("==")])]
agg = collections.OrderedDict()
for r in rows:
    k = (r["ID"], re.sub(r"\(.*", "", r["Kernel Name"]).replace("void dlv::", ""))
    agg.setdefault(k, {})[r["Metric Name"]] = (float(r["Metric Value"].replace(",", "")), r["Metric Unit"])
for (i, k), m in agg.items():
    print(i, k, {a: f"{v[0]:.3f} {v[1]}" for a, v in m.items()})
PY
