#!/usr/bin/env python
"""ncu CSV (dram__bytes_read.sum, dram__bytes_write.sum, gpu__time_duration.sum of the conv launches of ONE window
batch) -> a text summary on stdout and, with --json, the `cfg2` entry of profiles/conv_traffic.json that bench.py reads
for `roofline.traffic`.

usage: python tools/conv_traffic_summary.py gpurun_out/conv_traffic_TAG.csv BATCH_WINDOWS [--json profiles/conv_traffic.json --source profiles/NAME.csv]
"""
import argparse
import collections
import csv
import json
import re

UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("csv")
    ap.add_argument("batch_windows", type=int)
    ap.add_argument("--json")
    ap.add_argument("--source")
    a = ap.parse_args()
    rows = list(csv.DictReader(l for l in open(a.csv) if l.startswith('"')))
    per = collections.OrderedDict()
    for r in rows:
        d = per.setdefault(int(r["ID"]), {"name": re.sub(r"\(.*", "", r["Kernel Name"]).replace("void dlv::", "").replace("void ", "").strip(),
                                          "grid": r.get("Grid Size", "")})
        d[r["Metric Name"]] = float(r["Metric Value"].replace(",", "")) * UNIT.get(r["Metric Unit"], 1.0)
    tot_b = tot_us = 0.0
    for i, d in per.items():
        b = d.get("dram__bytes_read.sum", 0.0) + d.get("dram__bytes_write.sum", 0.0)
        us = d.get("gpu__time_duration.sum", 0.0)
        tot_b += b
        tot_us += us
        print(f"{i:3d} {d['name']:34s} grid {d['grid']:14s} {us:9.1f} us {b / 1e6:10.1f} MB {b / us / 1e6 if us else 0:6.2f} TB/s")
    flop = a.batch_windows * 589824 * 285104
    print(f"total {tot_us:.1f} us, {tot_b / 1e6:.1f} MB -> {tot_b / a.batch_windows / 1e6:.1f} MB per window; algorithmic FLOP per batch "
          f"{a.batch_windows} x 589824 x 285104 = {flop:.3e} -> {flop / tot_us / 1e6:.0f} TFLOP/s under ncu ({len(per)} launches)")
    if a.json:
        j = json.load(open(a.json))
        j["cfg2"] = {"dram_bytes_per_window": tot_b / a.batch_windows, "batch_windows": a.batch_windows,
                     "conv_launches_per_batch": len(per), "ncu_time_us_per_batch": tot_us,
                     "source": f"{a.source or a.csv} (ncu dram__bytes_read.sum + dram__bytes_write.sum of the {len(per)} conv/deconv launches of one "
                               f"full {a.batch_windows}-window batch of cfg2, --clock-control none)"}
        json.dump(j, open(a.json, "w"), indent=1)
        print("updated", a.json)


if __name__ == "__main__":
    main()
