#!/usr/bin/env python
"""Top stalled SASS instructions of one kernel launch from an .ncu-rep (needs -lineinfo + --import-source on).
usage: ncu_stalls.py rep [launch_skip] [topN]"""
import csv, io, subprocess, sys
rep = sys.argv[1]; skip = sys.argv[2] if len(sys.argv) > 2 else "0"; topn = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--launch-skip", skip, "--launch-count", "1"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = next(r for r in rows if "Address" in r and "# Samples" in r)
idx = {h: i for i, h in enumerate(hdr)}
data = [r for r in rows if len(r) == len(hdr) and r[idx["# Samples"]].isdigit()]
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
tot = sum(int(r[idx["# Samples"]]) for r in data)
print(rows[0][1] if len(rows[0]) > 1 else "", "total samples", tot, "instructions", len(data))
agg = {h: sum(int(r[idx[h]]) for r in data) for h in stalls}
print("stall mix:", {k: v for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]})
for r in sorted(data, key=lambda r: -int(r[idx["# Samples"]]))[:topn]:
    st = sorted(((h, int(r[idx[h]])) for h in stalls if int(r[idx[h]]) > 0), key=lambda kv: -kv[1])[:3]
    print(r[idx["# Samples"]].rjust(7), r[idx["Address"]][-5:], r[idx["Source"]].strip()[:78].ljust(78), st)
