#!/usr/bin/env python
"""Instruction histogram per kernel of the in-tree library (cuobjdump -sass): the mnemonics that prove the tcgen05 / TMEM /
bulk-TMA path (UTCHMMA, UTCBAR, LDTM, STTM, UBLKCP) next to the usual suspects.  usage: python tools/sass_histogram.py > profiles/NAME.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "delivr_cfos_b200", "libdelivr_b200.so")
COLS = "UTCHMMA UTCBAR LDTM STTM UBLKCP UTMALDG UTCCP SYNCS MUFU FFMA2 FADD2 FMUL2 ATOMG RED REDG LDS STS LDG STG SHFL".split()


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    names = subprocess.run(["cu++filt"], input="\n".join(re.findall(r"Function : (\S+)", sass)), capture_output=True, text=True).stdout.split("\n")
    demangled = dict(zip(re.findall(r"Function : (\S+)", sass), names))
    hist = collections.OrderedDict()
    cur = None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = hist.setdefault(m.group(1), collections.Counter())
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", line)
        if m and cur is not None:
            cur[m.group(1)] += 1
            cur["__total__"] += 1
    print(f"SASS instruction histogram of delivr_cfos_b200/libdelivr_b200.so (cuobjdump -sass, sm_100a), {sys.argv[1] if len(sys.argv) > 1 else 'HEAD'}")
    print("columns: " + " ".join(COLS))
    print()
    tot = collections.Counter()
    rows = []
    for k, c in hist.items():
        name = re.sub(r"\(.*", "", demangled.get(k, k).replace("(int)", "").replace("(bool)", "")).replace("void dlv::", "").strip()
        rows.append((name, c))
        tot.update(c)
    for name, c in sorted(rows):
        print(f"{name[:60]:60s} " + " ".join(f"{c[x]:5d}" for x in COLS) + f"   total {c['__total__']}")
    print()
    print(f"{'library':60s} " + " ".join(f"{tot[x]:5d}" for x in COLS) + f"   total {tot['__total__']}")


if __name__ == "__main__":
    main()
