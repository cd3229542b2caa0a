#!/usr/bin/env python
"""Per-CUDA-source-line instruction / stall shares of one kernel from an ncu report:
  ncu -i rep.ncu-rep --page source --print-source cuda,sass --csv --kernel-name K --launch-count 1 > k.csv ; python tools/ncu_lines.py k.csv [N]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
agg = {}
hdr = None
for r in rows:
    if r and r[0] == "Line No":
        hdr = r
        ie, iss = hdr.index("Instructions Executed"), hdr.index("Warp Stall Sampling (All Samples)")
        continue
    if hdr and len(r) > ie and r[0].isdigit() and r[ie].isdigit() and r[iss].isdigit():
        a = agg.setdefault(int(r[0]), [0, 0, r[1][:120]])
        a[0] += int(r[ie])
        a[1] += int(r[iss])
tot = sum(a[0] for a in agg.values())
ts = sum(a[1] for a in agg.values())
print("total warp instructions", tot)
for ln, a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:int(sys.argv[2]) if len(sys.argv) > 2 else 40]:
    print("%5d %5.1f%% inst %5.1f%% stall | %s" % (ln, 100 * a[0] / tot, 100 * a[1] / max(ts, 1), a[2]))
