#!/usr/bin/env python
"""Dynamic opcode histogram of one kernel from an ncu report's source page (needs -lineinfo + --import-source on):
  ncu -i rep.ncu-rep --page source --csv --kernel-name K --launch-count 1 > k.csv ;  python tools/ncu_hot.py k.csv"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
h = [i for i, r in enumerate(rows) if "Instructions Executed" in r][0]
hdr = rows[h]
ia, isrc, iss = hdr.index("Instructions Executed"), hdr.index("Source"), hdr.index("Warp Stall Sampling (All Samples)")
data = []
for r in rows[h + 1:]:
    if len(r) > ia and r[ia].isdigit():
        data.append((int(r[ia]), int(r[iss]), r[isrc].strip()))
tot = sum(d[0] for d in data)
tots = sum(d[1] for d in data)
print("total warp instructions", tot, "static", len(data), "stall samples", tots)
op, st = collections.Counter(), collections.Counter()
for e, s, src in data:
    t = src.split()
    o = t[1] if t[0].startswith("@") else t[0]
    o = o.split(".")[0].rstrip(";")
    op[o] += e
    st[o] += s
for o, e in op.most_common(int(sys.argv[2]) if len(sys.argv) > 2 else 30):
    print("%-12s %5.1f%% of instructions   %5.1f%% of stall samples" % (o, 100 * e / tot, 100 * st[o] / max(tots, 1)))
