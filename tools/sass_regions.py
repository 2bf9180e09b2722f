#!/usr/bin/env python
"""Instruction-count breakdown of one kernel of an ncu --set full --import-source on capture, by SASS region.
   python tools/sass_regions.py rep.ncu-rep [kernel_index] [bucket] [tiles]"""
import csv, subprocess, sys
rep = sys.argv[1]; kidx = int(sys.argv[2]) if len(sys.argv) > 2 else 0
B = int(sys.argv[3]) if len(sys.argv) > 3 else 250
tiles = float(sys.argv[4]) if len(sys.argv) > 4 else 32768.0
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
secs = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"] + [len(rows)]
st, en = secs[kidx], secs[kidx + 1]
hdr = rows[st + 1]
ia, isrc, ismp = hdr.index("Instructions Executed"), hdr.index("Source"), hdr.index("# Samples")
data = [(int(r[ia] or 0), r[isrc], int(r[ismp] or 0)) for r in rows[st + 2:en] if len(r) > ia]
tot = sum(d[0] for d in data); stot = max(1, sum(d[2] for d in data))
print(rows[st][1][:60], "total inst", tot, "per tile", tot / tiles, "sass", len(data))
for b in range(0, len(data), B):
    seg = data[b:b + B]
    n = sum(d[0] for d in seg); sm = sum(d[2] for d in seg)
    ops = {}
    for d in seg:
        t = d[1].split()
        op = (t[1] if t[0].startswith("@") else t[0]).split(".")[0]
        ops[op] = ops.get(op, 0) + d[0]
    top = sorted(ops.items(), key=lambda kv: -kv[1])[:7]
    print(f"{b:5d} inst {n / tot * 100:5.1f}% smp {sm / stot * 100:5.1f}% per-tile {n / tiles:7.0f}", [(k, round(v / tiles)) for k, v in top])
if len(sys.argv) > 5:
    lo, hi = [int(x) for x in sys.argv[5].split(":")]
    for k in range(lo, hi):
        print(k, data[k][0], data[k][2], data[k][1][:110])
