#!/usr/bin/env python
"""Per-CUDA-source-line cost of one kernel from an `ncu --set full --import-source on` capture (needs -lineinfo).

  python tools/ncu_lines.py rep.ncu-rep <kernel regex> [top N] [launch index]

Prints, for the first matching launch, the source lines ordered by stall samples with their share of the
warp-instructions executed.  Read here, no GPU needed.
"""
import csv
import subprocess
import sys

rep, pat = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
which = int(sys.argv[4]) if len(sys.argv) > 4 else 0
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass",
                      "--kernel-name", "regex:" + pat], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
# the report repeats [File Path / header / lines...] blocks per file and per launch; a launch starts with "Kernel Name"
launches, cur = [], None
fname = None
for r in rows:
    if not r:
        continue
    if r[0] == "Kernel Name":
        cur = []
        launches.append((r[1], cur))
        continue
    if r[0] == "File Path":
        fname = r[1].split("/")[-1]
        if cur is None:
            cur = []
            launches.append(("?", cur))
        continue
    if r[0] == "Line No":
        hdr = r
        continue
    if r[0] == "" or cur is None:
        continue  # SASS rows
    try:
        ln = int(r[0])
    except ValueError:
        continue
    def num(v):
        try:
            return int(v.replace(",", ""))
        except ValueError:
            return 0
    smp = num(r[hdr.index("# Samples")])
    ins = num(r[hdr.index("Instructions Executed")])
    cur.append((fname, ln, r[1].strip(), smp, ins))
name, data = launches[which]
agg = {}
for f, ln, src, smp, ins in data:  # launches of the same kernel repeat the block: sum them (shares are unchanged)
    a = agg.setdefault((f, ln), [f, ln, src, 0, 0])
    a[3] += smp
    a[4] += ins
data = [tuple(a) for a in agg.values()]
ts = max(1, sum(d[3] for d in data))
ti = max(1, sum(d[4] for d in data))
print(f"{name[:70]}  lines={len(data)} samples={ts} warp-inst={ti}")
print(f"{'file:line':28s} {'smp%':>6s} {'inst%':>6s}  source")
for f, ln, src, smp, ins in sorted(data, key=lambda d: -d[3])[:top]:
    print(f"{(f + ':' + str(ln))[:28]:28s} {100 * smp / ts:6.2f} {100 * ins / ti:6.2f}  {src[:100]}")
