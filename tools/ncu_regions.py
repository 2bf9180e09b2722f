#!/usr/bin/env python
"""Cost of one kernel by source-line regions (from tools/ncu_lines.py output on stdin).

  python tools/ncu_lines.py rep.ncu-rep <kernel> 1000 | python tools/ncu_regions.py name:lo-hi name:lo-hi ...
Lines of other files (inlined headers) and helper lines outside every range go to "other"."""
import sys
regs = []
for a in sys.argv[1:]:
    n, r = a.split(":")
    lo, hi = r.split("-")
    regs.append((n, int(lo), int(hi)))
acc = {n: [0.0, 0.0] for n, _, _ in regs}
acc["other"] = [0.0, 0.0]
for ln in sys.stdin.read().splitlines()[2:]:
    f = ln.split()
    if len(f) < 3 or ":" not in f[0]:
        continue
    fn, l = f[0].rsplit(":", 1)
    try:
        l, smp, ins = int(l), float(f[1]), float(f[2])
    except ValueError:
        continue
    key = "other"
    if fn.endswith(".cuh") or fn.endswith(".cu"):
        for n, lo, hi in regs:
            if lo <= l <= hi:
                key = n
                break
    acc[key][0] += smp
    acc[key][1] += ins
print(f"{'region':28s} {'smp%':>7s} {'inst%':>7s}")
for n in [r[0] for r in regs] + ["other"]:
    print(f"{n:28s} {acc[n][0]:7.2f} {acc[n][1]:7.2f}")
