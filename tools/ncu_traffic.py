#!/usr/bin/env python
"""DRAM traffic per launch of the dominant kernels from an `ncu --set full` capture -> profiles/traffic.json
(read by bench.py for roofline.traffic).

  python tools/ncu_traffic.py rep.ncu-rep <workload> <precision>      e.g.  ... gpurun_out/full.ncu-rep c5 64
"""
import csv
import json
import os
import re
import subprocess
import sys

rep, wl, prec = sys.argv[1], sys.argv[2], sys.argv[3]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
ik, ir, iw = hdr.index("Kernel Name"), hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
acc = {}
for r in rows[2:]:
    name = re.sub(r"[<(].*", "", r[ik]).replace("void ", "")
    b = float(r[ir].replace(",", "")) * scale[units[ir]] + float(r[iw].replace(",", "")) * scale[units[iw]]
    acc.setdefault(name, []).append(b)
phase = {"k_knn_tile": "knn", "k_force_st": "force", "k_force_st32": "force", "k_reorder": "reorder"}
path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "traffic.json")
d = json.load(open(path)) if os.path.exists(path) else {}
for k, v in acc.items():
    if k in phase:
        d[f"{phase[k]}_{wl}_f{prec}"] = sum(v) / len(v)
        print(k, f"{sum(v) / len(v) / 1e9:.3f} GB per launch over {len(v)} launches")
json.dump(d, open(path, "w"), indent=1, sort_keys=True)
