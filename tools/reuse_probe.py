#!/usr/bin/env python
"""Per-step view of the list-reuse cycle on bench.py's workload (one GPU): phase times of every step, whether it was a
rebuild or a reuse evaluation, and how many particles the certificate refused.

    SPHB_REUSE_PERIOD=10 python tools/reuse_probe.py [--workload c5] [--steps 24] [--precision 64]
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402

import bench as B  # noqa: E402
from sphugo_b200 import _lib as L  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="c5")
    ap.add_argument("--steps", type=int, default=24)
    ap.add_argument("--precision", type=int, default=64)
    a = ap.parse_args()
    pass  # (SPHB_REUSE_PERIOD fixes the cycle length; default: the library's own schedule)
    nx, ny, box, phys, desc = B.workload(a.workload, 1)
    pos = B.make_ic_c4(0, 1)[0] if a.workload == "c4" else B.make_ic(nx, ny, box, 0, 1)[0]
    n = len(pos)
    g = L.Handle(L.make_params(hor=(0.0, box[0]), ver=(0.0, box[1]), precision=a.precision, **phys), pos, None, np.full(n, 0.01))
    g.step(1)
    g.sync()
    c0 = g.counters()
    print(desc)
    print("step kind   total   keys   sort  reorder  knn   force   refused  refused/n")
    for k in range(a.steps):
        g.step(1)
        g.sync()
        pt, c1 = g.phase_times(), g.counters()
        kind = "reuse" if c1["reuse_steps"] > c0["reuse_steps"] else "build"
        ref = c1["knn_fallback"] - c0["knn_fallback"]
        print(f"{k + 1:4d} {kind} {pt['total']:7.3f} {pt['keys']:6.3f} {pt['sort']:6.3f} {pt['reorder']:6.3f} {pt['knn']:6.3f} "
              f"{pt['force']:6.3f} {ref:9d} {ref / n:9.2e}")
        c0 = c1
    g.close()


if __name__ == "__main__":
    main()
