#!/usr/bin/env python
"""A/B sweep of the library's tuning knobs on one GPU (round-2 tool; not part of bench.py's contract).

libsphb reads its knobs from the environment when a handle is created (sphb.cu, create_common):
  SPHB_GUESS_MARGIN  kNN search-radius margin over the previous h      (default 0.02)
  SPHB_KNN_CAP       column slots per lane of k_knn_tile               (47 fp64 / 50 fp32)
  SPHB_KNN_NCW       staged candidates per warp                        (224 fp64 / 256 fp32)
  SPHB_CELL_PER_H    cell row height in units of the mean h            (1.15)
  SPHB_CELL_ASPECT   cell width / height                               (0.32)
  SPHB_FORCE_NREC    staged neighbour records per force block          (672)
Each setting gets a fresh handle on the same jittered-lattice box (bench.py's C5 share at reduced or full size), one
step 0 + 3 warm-up steps, then K steps timed with CUDA events by the library's phase timers.  Prints one line per
setting: ms per phase, fallback particles per step, and the relative change against the first (default) setting.

  python tools/tune_sweep.py --log2n 23 --precision 64 SPHB_GUESS_MARGIN=0.01,0.015,0.02,0.03 SPHB_KNN_CAP=44,47,52
  (knobs are swept one at a time around the defaults; --grid sweeps the full product)
"""
import argparse
import itertools
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402


def run(setting, pos, box, precision, steps):
    from sphugo_b200 import _lib as L
    saved = {k: os.environ.get(k) for k in setting}
    os.environ.update({k: str(v) for k, v in setting.items()})
    try:
        prm = L.make_params(hor=(0.0, box[0]), ver=(0.0, box[1]), accel=(0.0, 0.2), dt_half=6e-5, precision=precision)  # bench.py's C5 physics
        g = L.Handle(prm, pos, None, np.full(len(pos), 0.01))
        g.step(4)
        g.sync()
        f0 = g.counters()["knn_fallback"]
        acc = {k: 0.0 for k in L.PHASES}
        for _ in range(steps):
            g.step(1)
            pt = g.phase_times()
            for k in L.PHASES:
                acc[k] += pt[k]
        g.sync()
        f1 = g.counters()["knn_fallback"]
        g.close()
        out = {k: acc[k] / steps for k in L.PHASES}
        out["fallback_per_step"] = (f1 - f0) / steps
        return out
    finally:
        for k, v in saved.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--log2n", type=int, default=23)
    ap.add_argument("--precision", type=int, default=64)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--grid", action="store_true")
    ap.add_argument("knobs", nargs="*", help="NAME=v1,v2,...")
    a = ap.parse_args()
    from sphugo_b200 import build, gen
    build.build()
    nx = 1 << ((a.log2n + 1) // 2)
    ny = 1 << (a.log2n // 2)
    s = 2.0 ** -14  # bench.py's C5 lattice spacing
    box = (nx * s, ny * s)
    pos = gen.jittered_lattice(nx, ny, (0.0, 0.0), box, 0.25)
    knobs = [(k.split("=")[0], k.split("=")[1].split(",")) for k in a.knobs]
    settings = [{}]
    if a.grid:
        settings += [dict(zip([k for k, _ in knobs], combo)) for combo in itertools.product(*[v for _, v in knobs])]
    else:
        settings += [{k: v} for k, vals in knobs for v in vals]
    base = None
    for st in settings:
        r = run(st, pos, box, a.precision, a.steps)
        base = base or r
        r["setting"] = st or "defaults"
        r["total_vs_default"] = r["total"] / base["total"]
        print(json.dumps(r), flush=True)


if __name__ == "__main__":
    main()
