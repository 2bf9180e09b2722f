#!/usr/bin/env python
"""Small workload that touches every kernel of libsphb.so, for compute-sanitizer:

  compute-sanitizer --tool memcheck  --error-exitcode 9 python tools/sanitize_workload.py
  compute-sanitizer --tool racecheck --error-exitcode 9 python tools/sanitize_workload.py

Both builds: non-uniform periodic box (group halving, force rows -2..+2, fallback), open box (statistics pass,
clamped border cells), frame / neighbour-list download, a 2-slab ring with migration (ghost packing, in-place
ghost removal), and appends inside and past the capacity (sources)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402

from sphugo_b200 import _lib as L, gen, slab  # noqa: E402

for prec in (64, 32):
    pos = gen.shock_tube(6000)
    g = L.Handle(L.make_params(hor=(0.0, 1.0), ver=(0.0, 1.0), dt_half=2e-3, precision=prec), pos, None, np.full(len(pos), 0.01))
    g.step(2); g.sync(); g.frame(64, 64, ids=False); g.state(["pos", "rho", "nn_idx"]); g.close()
    ic = gen.spawn([(1500, (0, 0), (1, 1))])
    g = L.Handle(L.make_params(precision=prec, accel=(0.0, 0.2)), ic["pos"], None, ic["e"])
    g.step(2); g.sync(); g.close()
    pos = gen.jittered_lattice(96, 96)
    n = len(pos)
    pg = L.make_params(hor=(0.0, 1.0), ver=(0.0, 1.0), dt_half=0.002, precision=prec)
    sim = slab.LocalSlabSim(pg, slab.Topology(2, [0.0, 0.5, 1.0], True), pos, np.tile([[3.0, 1.0]], (n, 1)), np.full(n, 0.01),
                            h_max_hint=slab.default_h_hint(n, 1.0))
    sim.step(3)
    assert sum(sim.counts()) == n
    sim.close()
    # sources: append inside the capacity, then past it (reallocation at twice the size), a step after each
    ic = gen.spawn([(1200, (0, 0), (1, 1))])
    g = L.Handle(L.make_params(precision=prec, accel=(0.0, 0.2), dt_half=0.002), ic["pos"], None, ic["e"], capacity=1300)
    g.step(2)
    extra = gen.uniform_rect(1500, (0.2, 0.2), (0.8, 0.8), seed=99)
    for lo, hi in ((0, 100), (100, 1500)):
        g.append(extra[lo:hi], None, np.full(hi - lo, 0.01), None, np.arange(1200 + lo, 1200 + hi, dtype=np.int64))
        g.step(1); g.sync()
    assert g.n == 2700
    g.frame(64, 64, ids=False); g.close()
    # the same on a periodic box, where steps are fused (the force epilogue prepares the next step's keys and grid)
    pos = gen.jittered_lattice(48, 48)
    g = L.Handle(L.make_params(precision=prec, hor=(0.0, 1.0), ver=(0.0, 1.0), dt_half=0.002), pos, None, np.full(len(pos), 0.01))
    g.step(3)
    extra = gen.jittered_lattice(20, 20, jitter=0.4, seed=5)
    g.append(extra, None, np.full(len(extra), 0.01), None, np.arange(len(pos), len(pos) + len(extra), dtype=np.int64))
    g.step(2); g.sync()
    assert g.n == len(pos) + len(extra)
    g.close()
# list reuse (rebuild on the wide block + annulus pass, reuse evaluations, stale-cell fallback), single handle and ring
os.environ["SPHB_REUSE_PERIOD"] = "4"
for prec in (64, 32):
    for pos, kw in ((gen.jittered_lattice(64, 64), dict(hor=(0.0, 1.0), ver=(0.0, 1.0), accel=(0.0, 0.2), dt_half=0.01)),
                    (gen.shock_tube(6000), dict(hor=(0.0, 1.0), ver=(0.0, 1.0), dt_half=2e-3)),
                    (gen.spawn([(3000, (0, 0), (1, 1))])["pos"], dict(accel=(0.0, 0.2), dt_half=0.002))):
        g = L.Handle(L.make_params(precision=prec, **kw), pos, None, np.full(len(pos), 0.01))
        g.step(7); g.sync()
        assert g.counters()["reuse_steps"] >= 4
        g.state(["pos", "rho", "nn_idx"]); g.close()
    pos = gen.jittered_lattice(96, 96)
    n = len(pos)
    pg = L.make_params(hor=(0.0, 1.0), ver=(0.0, 1.0), dt_half=0.002, precision=prec)
    sim = slab.LocalRingSim(pg, slab.Topology(2, [0.0, 0.5, 1.0], True), pos, np.tile([[3.0, 1.0]], (n, 1)), np.full(n, 0.01),
                            h_max_hint=slab.default_h_hint(n, 1.0))
    sim.step(6)
    sim.state(["pos", "rho"])
    assert sum(sim.counts()) == n
    sim.close()
del os.environ["SPHB_REUSE_PERIOD"]
print("sanitizer workload done")
