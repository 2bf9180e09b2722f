import sys, numpy as np
sys.path.insert(0, '.')
from sphugo_b200 import _lib as L, gen, slab
pos = gen.jittered_lattice(128, 64); n = len(pos)
vel = np.tile([[6.0, 1.0]], (n, 1)); e = np.full(n, 0.01)
kw = dict(hor=(0.0, 1.0), ver=(0.0, 1.0), dt_half=0.002)
pg = L.make_params(**kw)
g = L.Handle(pg, pos, vel, e)
topo = slab.Topology(4, [0, .25, .5, .75, 1.0], True)
sim = slab.LocalSlabSim(pg, topo, pos, vel, e, h_max_hint=slab.default_h_hint(n, 1.0))
F = ["pos", "vel", "rho", "c", "e", "edot", "vdot", "h", "id"]
for k in range(3):
    before = [s.h.download(["id"])["id"].copy() for s in sim.slabs]
    g.step(1); sim.step(1)
    a, b = g.state(F), sim.state(F)
    after = [set(s.h.download(["id"])["id"].tolist()) for s in sim.slabs]
    arrived = set().union(*[after[r] - set(before[r].tolist()) for r in range(4)])
    for f in ["pos", "vel", "vdot", "e", "edot", "rho", "h"]:
        d = np.abs(a[f] - b[f]); d = d.reshape(n, -1).max(1)
        bad = np.nonzero(d > 1e-9 * max(1e-30, np.abs(a[f]).max()))[0]
        print(k + 1, f, d.max(), len(bad), "arrived", len(arrived), "bad&arrived", len(set(bad.tolist()) & arrived), "x of bad", np.round(a["pos"][bad[:6], 0], 3))
    print(sim.counts())
