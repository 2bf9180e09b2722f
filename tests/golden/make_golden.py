#!/usr/bin/env python
"""Regenerates tests/golden/*.npz.

The reference is pure Go and cannot run in this image (no Go toolchain), so the golden vectors are outputs of
the CPU restatement in oracle/ (faithful mode: the reference's own tree walk and queue) on the reference's
configurations.  They pin (a) the oracle against accidental change and (b) the CUDA path on the
GPU box, where neither /root/reference nor a rebuilt oracle is needed to read them.
    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import oracle as orc  # noqa: E402
from sphugo_b200 import gen, gorand  # noqa: E402


def save(name, **arrs):
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **arrs)
    print(name, {k: getattr(v, "shape", None) for k, v in arrs.items()})


def c1_density():
    """examples/density main: 1000 + 200 particles, periodic [0,1]^2, three kernels (density.go:41-97)"""
    ic = gen.spawn([(1000, (0.0, 0.0), (1.0, 1.0)), (200, (0.1, 0.0), (0.3, 0.4))])
    o = orc.Oracle(orc.make_params(hor=(0, 1), ver=(0, 1)), ic["pos"], ids=ic["id"])
    o.knn((0.0, 1.0), (0.0, 1.0), mode=0)
    st = o.state(neighbours=True)
    rho = []
    for k in (0, 1, 2):
        o.density(k)
        rho.append(o.state()["rho"])
    save("c1_density", pos=st["pos"], id=st["id"], h=st["h"], nn_id=np.sort(st["nn_id"], 1).astype(np.int32),
         nn_dist=st["nn_dist"], rho_tophat=rho[0], rho_monaghan=rho[1], rho_wendland=rho[2])


def c2_default():
    """sim.MakeSimulation(): 1000 U([0,1]^2), MakeConfig defaults; states after 1 and 5 steps"""
    ic = gen.spawn([(1000, (0, 0), (1, 1))])
    o = orc.Oracle(orc.make_params(), ic["pos"], ic["vel"], ic["e"], None, ic["id"])
    out = dict(pos0=ic["pos"], id=ic["id"])
    for steps in (1, 5):
        o.step(steps - o.current_step)
        st = o.state()
        for f in ("pos", "vel", "e", "rho", "h", "vdot", "edot"):
            out[f"{f}_{steps}"] = st[f]
    save("c2_default", **out)


def c2_example_config():
    """generated example.sph-config (config-parser.go:872-924), 4 steps"""
    ic = gen.spawn([(260, (0.2, 0.3), (0.8, 0.4)), (700, (0.2, 0.6), (0.8, 0.99))])
    kw = dict(gamma=4.666, particle_mass=1e6, accel=(0.0, 0.55), dt_half=0.00324, kernel=2, hor=(0.2, 0.8),
              ver=(-100.0, 100.0), refl=(orc.OPEN[0], orc.OPEN[1], orc.OPEN[0], 0.99))
    o = orc.Oracle(orc.make_params(**kw), ic["pos"], ic["vel"], ic["e"], None, ic["id"])
    o.step(4)
    st = o.state()
    save("c2_example_config", pos0=ic["pos"], id=ic["id"], **{f: st[f] for f in ("pos", "vel", "e", "rho", "h", "vdot", "edot")})


def c3_c4_small():
    """the two large BASELINE shapes at reduced N, exact-kNN mode (the GPU's contract): C3, a periodic
    jittered lattice (64 x 64) drifting so that particles wrap, 3 steps; C4, a shock tube with a 4:1 number-density
    contrast (8000 particles), 2 steps"""
    pos = gen.jittered_lattice(64, 64)
    n = len(pos)
    vel = np.tile(np.array([[3.0, -2.0]]), (n, 1))
    o = orc.Oracle(orc.make_params(hor=(0.0, 1.0), ver=(0.0, 1.0), accel=(0.0, 0.2), dt_half=0.001), pos, vel, np.full(n, 0.01))
    o.step(3, 1)
    st = o.state()
    save("c3_small", pos0=pos, vel0=vel, **{f: st[f] for f in ("id", "pos", "vel", "e", "rho", "h")})
    pos = gen.shock_tube(8000)
    n = len(pos)
    o = orc.Oracle(orc.make_params(hor=(0.0, 1.0), ver=(0.0, 1.0), dt_half=2e-3), pos, None, np.full(n, 0.01))
    o.step(2, 1)
    st = o.state()
    save("c4_small", pos0=pos, **{f: st[f] for f in ("id", "pos", "vel", "e", "rho", "h")})


def go_scenes():
    """the same two scenes with the particles the Go binary itself spawns: Go's math/rand after rand.Seed(12345678)
    (config-parser.go:60-64), reconstructed in sphugo_b200/gorand.py.  `z` is Particle.Z; ids are spawn indices."""
    a, b = gorand.uniform_rect_spawn(1000), gorand.uniform_rect_spawn(200, (0.1, 0.0), (0.3, 0.4))  # density.go:50-63
    pos, z = np.concatenate([a["pos"], b["pos"]]), np.concatenate([a["z"], b["z"]])
    o = orc.Oracle(orc.make_params(hor=(0, 1), ver=(0, 1)), pos)
    o.knn((0.0, 1.0), (0.0, 1.0), mode=0)
    st = o.state(neighbours=True)
    rho = []
    for k in (0, 1, 2):
        o.density(k)
        rho.append(o.state()["rho"])
    save("c1_density_go", pos=st["pos"], id=st["id"], z=z, h=st["h"], nn_id=np.sort(st["nn_id"], 1).astype(np.int32),
         rho_tophat=rho[0], rho_monaghan=rho[1], rho_wendland=rho[2])
    ic = gorand.uniform_rect_spawn(1000)  # sim.MakeSimulation(), sph.go:23-30
    o = orc.Oracle(orc.make_params(), ic["pos"], ic["vel"], ic["e"])
    out = dict(pos0=ic["pos"], z=ic["z"])
    for steps in (1, 5):
        o.step(steps - o.current_step)
        st = o.state()
        for f in ("pos", "vel", "e", "rho", "h", "vdot", "edot"):
            out[f"{f}_{steps}"] = st[f]
    out["id"] = st["id"]
    save("c2_default_go", **out)
    # the two .sph-config files the reference generates, with their own rectangles (config-parser.go:872-973) on the Go
    # stream; 4 steps.  example: faithful mode (= exact here).  tube: EXACT mode, the GPU's contract - on this scene the
    # reference's own tree walk misses neighbours within the first steps (non-enclosing circle merge, core.go:300-311);
    # `n_faithful_differs` counts the particles whose neighbour set differs between the two modes in the INITIAL state
    # (2 of 4700 on the tube scene; the difference then spreads through the forces)
    for name, kw, rects in REAL_CONFIGS:
        parts = [gorand.uniform_rect_spawn(n, ul, lr) for n, ul, lr in rects]
        pos = np.concatenate([p["pos"] for p in parts])
        z = np.concatenate([p["z"] for p in parts])
        sets = []
        for mode in (0, 1):
            o = orc.Oracle(orc.make_params(**kw), pos, None, np.full(len(pos), 0.01))
            o.knn(mode=mode)
            sets.append(np.sort(o.state(neighbours=True)["nn_id"], 1))
            o.close()
        o = orc.Oracle(orc.make_params(**kw), pos, None, np.full(len(pos), 0.01))
        o.step(4, 0 if name == "c2_example_config_go" else 1)
        st = o.state()
        o.close()
        save(name, pos0=pos, z=z, n_faithful_differs=np.array(int((sets[0] != sets[1]).any(1).sum())),
             **{f: st[f] for f in ("id", "pos", "vel", "e", "rho", "h", "vdot", "edot")})


O = orc.OPEN
REAL_CONFIGS = [
    ("c2_example_config_go", dict(gamma=4.666, particle_mass=1e6, accel=(0.0, 0.55), dt_half=0.00324, kernel=2, hor=(0.2, 0.8),
                                  ver=(-100.0, 100.0), refl=(O[0], O[1], O[0], 0.99)),
     [(260, (0.6, 0.2), (0.79, 0.3)), (700, (0.27, 0.3), (0.4, 0.9))]),
    ("c2_tube_config_go", dict(gamma=4.666, particle_mass=1e5, accel=(0.0, 0.05), dt_half=0.00424, kernel=2,
                               refl=(0.2, O[1], 0.25, 0.5)),
     [(4000, (0.3, 0.3), (0.7, 0.4)), (700, (0.3, 0.3), (0.7, 0.5))]),
]


if __name__ == "__main__":
    if sys.argv[1:] == ["go"]:
        go_scenes()
        sys.exit(0)
    c1_density()
    c2_default()
    c2_example_config()
    c3_c4_small()
    go_scenes()
