#!/usr/bin/env python
"""The reference's own benchmark program, like for like: examples/speed-test (speed-test.go:22-45).

100000 particles of the default spawner (the Go math/rand stream, so the very particles the Go program steps), MakeConfig
defaults with DeltaTHalf = 0.02 and g = (0, 0.2), open boundaries, 20 Step() calls from a fresh simulation (the first one
evaluates the forces twice, sph.go:89-103), "FPS" = steps per second - the only performance figure the reference prints.

    python tools/speed_test.py --impl cpu     # the C restatement of the serial Go path (oracle/), full size, ~20 s
    python tools/speed_test.py --impl gpu     # libsphb.so through the mirrored sim API (fails loudly without a GPU)

Not part of bench.py's contract (SURVEY 8d lists it next to C3 as the like-for-like CPU comparison); prints one JSON line.
"""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--impl", choices=["gpu", "cpu"], default="gpu")
    ap.add_argument("--n", type=int, default=100000)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--precision", type=int, default=64, choices=[64, 32])
    a = ap.parse_args()
    from sphugo_b200 import gorand
    ic = gorand.uniform_rect_spawn(a.n)  # spwn := MakeUniformRectSpawner(); spwn.NParticles = 100000
    fps = []
    if a.impl == "cpu":
        from oracle import oracle as orc
        o = orc.Oracle(orc.make_params(dt_half=0.02, accel=(0.0, 0.2)), ic["pos"], ic["vel"], ic["e"])
        for i in range(a.steps):
            t0 = time.perf_counter()
            o.step(1)
            fps.append(1.0 / (time.perf_counter() - t0))
        total_e = o.total_energy()
        o.close()
        what = "C restatement of the serial Go path (oracle/sph_oracle.c), 1 core"
    else:
        from sphugo_b200 import sim
        conf = sim.MakeConfig()
        conf.DeltaTHalf, conf.Acceleration = 0.02, (0.0, 0.2)
        for run in range(2):  # the first run warms the device up (context, clocks); the second is reported
            s = sim.Simulation(conf, ic, precision=a.precision)
            fps = []
            for i in range(a.steps):
                t0 = time.perf_counter()
                s.Step()
                total_e = s.TotalEnergy()  # waits for the asynchronous step (simviewer reads it every step too)
                fps.append(1.0 / (time.perf_counter() - t0))
            s.Close()
        what = f"libsphb.so (fp{a.precision} build) through sphugo_b200.sim, wall clock per Step() + TotalEnergy()"
    total = sum(1.0 / f for f in fps)
    for i, f in enumerate(fps):
        print(f"Step {i} FPS {f:.6g}", file=sys.stderr)
    print(json.dumps({"benchmark": "examples/speed-test (speed-test.go:22-45)", "impl": a.impl, "what": what, "particles": a.n,
                      "steps": a.steps, "seconds": total, "average_fps": a.steps / total,
                      "particle_updates_per_s": a.n * a.steps / total, "sum_E": total_e}))


if __name__ == "__main__":
    main()
