"""CPU: the C oracle against a second, independent numpy restatement of the same Go functions (tests/np_restatement.py).
The reference pins none of kNN / density / force / integrator with a test of its own, so two restatements
written separately from the Go source agreeing bit for bit is the strongest pin available without a Go toolchain."""
import numpy as np
import pytest

from oracle import oracle as orc
from sphugo_b200 import gen, gorand
from tests import np_restatement as npr

FIELDS = ("pos", "vel", "e", "rho", "h", "c", "vdot", "edot", "vpred", "epred")


def _both(pos, vel, e, steps, mode, **kw):
    cfg = dict(dt_half=0.001, gamma=1.66666, particle_mass=1.0, accel=(0.0, 0.0), hor=(-npr.MAXF, npr.MAXF), ver=(-npr.MAXF, npr.MAXF),
               refl=(-npr.MAXF, npr.MAXF, -npr.MAXF, npr.MAXF), kernel=1)
    cfg.update(kw)
    o = orc.Oracle(orc.make_params(**cfg), pos, vel, e)
    st = npr.make_state(pos, vel, e)
    for k in range(steps):
        o.step(1, mode)
        npr.step(st, cfg)
        ref = o.state()  # sorted by id = spawn index, the numpy state's order
        for f in FIELDS:
            assert np.array_equal(ref[f], st[f]), (f, k, float(np.abs(ref[f] - st[f]).max()))
    o.close()


@pytest.mark.parametrize("mode", [0, 1])
def test_default_scene_bit_for_bit(mode):
    """sim.MakeSimulation(): the Go stream's 1000 particles, MakeConfig defaults, open box (sph.go:23-30)"""
    ic = gorand.uniform_rect_spawn(1000)
    _both(ic["pos"], ic["vel"], ic["e"], 3, mode)


def test_example_config_bit_for_bit():
    """generated example.sph-config (config-parser.go:872-924): Wendland, periodic x, gravity, the floor reflection"""
    a, b = gorand.uniform_rect_spawn(260, (0.2, 0.3), (0.8, 0.4)), gorand.uniform_rect_spawn(700, (0.2, 0.6), (0.8, 0.99))
    pos = np.concatenate([a["pos"], b["pos"]])
    _both(pos, np.zeros_like(pos), np.full(len(pos), 0.01), 4, 0, gamma=4.666, particle_mass=1e6, accel=(0.0, 0.55), dt_half=0.00324,
          kernel=2, hor=(0.2, 0.8), ver=(-100.0, 100.0), refl=(-npr.MAXF, npr.MAXF, -npr.MAXF, 0.99))


def test_periodic_box_with_wraps_and_reflections_bit_for_bit():
    """periodic box, particles drifting fast enough to wrap on both axes (the `continue` quirk, sph.go:147-167)"""
    pos = gen.jittered_lattice(36, 36)
    vel = np.tile(np.array([[9.0, -7.0]]), (len(pos), 1))
    _both(pos, vel, np.full(len(pos), 0.01), 3, 1, hor=(0.0, 1.0), ver=(0.0, 1.0), accel=(0.0, 0.2), dt_half=0.002)
    # walls on all four sides inside an open box: particles leave through the walls and are put back (sph.go:170-193)
    vel = (gen.uniform01(5, 2 * len(pos)).reshape(-1, 2) - 0.5) * 40.0
    _both(pos, vel, np.full(len(pos), 0.01), 3, 0, dt_half=0.002, refl=(0.02, 0.97, 0.03, 0.98))
