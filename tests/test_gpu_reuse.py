"""-m gpu: certified reuse of the neighbour lists (sphb_kernels.cuh ReuseState, sphb_reuse.cuh).

A reuse evaluation takes the exact kNN of every particle from the <= 48 candidates stored by the last rebuild and
certifies it with  h' + D < dexcl;  refused particles go to the ring-expansion search on the stale cells.  Accepted or
refused, the result must be the reference's contract (nearest-neighbour.go:28-165): the exact 32 nearest.  These tests
step across several rebuild / reuse cycles and check EVERY step's neighbour sets and smoothing lengths against an
independent exact search (scipy cKDTree on the evaluation positions x + v dt, which the host can form bit for bit from
the state before the step), and the trajectories against the oracle and against a run with the reuse switched off."""
import contextlib
import os

import numpy as np
import pytest
from scipy.spatial import cKDTree

from tests import util as U
from oracle import oracle as orc
from sphugo_b200 import _lib as L
from sphugo_b200 import gen

pytestmark = pytest.mark.gpu


@contextlib.contextmanager
def env(**kw):
    old = {k: os.environ.get(k) for k in kw}
    os.environ.update({k: str(v) for k, v in kw.items()})
    try:
        yield
    finally:
        for k, v in old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v


def exact_knn(p, hor, ver):
    """ids of the 32 nearest others and the 33 smallest distances, through the 3 x 3 periodic images where hor / ver are
    periodic (nearest-neighbour.go:57-61), from a cKDTree over the image copies"""
    n = len(p)
    sx = (0.0,) if hor[0] == L.OPEN_LO else (hor[0] - hor[1], 0.0, hor[1] - hor[0])
    sy = (0.0,) if ver[0] == L.OPEN_LO else (ver[0] - ver[1], 0.0, ver[1] - ver[0])
    pts = np.concatenate([p + np.array([a, b]) for a in sx for b in sy])
    ids = np.tile(np.arange(n), len(sx) * len(sy))
    d, j = cKDTree(pts).query(p, k=34)
    assert (d[:, 0] == 0.0).all() and (d[:, 1] > 0.0).all()  # column 0 is the particle itself (no coincident particles)
    return ids[j[:, 1:33]], d[:, 1:34]


def _cycle_case(pos, steps, period, vel=None, precision=64, h_tol=1e-13, check_oracle=True, **cfg):
    n = len(pos)
    ic_e = np.full(n, 0.01)
    po, pg = U.params_pair(**cfg)
    pg.precision = precision
    hor, ver = tuple(pg.hor), tuple(pg.ver)
    with env(SPHB_REUSE_PERIOD=period):
        g = L.Handle(pg, pos, vel, ic_e)
    with env(SPHB_REUSE=0):
        g0 = L.Handle(pg, pos, vel, ic_e)
    o = orc.Oracle(po, pos, vel, ic_e) if check_oracle else None
    dt = pg.dt_half
    for k in range(steps):
        st = g.state(("pos", "vel", "id"))
        # evaluation positions of the coming step: drift-1, product then sum (sph.go:112-113)
        p_eval = st["pos"] + st["vel"] * dt
        if hor[0] != L.OPEN_LO:
            p_eval[:, 0] = hor[0] + np.mod(p_eval[:, 0] - hor[0], hor[1] - hor[0])
        if ver[0] != L.OPEN_LO:
            p_eval[:, 1] = ver[0] + np.mod(p_eval[:, 1] - ver[0], ver[1] - ver[0])
        want_i, want_d = exact_knn(p_eval, hor, ver)
        g.step(1)
        g0.step(1)
        got = g.state(("h", "rho", "id", "pos", "vel", "e", "nn_idx", "nn_dist"))
        assert np.array_equal(got["id"], st["id"])
        # (1) exactness, every step: the 32 nearest of the evaluation positions
        assert np.abs(got["h"] / want_d[:, 31] - 1.0).max() <= (h_tol if precision == 64 else 2e-6), f"h, step {k + 1}"
        bad = np.nonzero((np.sort(got["nn_id"], 1) != np.sort(want_i, 1)).any(1))[0]
        for a in bad:  # only a tie between the 32nd and the 33rd may differ (fp32 build: a near tie)
            gap = want_d[a, 32] / want_d[a, 31] - 1.0
            extra = set(got["nn_id"][a]) ^ set(want_i[a])
            assert len(extra) == 2 and gap <= (1e-15 if precision == 64 else 4e-6), \
                f"neighbour set of particle {a}, step {k + 1}: {sorted(extra)}, gap {gap:.3g}"
        # (2) the run without reuse: same sets, same numbers up to the summation order of the density / force terms
        ref = g0.state(("h", "rho", "pos", "vel", "e"))
        tol = 1e-10 if precision == 64 else 5e-4
        for f in ("h", "rho", "pos", "e"):
            assert U.rel_err(got[f], ref[f], np.abs(ref[f]).max() * 1e-3) <= tol, (f, k + 1)
        if o is not None and precision == 64:
            o.step(1)
            oref = o.state()
            for f in ("h", "rho", "pos", "e"):
                assert U.rel_err(got[f], oref[f], np.abs(oref[f]).max() * 1e-3) <= (U.TOL64 if k == 0 else 1e-9), (f, k + 1)
    c = g.counters()
    assert c["steps"] == steps and g0.counters()["reuse_steps"] == 0
    if period > 1:  # the first step follows the upload of the particles: a plain rebuild; cycles of `period` after it
        assert c["reuse_steps"] == (steps - 1) - (steps - 1 + period - 1) // period, c
    g.close(); g0.close()
    if o is not None:
        o.close()
    return c


def test_reuse_cycles_periodic_lattice_every_step_exact():
    """bench.py's workload shape in scaled units: 13 steps = 4 rebuilds + 9 reuse evaluations"""
    n = 96
    c = _cycle_case(gen.jittered_lattice(n, n), steps=13, period=5, hor=(0.0, 1.0), ver=(0.0, 1.0), accel=(0.0, 0.2),
                    dt_half=0.98 / n)
    assert c["knn_fallback"] < 0.05 * n * n * 13  # the certificate accepts nearly everything at this period


def test_reuse_long_cycle_refusals_stay_exact():
    """a cycle far longer than the skin allows: most particles are refused near its end and take the stale-cell search"""
    n = 64
    c = _cycle_case(gen.jittered_lattice(n, n), steps=16, period=16, hor=(0.0, 1.0), ver=(0.0, 1.0), accel=(0.0, 0.2),
                    dt_half=0.98 / n, check_oracle=False)
    assert c["knn_fallback"] > 0


def test_reuse_seam_crossing_and_bulk_flow():
    """a fast bulk flow: particles cross the periodic seam inside a cycle (entries change their image, the stale-cell
    lookups of the force staging are shifted by the mean displacement)"""
    n = 64
    pos = gen.jittered_lattice(n, n)
    vel = np.tile([[3.0, -1.7]], (len(pos), 1))
    _cycle_case(pos, steps=12, period=6, vel=vel, hor=(0.0, 1.0), ver=(0.0, 1.0), dt_half=0.5 / n)


def test_reuse_open_box_reflections_nonuniform():
    """open box, reflecting walls, a 6:1 density contrast (the tube config's physics, config-parser.go:926-973)"""
    ic = gen.spawn([(3000, (0.2, 0.25), (0.5, 0.5)), (500, (0.5, 0.25), (0.8, 0.5))])
    _cycle_case(ic["pos"], steps=9, period=4, particle_mass=1e5, accel=(0.0, 0.05), dt_half=0.00424, kernel=2,
                refl=(0.2, L.OPEN_HI, 0.25, 0.5), check_oracle=False)


def test_reuse_tiny_periodic_box_multi_image():
    """a box of a few smoothing lengths: a particle is its own neighbour's neighbour through several images; the reuse
    kernel must refuse (nearest image not unique) and the stale-cell search scan all images"""
    pos = gen.spawn([(60, (0, 0), (1, 1))], seed=11)["pos"]
    _cycle_case(pos, steps=6, period=3, hor=(0.0, 1.0), ver=(0.0, 1.0), dt_half=0.002, check_oracle=False)


def test_reuse_with_the_local_displacement_bound():
    """SPHB_REUSE_LOCAL=1: certificates use the bound from the 3 x 3 block of coarse cells; long cycle, every step exact"""
    n = 96
    with env(SPHB_REUSE_LOCAL=1):
        _cycle_case(gen.jittered_lattice(n, n), steps=14, period=12, hor=(0.0, 1.0), ver=(0.0, 1.0), accel=(0.0, 0.2),
                    dt_half=0.98 / n, check_oracle=False)
        _cycle_case(gen.spawn([(3000, (0, 0), (1, 1))])["pos"], steps=7, period=5, accel=(0.0, 0.2), dt_half=0.002, check_oracle=False)


def test_reuse_fp32_build():
    n = 96
    _cycle_case(gen.jittered_lattice(n, n), steps=9, period=5, precision=32, hor=(0.0, 1.0), ver=(0.0, 1.0),
                accel=(0.0, 0.2), dt_half=0.98 / n)


def test_reuse_adaptive_schedule_and_invalidation():
    """default (adaptive) schedule: reuse evaluations happen, an upload or an append in between forces a rebuild, and the
    trajectory equals the one without reuse"""
    n = 1456  # (handles below 2^21 particles do not reuse on their own)
    pos = gen.jittered_lattice(n, n)
    kw = dict(hor=(0.0, 1.0), ver=(0.0, 1.0), accel=(0.0, 0.2), dt_half=0.98 / n)
    with env(SPHB_REUSE=1):
        g = L.Handle(L.make_params(**kw), pos, None, np.full(len(pos), 0.01), capacity=len(pos) + 64)
    with env(SPHB_REUSE=0):
        g0 = L.Handle(L.make_params(**kw), pos, None, np.full(len(pos), 0.01), capacity=len(pos) + 64)
    for h in (g, g0):
        for _ in range(12):  # (a caller that looks at every step, like simviewer: the schedule learns from finished steps)
            h.step(1); h.sync()
    # two calm rebuilds first, then cycles of 2, 3, 4 evaluations: at least four reuse evaluations in twelve steps
    assert g.counters()["reuse_steps"] >= 4
    extra = np.array([[0.50003, 0.50001], [0.25, 0.75]])
    for h in (g, g0):
        h.append(extra, None, np.full(2, 0.01), None, np.arange(len(pos), len(pos) + 2, dtype=np.int64))
        h.step(5)
        st = h.download(["pos", "vel", "e"])  # device order, as sphb_upload expects
        h.upload(pos=st["pos"], vel=st["vel"] * 1.01, e=st["e"])
        h.step(4)
    a, b = g.state(), g0.state()
    for f in ("pos", "vel", "e", "rho", "h"):
        assert U.rel_err(a[f], b[f], np.abs(b[f]).max() * 1e-3) <= 1e-9, f
    g.close(); g0.close()
