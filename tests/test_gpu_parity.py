"""-m gpu parity tests: libsphb.so (CUDA, through the C ABI) against the CPU oracle on identical inputs.

Bars (north_star): neighbour index sets bit-exact excluding exact distance ties; h, rho, accelerations and
energies within 1e-12 (fp64 build), accelerations / EDot measured relative to sum|term| because the
32-term sums nearly cancel on near-uniform data."""
import numpy as np
import pytest

from tests import util as U
from oracle import oracle as orc
from sphugo_b200 import _lib as L
from sphugo_b200 import gen

pytestmark = pytest.mark.gpu
TOL = U.TOL64


def _knn_density_case(ic, hor, ver, kernels=(0, 1, 2), mass=1.0):
    po, pg = U.params_pair(hor=hor, ver=ver, particle_mass=mass)
    o = orc.Oracle(po, ic["pos"], ids=ic["id"])
    o.knn(hor, ver, mode=0)
    assert o.underfull == 0
    g = L.Handle(pg, ic["pos"], ids=ic["id"])
    g.knn(hor, ver)
    ref = o.state(neighbours=True)
    got = g.state(U.FIELDS_STATE + U.FIELDS_NN)
    hard, excused = U.neighbour_sets_equal(got, ref)
    assert hard == 0, f"{hard} particles with different neighbour sets ({excused} exact ties excused)"
    assert U.rel_err(got["h"], ref["h"]) <= TOL
    # lists are both sorted by descending distance: distances must agree slot by slot
    assert U.rel_err(got["nn_dist"], ref["nn_dist"]) <= TOL
    assert np.abs(got["nn_pos"] - ref["nn_pos"]).max() <= 1e-15 * max(1.0, np.abs(ref["nn_pos"]).max()) * 4
    for k in kernels:
        o.density(k)
        g.density(k)
        r = o.state()["rho"]
        assert U.rel_err(g.state(["rho"])["rho"], r) <= TOL, f"density kernel {k}"
    g.close(); o.close()


def test_c1_density_example_periodic():
    """examples/density main: N=1200, periodic [0,1]^2, TopHat/Monaghan/Wendland (density.go:41-97)."""
    _knn_density_case(U.c1_density(), (0.0, 1.0), (0.0, 1.0))


def test_c1_periodic_visual_open_and_periodic():
    """periodicVisualTest: N=11200 on [0.1,0.9]^2, open vs periodic, TopHat (density.go:99-141)."""
    ic = U.c1_periodic_visual()
    _knn_density_case(ic, U.OPEN, U.OPEN, kernels=(0,))
    _knn_density_case(ic, (0.1, 0.9), (0.1, 0.9), kernels=(0,))


def test_mixed_open_periodic_axes():
    ic = gen.spawn([(3000, (0.0, 0.0), (1.0, 1.0))], seed=7)
    _knn_density_case(ic, (0.0, 1.0), U.OPEN, kernels=(1,))
    _knn_density_case(ic, U.OPEN, (0.0, 1.0), kernels=(2,))


def test_knn_matches_exact_bruteforce_clustered():
    """non-uniform data (h varies x10): the stencil scan must stay exact (ring-expansion fallback)."""
    rng = np.random.default_rng(3)
    pos = np.concatenate([rng.random((1500, 2)), 0.5 + 0.01 * rng.standard_normal((1500, 2))])
    ic = dict(pos=pos, id=np.arange(len(pos), dtype=np.int64))
    po, pg = U.params_pair()
    o = orc.Oracle(po, ic["pos"], ids=ic["id"])
    o.knn(U.OPEN, U.OPEN, mode=1)
    g = L.Handle(pg, ic["pos"], ids=ic["id"])
    g.knn(U.OPEN, U.OPEN)
    ref, got = o.state(neighbours=True), g.state(U.FIELDS_STATE + U.FIELDS_NN)
    hard, _ = U.neighbour_sets_equal(got, ref)
    assert hard == 0
    assert U.rel_err(got["h"], ref["h"]) <= TOL


def _step_case(ic, steps, check_every=1, tol_first=TOL, tol_traj=1e-9, **cfg):
    po, pg = U.params_pair(**cfg)
    o = orc.Oracle(po, ic["pos"], ic.get("vel"), ic.get("e"), None, ic["id"])
    g = L.Handle(pg, ic["pos"], ic.get("vel"), ic.get("e"), None, ic["id"])
    done = 0
    while done < steps:
        k = min(check_every, steps - done)
        o.step(k)
        g.step(k)
        done += k
        ref = o.state(neighbours=True)
        got = g.state(U.FIELDS_STATE)
        assert g.current_step == o.current_step == done
        tol = tol_first if done == 1 else tol_traj
        asc, esc = U.force_scales(ref, po)
        L_ = max(1.0, float(np.abs(ref["pos"]).max()))
        assert np.abs(got["pos"] - ref["pos"]).max() <= tol * L_, f"pos step {done}"
        assert U.rel_err(got["h"], ref["h"]) <= tol, f"h step {done}"
        assert U.rel_err(got["rho"], ref["rho"]) <= tol, f"rho step {done}"
        assert U.rel_err(got["c"], ref["c"]) <= tol, f"c step {done}"
        assert U.rel_err(got["vdot"], ref["vdot"], asc) <= tol, f"vdot step {done}"
        assert U.rel_err(got["edot"], ref["edot"], esc) <= tol, f"edot step {done}"
        vs = np.abs(ref["vel"]).max() + asc.max() * 2 * po.dt_half
        assert np.abs(got["vel"] - ref["vel"]).max() <= tol * vs, f"vel step {done}"
        assert U.rel_err(got["e"], ref["e"], esc * 2 * po.dt_half) <= tol, f"e step {done}"
        assert abs(g.reduce(L.SUM_E) - o.total_energy()) <= 1e-12 * abs(o.total_energy())
        assert abs(g.reduce(L.SUM_RHO) - o.total_density()) <= 1e-12 * abs(o.total_density())
    g.close(); o.close()


def test_c2_default_simulation_steps():
    """sim.MakeSimulation(): 1000 U([0,1]^2), MakeConfig defaults, open boundaries (sph.go:23-30)."""
    _step_case(gen.spawn([(1000, (0, 0), (1, 1))]), steps=10)


def test_c2_example_config_steps():
    """the physics of the generated example.sph-config (config-parser.go:872-924: Wendland, periodic x, gravity, floor) on
    two wide rectangles; the config's own rectangles on the Go stream are tests/golden/c2_example_config_go.npz"""
    ic = gen.spawn([(260, (0.2, 0.3), (0.8, 0.4)), (700, (0.2, 0.6), (0.8, 0.99))])
    _step_case(ic, steps=8, gamma=4.666, particle_mass=1e6, accel=(0.0, 0.55), dt_half=0.00324, kernel=2,
               hor=(0.2, 0.8), ver=(-100.0, 100.0), refl=(L.OPEN_LO, L.OPEN_HI, L.OPEN_LO, 0.99))


def test_c2_tube_config_steps():
    """the physics of the generated tube.sph-config (config-parser.go:926-973: reflections L, U, D) on a dense and a dilute
    rectangle side by side; the config's own rectangles on the Go stream are tests/golden/c2_tube_config_go.npz"""
    ic = gen.spawn([(4000, (0.2, 0.25), (0.5, 0.5)), (700, (0.5, 0.25), (0.8, 0.5))])
    _step_case(ic, steps=6, particle_mass=1e5, accel=(0.0, 0.05), dt_half=0.00424, kernel=2,
               refl=(0.2, L.OPEN_HI, 0.25, 0.5))


def test_speed_test_shape_small():
    """examples/speed-test shape at reduced N: open box, g = (0, 0.2) (speed-test.go:22-30)."""
    ic = gen.spawn([(20000, (0, 0), (1, 1))])
    _step_case(ic, steps=3, accel=(0.0, 0.2), dt_half=0.002)


def test_periodic_box_steps_with_wrap():
    """C3 shape at reduced N: periodic [0,1]^2, jittered lattice, moving so that particles wrap."""
    pos = gen.jittered_lattice(128, 128)
    n = len(pos)
    vel = np.tile(np.array([[3.0, -2.0]]), (n, 1))
    ic = dict(pos=pos, vel=vel, e=np.full(n, 0.01), id=np.arange(n, dtype=np.int64))
    _step_case(ic, steps=5, hor=(0.0, 1.0), ver=(0.0, 1.0), accel=(0.0, 0.2), dt_half=0.001)


def test_errors_and_edge_cases():
    po, pg = U.params_pair()
    with pytest.raises(L.SphbError) as ei:  # half-open axis panics in the reference (nearest-neighbour.go:44)
        L.Handle(L.make_params(hor=(L.OPEN_LO, 1.0)), np.random.rand(100, 2))
    assert ei.value.code == L.E_INVALID
    g = L.Handle(pg, np.random.default_rng(0).random((20, 2)))
    with pytest.raises(L.SphbError) as ei:  # fewer than 32 candidates
        g.knn(U.OPEN, U.OPEN)
    assert ei.value.code == L.E_KNN_UNDERFULL
    g.close()
    g = L.Handle(L.make_params(kernel=0), np.random.default_rng(0).random((200, 2)))
    with pytest.raises(L.SphbError) as ei:  # TopHat2D.DF panics (sph.go:251)
        g.step(1)
    assert ei.value.code == L.E_KERNEL
    with pytest.raises(L.SphbError) as ei:
        g.density(1)
    assert ei.value.code == L.E_STATE
    g.close()
    # empty simulation: Step panics "Simulation not initialized" (sph.go:92-94)
    g = L.Handle(pg, np.zeros((0, 2)), capacity=64)
    with pytest.raises(L.SphbError):
        g.step(1)
    g.close()


def test_tiny_periodic_box_multi_image():
    """N = 40 in a periodic box: the 3x3 images supply neighbours, one particle can appear via two images."""
    ic = gen.spawn([(40, (0, 0), (1, 1))], seed=5)
    po, pg = U.params_pair(hor=(0.0, 1.0), ver=(0.0, 1.0))
    o = orc.Oracle(po, ic["pos"], ids=ic["id"])
    o.knn((0.0, 1.0), (0.0, 1.0), mode=1)
    g = L.Handle(pg, ic["pos"], ids=ic["id"])
    g.knn((0.0, 1.0), (0.0, 1.0))
    ref, got = o.state(neighbours=True), g.state(U.FIELDS_STATE + U.FIELDS_NN)
    assert U.rel_err(got["h"], ref["h"]) <= TOL
    assert U.rel_err(got["nn_dist"], ref["nn_dist"]) <= TOL
    assert (np.sort(got["nn_id"], 1) == np.sort(ref["nn_id"], 1)).all()


def _golden(name):
    import os
    return np.load(os.path.join(os.path.dirname(__file__), "golden", name + ".npz"))


def test_golden_c1_density():
    """committed golden vectors (tests/golden/make_golden.py): kNN sets, h and the three densities"""
    g = _golden("c1_density")
    h = L.Handle(L.make_params(hor=(0.0, 1.0), ver=(0.0, 1.0)), g["pos"], ids=g["id"])
    h.knn((0.0, 1.0), (0.0, 1.0))
    st = h.state(U.FIELDS_STATE + U.FIELDS_NN)
    assert (np.sort(st["nn_id"], 1) == g["nn_id"]).all()
    assert U.rel_err(st["h"], g["h"]) <= TOL
    assert U.rel_err(st["nn_dist"], g["nn_dist"]) <= TOL
    for k, name in ((0, "rho_tophat"), (1, "rho_monaghan"), (2, "rho_wendland")):
        h.density(k)
        assert U.rel_err(h.state(["rho"])["rho"], g[name]) <= TOL, name


def test_golden_c2_default_and_example_config():
    g = _golden("c2_default")
    n = len(g["pos0"])
    h = L.Handle(L.make_params(), g["pos0"], None, np.full(n, 0.01), None, g["id"])
    h.step(1)
    st = h.state()
    for f in ("pos", "vel", "e", "rho", "h"):
        assert U.rel_err(st[f], g[f + "_1"], np.abs(g[f + "_1"]).max() * 1e-3) <= TOL, f
    h.step(4)
    st = h.state()
    for f in ("pos", "vel", "e", "rho", "h"):
        assert U.rel_err(st[f], g[f + "_5"], np.abs(g[f + "_5"]).max() * 1e-3) <= 1e-9, f
    g = _golden("c2_example_config")
    n = len(g["pos0"])
    kw = dict(gamma=4.666, particle_mass=1e6, accel=(0.0, 0.55), dt_half=0.00324, kernel=2, hor=(0.2, 0.8),
              ver=(-100.0, 100.0), refl=(L.OPEN_LO, L.OPEN_HI, L.OPEN_LO, 0.99))
    h = L.Handle(L.make_params(**kw), g["pos0"], None, np.full(n, 0.01), None, g["id"])
    h.step(4)
    st = h.state()
    for f in ("pos", "vel", "e", "rho", "h"):
        assert U.rel_err(st[f], g[f], np.abs(g[f]).max() * 1e-3) <= 1e-9, f


def test_full_size_properties_c3():
    """BASELINE size (2^20, periodic): properties that need no oracle: every list holds 32 distinct real
    neighbours, h equals the largest listed distance, kNN is idempotent (recomputing from scratch with a cold
    radius guess gives identical h), total energy and density are finite and positive."""
    pos = gen.jittered_lattice(1024, 1024)
    n = len(pos)
    h = L.Handle(L.make_params(hor=(0.0, 1.0), ver=(0.0, 1.0), accel=(0.0, 0.2), dt_half=0.001), pos, None, np.full(n, 0.01))
    h.step(2)
    h1 = h.state(["h", "rho"])
    assert np.isfinite(h1["h"]).all() and (h1["h"] > 0).all() and (h1["rho"] > 0).all()
    h.knn((0.0, 1.0), (0.0, 1.0))
    st = h.state(["h", "nn_idx", "nn_dist", "pos"])
    assert (st["nn_idx"] >= 0).all()
    srt = np.sort(st["nn_idx"], 1)
    assert (np.diff(srt, axis=1) > 0).all()  # 32 distinct neighbours
    assert np.array_equal(st["nn_dist"][:, 0], st["h"]) or U.rel_err(st["nn_dist"][:, 0], st["h"]) <= 1e-15
    assert (np.diff(st["nn_dist"], axis=1) <= 0).all()
    # cold start on the same positions: a fresh handle has no radius guess at all
    g2 = L.Handle(L.make_params(hor=(0.0, 1.0), ver=(0.0, 1.0)), st["pos"], ids=st["id"])
    g2.knn((0.0, 1.0), (0.0, 1.0))
    s2 = g2.state(["h"])
    assert U.rel_err(s2["h"], st["h"]) <= 1e-15
    # mean neighbour count check: pi h^2 n ~ 33
    assert abs(np.mean(np.pi * st["h"] ** 2 * n) - 33) < 1.5
    assert np.isfinite(h.reduce(L.SUM_E)) and h.reduce(L.SUM_RHO) > 0


def test_non_finite_input_does_not_fault_the_device():
    """garbage in the mutable fields (Root.Particles is public: density.go:12-15) may give garbage results or a
    status code, but never an illegal memory access: the next handle on the same device must work."""
    pos = gen.jittered_lattice(64, 64)
    n = len(pos)
    vel = np.zeros((n, 2))
    vel[::7] = [1e300, -1e300]
    vel[5] = [np.nan, np.inf]
    for prec in (64, 32):
        pg = L.make_params(hor=(0.0, 1.0), ver=(0.0, 1.0), dt_half=0.001, precision=prec)
        g = L.Handle(pg, pos, vel, np.full(n, 0.01))
        try:
            g.step(2)
            g.sync()
        except L.SphbError as ex:
            assert ex.code != L.E_CUDA, ex
        g.close()
        g2 = L.Handle(pg, pos, None, np.full(n, 0.01))
        g2.step(1)
        assert np.isfinite(g2.state(["rho"])["rho"]).all()
        g2.close()


def test_frame_extraction_matches_the_animator_formula():
    """sphb_frame against (*Animator).CurrentFrame's per-particle arithmetic (animator.go:75-101) on oracle state"""
    ic = gen.spawn([(260, (0.2, 0.3), (0.8, 0.4)), (700, (0.2, 0.6), (0.8, 0.99))])
    kw = dict(gamma=4.666, particle_mass=1e6, accel=(0.0, 0.55), dt_half=0.00324, kernel=2,
              hor=(0.2, 0.8), ver=(-100.0, 100.0), refl=(L.OPEN_LO, L.OPEN_HI, L.OPEN_LO, 0.99))
    po, pg = U.params_pair(**kw)
    o = orc.Oracle(po, ic["pos"], ic["vel"], ic["e"], None, ic["id"])
    g = L.Handle(pg, ic["pos"], ic["vel"], ic["e"], None, ic["id"])
    o.step(3); g.step(3)
    ref = o.state()
    fr = g.frame(1280, 720)
    order = np.argsort(fr["id"], kind="stable")
    n = len(ref["id"])
    x = ref["pos"][:, 0].astype(np.float32) * np.float32(1280)
    y = ref["pos"][:, 1].astype(np.float32) * np.float32(720)
    cf = ref["rho"] / (po.particle_mass * float(n * 10)) * 256
    ci = np.minimum(cf, 255).astype(np.uint8)
    assert (fr["id"][order] == ref["id"]).all()
    assert np.abs(fr["xy"][order, 0] - x).max() <= 1e-3 and np.abs(fr["xy"][order, 1] - y).max() <= 1e-3
    # rho agrees to 1e-9 along the trajectory: the truncated index may differ only where cf sits on an integer
    diff = fr["colour"][order].astype(int) - ci.astype(int)
    near = np.abs(cf - np.round(cf)) < 1e-6
    assert (diff[~near] == 0).all() and np.abs(diff).max() <= 1
    assert ci.max() > ci.min()  # the scene spans several ramp entries
    g.close(); o.close()


def test_by_id_transfers_follow_the_callers_order():
    """sphb_upload_by_id / sphb_frame(id_out = NULL): host arrays indexed by particle id, whatever the device order"""
    pos = gen.jittered_lattice(48, 48)
    n = len(pos)
    kw = dict(hor=(0.0, 1.0), ver=(0.0, 1.0), accel=(0.0, 0.2), dt_half=0.002)
    pg = L.make_params(**kw)
    a = L.Handle(pg, pos, None, np.full(n, 0.01))   # ids = 0..n-1
    b = L.Handle(pg, pos, None, np.full(n, 0.01))
    a.step(3); b.step(3)
    st = a.state(["pos", "vel", "e", "id"])          # sorted by id
    # b receives a's state in id order although its own device order is a cell order
    b.upload_by_id(pos=st["pos"], vel=st["vel"], e=st["e"])
    a.step(2); b.step(2)
    sa, sb = a.state(["pos", "vel", "e", "rho", "id"]), b.state(["pos", "vel", "e", "rho", "id"])
    # (a is in the middle of a list-reuse cycle, b rebuilt its lists after the upload: same neighbours, but the 32-term
    # sums run in a different order)
    for f in ("pos", "vel", "e", "rho"):
        assert U.rel_err(sa[f], sb[f], np.abs(sb[f]).max() * 1e-3) <= 1e-12, f
    f_dev, f_id = a.frame(640, 480, ids=True), a.frame(640, 480, ids=False)
    o = np.argsort(f_dev["id"], kind="stable")
    assert np.array_equal(f_dev["xy"][o], f_id["xy"]) and np.array_equal(f_dev["colour"][o], f_id["colour"])
    a.close(); b.close()
    # sparse ids are refused
    c = L.Handle(pg, pos, None, np.full(n, 0.01), None, np.arange(n, dtype=np.int64) * 3 + 7)
    with pytest.raises(L.SphbError) as ei:
        c.frame(640, 480, ids=False)
    assert ei.value.code == L.E_STATE
    c.close()


def test_shock_tube_density_contrast_steps():
    """C4 shape at reduced N (BASELINE configs[3]): number-density ratio 4:1 across x = 0.5, periodic box.  The grid
    follows the mean h, so the dilute half has h ~ 2 cell rows: kNN tiles whose union block does not fit the staging
    area are halved and retried, force blocks stage rows -2..+2 - and only a small share of the particles may need the
    ring-expansion fallback."""
    pos = gen.shock_tube(40000)
    n = len(pos)
    ic = dict(pos=pos, vel=np.zeros((n, 2)), e=np.full(n, 0.01), id=np.arange(n, dtype=np.int64))
    cfg = dict(hor=(0.0, 1.0), ver=(0.0, 1.0), dt_half=2e-3)
    _step_case(ic, steps=3, **cfg)
    g = L.Handle(L.make_params(**cfg), pos, None, ic["e"])
    g.step(3)
    g.sync()
    c = g.counters()
    assert c["knn_fallback"] < 0.05 * n * 4, c  # 4 evaluations; the first one guesses its radii from the cell counts
    g.close()


def test_host_calls_between_steps_do_not_disturb_the_prepared_keys():
    """after a periodic step the next step's cell keys are already on the device (force epilogue): read-only calls
    in between (frame by id incl. its first-use dense-id check, reductions, downloads) must leave them intact, and
    state-changing calls must discard them"""
    pos = gen.jittered_lattice(64, 64)
    n = len(pos)
    kw = dict(hor=(0.0, 1.0), ver=(0.0, 1.0), accel=(0.0, 0.2), dt_half=0.002)
    a = L.Handle(L.make_params(**kw), pos, None, np.full(n, 0.01))
    b = L.Handle(L.make_params(**kw), pos, None, np.full(n, 0.01))
    a.step(4)
    for k in range(4):
        b.step(1)
        b.frame(320, 200, ids=False)        # first call runs the dense-id check
        b.reduce(L.SUM_E); b.state(["pos", "nn_idx"])
    sa, sb = a.state(), b.state()
    for f in ("pos", "vel", "e", "rho", "h"):
        assert np.array_equal(sa[f], sb[f]), f
    # a parameter change (dt) and an upload invalidate the prepared keys
    st = b.state(["pos", "vel", "e", "id"])
    p2 = L.make_params(**dict(kw, dt_half=0.001))
    a.set_params(p2); b.set_params(p2)
    b.upload_by_id(pos=st["pos"], vel=st["vel"], e=st["e"])   # same values: only the bookkeeping differs
    a.step(2); b.step(2)
    sa, sb = a.state(), b.state()
    for f in ("pos", "vel", "e", "rho", "h"):
        assert np.array_equal(sa[f], sb[f]), f
    a.close(); b.close()


@pytest.mark.parametrize("precision,tol", [(64, 1e-9), (32, 1e-4)])
def test_golden_c3_c4_small(precision, tol):
    """committed golden vectors of the two large BASELINE shapes at reduced N (tests/golden/make_golden.py), both
    builds: fp64 to 1e-9 along the 2-3 step trajectory, fp32 to 1e-4 (ten times its one-step bar)"""
    g = _golden("c3_small")
    n = len(g["pos0"])
    h = L.Handle(L.make_params(hor=(0.0, 1.0), ver=(0.0, 1.0), accel=(0.0, 0.2), dt_half=0.001, precision=precision),
                 g["pos0"], g["vel0"], np.full(n, 0.01))
    h.step(3)
    st = h.state()
    assert (st["id"] == g["id"]).all()
    assert np.abs(st["pos"] - g["pos"]).max() <= tol
    for f in ("vel", "e", "rho", "h"):
        assert U.rel_err(st[f], g[f], np.abs(g[f]).max() * 1e-3) <= tol, ("c3", f)
    h.close()
    g = _golden("c4_small")
    n = len(g["pos0"])
    h = L.Handle(L.make_params(hor=(0.0, 1.0), ver=(0.0, 1.0), dt_half=2e-3, precision=precision), g["pos0"], None, np.full(n, 0.01))
    h.step(2)
    st = h.state()
    assert np.abs(st["pos"] - g["pos"]).max() <= tol
    for f in ("e", "rho", "h"):
        assert U.rel_err(st[f], g[f], np.abs(g[f]).max() * 1e-3) <= tol, ("c4", f)
    # velocities started at zero: measured against the largest one
    assert np.abs(st["vel"] - g["vel"]).max() <= tol * np.abs(g["vel"]).max()
    h.close()


def test_split_upload_matches_the_serial_one():
    """sphb_upload_by_id_begin / _end: the copies run next to the step enqueued before them; same state as the serial call"""
    pos = gen.jittered_lattice(64, 64)
    n = len(pos)
    kw = dict(hor=(0.0, 1.0), ver=(0.0, 1.0), accel=(0.0, 0.2), dt_half=0.002)
    a = L.Handle(L.make_params(**kw), pos, None, np.full(n, 0.01))
    b = L.Handle(L.make_params(**kw), pos, None, np.full(n, 0.01))
    rng = np.random.default_rng(5)
    a.step(1); b.step(1)
    for k in range(3):
        st = a.state(["pos", "vel", "e", "id"])
        new = dict(pos=st["pos"] + 1e-4 * rng.standard_normal((n, 2)), vel=st["vel"] + 1e-3 * rng.standard_normal((n, 2)), e=st["e"] * 1.01)
        a.step(1); a.upload_by_id(**new); a.step(1)
        b.step(1)                      # asynchronous: still running when the copies start
        b.upload_by_id_begin(**new)
        with pytest.raises(L.SphbError):
            b.upload_by_id_begin(**new)  # one upload in flight per handle
        b.upload_by_id_end()
        b.step(1)
        sa, sb = a.state(["pos", "vel", "e", "rho", "h"]), b.state(["pos", "vel", "e", "rho", "h"])
        for f in ("pos", "vel", "e", "rho", "h"):
            assert np.array_equal(sa[f], sb[f]), (f, k)
    a.close(); b.close()


def test_device_copy_made_in_the_middle_of_a_run():
    """sphb_set_current_step: a handle created from the state of a simulation that has already stepped (what the Go shim
    does when simviewer replaces the Simulation value) continues it without repeating the step-0 initialisation"""
    pos = gen.jittered_lattice(48, 48)
    n = len(pos)
    kw = dict(hor=(0.0, 1.0), ver=(0.0, 1.0), accel=(0.0, 0.2), dt_half=0.002)
    a = L.Handle(L.make_params(**kw), pos, None, np.full(n, 0.01))
    a.step(2)
    st = a.state(["pos", "vel", "e", "rho", "vdot", "edot", "id"])
    b = L.Handle(L.make_params(**kw), st["pos"], st["vel"], st["e"], st["rho"], st["id"])
    b.upload(vdot=st["vdot"], edot=st["edot"])  # (device order = creation order = id order here)
    assert L.lib().sphb_set_current_step(b._h, 2) == 0 and b.current_step == 2
    a.step(1); b.step(1)
    sa, sb = a.state(["pos", "vel", "e", "rho", "h"]), b.state(["pos", "vel", "e", "rho", "h"])
    for f in ("pos", "vel", "e", "rho", "h"):
        assert U.rel_err(sb[f], sa[f], np.abs(sa[f]).max() * 1e-3) <= 1e-12, f
    # without the counter the copy would start over: VPred = Vel and a force evaluation before the step
    c = L.Handle(L.make_params(**kw), st["pos"], st["vel"], st["e"], st["rho"], st["id"])
    c.upload(vdot=st["vdot"], edot=st["edot"])
    c.step(1)
    assert U.rel_err(c.state(["vel"])["vel"], sa["vel"], np.abs(sa["vel"]).max() * 1e-3) > 1e-9
    a.close(); b.close(); c.close()
