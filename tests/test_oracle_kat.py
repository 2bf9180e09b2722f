"""CPU: pins the oracle (oracle/sph_oracle.c) against every fixture the reference's own tests hold for the path
 and against independent exact computations.  Nothing here touches the CUDA library."""
import math

import numpy as np
import pytest

from oracle import oracle as orc
from sphugo_b200 import gen

VERT, HOR = 0, 1  # core.go:62-67

# sim/partition_test.go:11-16, 66-72, 128-134
PS_EVEN = [(1.0, 1.0), (0.9, 0.9), (0.8, 0.8), (0.7, 0.7)]
PS_ODD = [(0.0, 0.9), (0.5, -0.8), (1.7, 0.1), (0.7, -0.1), (-0.7, 0.1)]
PS_VAR = [(0.9, 0.0), (-0.8, 0.5), (0.1, 1.7), (-0.1, 0.7), (0.1, -0.7)]

PARTITION_KATS = [  # (fixture, orientation, pivot, len(a), len(b))  partition_test.go:18-160
    (PS_EVEN, VERT, 0.5, 0, 4), (PS_EVEN, HOR, 0.5, 0, 4), (PS_EVEN, VERT, 0.85, 2, 2), (PS_EVEN, HOR, 0.85, 2, 2),
    (PS_ODD, VERT, 0.100000000001, 4, 1), (PS_ODD, HOR, 0.601, 3, 2), (PS_ODD, VERT, -100, 0, 5), (PS_ODD, HOR, 100, 5, 0),
    (PS_VAR, HOR, 0.100000000001, 4, 1), (PS_VAR, VERT, 0.601, 3, 2), (PS_VAR, HOR, -100, 0, 5), (PS_VAR, VERT, 100, 5, 0),
    ([], HOR, 0.85, 0, 0),
]


@pytest.mark.parametrize("pts,ori,pivot,la,lb", PARTITION_KATS)
def test_partition_kats(pts, ori, pivot, la, lb):
    a, b, out = orc.partition(pts, ori, pivot)
    assert (a, b) == (la, lb)
    # beyond the reference's length checks: a permutation, split by the pivot on the right coordinate
    assert sorted(map(tuple, out.tolist())) == sorted(map(tuple, np.array(pts, float).reshape(-1, 2).tolist()))
    col = 1 if ori == VERT else 0
    assert (out[:a, col] <= pivot).all() and (out[a:, col] > pivot).all()


@pytest.mark.parametrize("n", [1, 60, 6000])
@pytest.mark.parametrize("stream", ["go", "splitmix"])
def test_inside_any_sphere(n, stream):
    """bounding-sphere_test.go:30-64: every particle lies inside some node circle.  "go": the test's own particles,
    MakeCellsUniform -> InitUniformly (core.go:76-91, 106-113) on Go's math/rand stream after rand.Seed(12345678)"""
    from sphugo_b200 import gorand
    pos = gorand.init_uniformly(n)["pos"] if stream == "go" else gen.uniform_rect(n)
    o = orc.Oracle(orc.make_params(), pos)
    assert o.outside_all_circles() == 0
    st = o.tree_stats()
    assert st["max_leaf"] <= 8  # MAX_PARTICLES_PER_CELL, core.go:11
    o.close()


def test_heap_readme_kat():
    """README.md:88-162"""
    init = [31, 37, 82, 83, 33, 54, 39, 42, 62, 49, 84, 59, 88, 26, 27, 21, 92, 97, 87, 49, 33, 9, 42, 49, 88, 67]
    heap = [9, 21, 26, 31, 33, 49, 27, 42, 62, 37, 33, 54, 67, 39, 82, 83, 92, 97, 87, 49, 49, 84, 42, 59, 88, 88]
    assert orc.heap_build(init) == heap
    ins = orc.heap_insert(heap, 0)
    assert ins == [0, 21, 9, 31, 33, 26, 27, 42, 62, 37, 33, 54, 49, 39, 82, 83, 92, 97, 87, 49, 49, 84, 42, 59, 88, 88, 67]
    ext, m = orc.heap_extract_min(ins)
    assert m == 0 and ext == heap
    rep, m = orc.heap_replace(heap, 33)
    assert m == 9
    assert rep == [21, 31, 26, 33, 33, 49, 27, 42, 62, 37, 33, 54, 67, 39, 82, 83, 92, 97, 87, 49, 49, 84, 42, 59, 88, 88]


def test_particle_layout_and_constants():
    assert orc.lib().orc_sizeof_particle() == 1136  # core.go:17-42 (SURVEY header table)
    # Go untyped-constant expressions rounded once (sph.go:249,261,275,293,303)
    from fractions import Fraction
    pi = Fraction("3.14159265358979323846264338327950288419716939937510582097494459")
    assert float(Fraction(240) / (pi * 7)) == float.fromhex("0x1.5d3b3e3583243p+3")
    assert float(Fraction(28) / (pi * 4)) == float.fromhex("0x1.1d34a60108f72p+1")
    assert float(Fraction(56) / (pi * 4)) == float.fromhex("0x1.1d34a60108f72p+2")
    assert float(1 / pi) == float.fromhex("0x1.45f306dc9c883p-2")


def _exact_periodic_knn(pos, box):
    from scipy.spatial import cKDTree
    t = cKDTree(np.mod(pos, box), boxsize=box)
    d, j = t.query(np.mod(pos, box), k=33)
    return d[:, 1:], j[:, 1:]


def test_faithful_knn_equals_exact_and_ckdtree_periodic():
    ic = gen.spawn([(1000, (0.0, 0.0), (1.0, 1.0)), (200, (0.1, 0.0), (0.3, 0.4))])
    p = orc.make_params(hor=(0, 1), ver=(0, 1))
    a, b = orc.Oracle(p, ic["pos"]), orc.Oracle(p, ic["pos"])
    a.knn(mode=0); b.knn(mode=1)
    sa, sb = a.state(neighbours=True), b.state(neighbours=True)
    assert (np.sort(sa["nn_id"], 1) == np.sort(sb["nn_id"], 1)).all()
    assert np.array_equal(sa["nn_dist"], sb["nn_dist"])
    assert a.underfull == 0
    d, j = _exact_periodic_knn(ic["pos"], 1.0)
    assert (np.sort(sa["nn_id"], 1) == np.sort(j, 1)).all()
    assert np.allclose(sa["h"], d[:, -1], rtol=1e-12, atol=0)
    # list order: descending distance, slot 0 = h (nearest-neighbour.go:139-153)
    assert (np.diff(sa["nn_dist"], axis=1) <= 0).all() and np.array_equal(sa["nn_dist"][:, 0], sa["h"])


def test_open_boundary_knn_matches_bruteforce():
    pos = gen.uniform_rect(700, seed=3)
    o = orc.Oracle(orc.make_params(), pos)
    o.knn(mode=0)
    s = o.state(neighbours=True)
    d2 = ((pos[:, None, :] - pos[None, :, :]) ** 2).sum(-1)
    np.fill_diagonal(d2, np.inf)
    assert np.allclose(np.sort(np.sqrt(np.sort(d2, 1)[:, :32]), 1), np.sort(s["nn_dist"], 1), rtol=1e-14)


def test_half_open_axis_panics():
    o = orc.Oracle(orc.make_params(), gen.uniform_rect(100))
    with pytest.raises(RuntimeError):
        o.knn(hor=(orc.OPEN[0], 1.0), ver=orc.OPEN)  # nearest-neighbour.go:44


def test_density_on_a_lattice_is_uniform_and_kernels_are_normalised():
    """analytic check: on a perfect periodic lattice every particle has the same density, and the kernel
    prefactors integrate to one: int_0^1 F(q) 2 pi q dq * pref = 1 (q = r / h)."""
    q = (np.arange(200000) + 0.5) / 200000
    mon = np.where(q < 0.5, q ** 3 - q ** 2 + 1 / 6, (1 - q) ** 3 / 3)
    wen = (1 - q) ** 4 * (1 + 4 * q)
    assert abs(np.sum(mon * 2 * np.pi * q) / 200000 * 6 * 40 / (7 * np.pi) - 1) < 1e-6
    assert abs(np.sum(wen * 2 * np.pi * q) / 200000 * 4 * 7 / (4 * np.pi) - 1) < 1e-6
    pos = gen.jittered_lattice(24, 24, jitter=0.0)
    o = orc.Oracle(orc.make_params(hor=(0, 1), ver=(0, 1)), pos)
    o.knn(mode=1)
    for k in (0, 1, 2):
        o.density(k)
        rho = o.state()["rho"]
        assert np.ptp(rho) <= 1e-9 * rho.mean()


def test_step_quirks():
    """leapfrog order, wrap `continue` quirk and reflections"""
    rng = np.random.default_rng(1)
    pos = rng.random((300, 2))
    vel = (rng.random((300, 2)) - 0.5) * 40
    kw = dict(hor=(0.0, 1.0), ver=(0.0, 1.0), dt_half=0.01)
    o = orc.Oracle(orc.make_params(**kw), pos, vel, np.full(300, 0.01))
    o.step(1)
    s = o.state()
    # a particle that left through a corner was shifted in X only (the `continue`, sph.go:147-167)
    assert ((s["pos"][:, 1] < 0) | (s["pos"][:, 1] > 1)).any()
    # reflections: pos == wall exactly and the velocity component flipped (sph.go:170-193)
    kw = dict(refl=(0.2, 0.8, 0.2, 0.8), dt_half=0.01)
    o = orc.Oracle(orc.make_params(**kw), 0.25 + 0.5 * pos, vel, np.full(300, 0.01))
    o.step(1)
    s = o.state()
    # `pos -= pos - wall` is not `pos = wall`: it may land one ulp beside the wall
    assert s["pos"].min() >= 0.2 - 1e-15 and s["pos"].max() <= 0.8 + 1e-15
    assert (np.isclose(s["pos"], 0.2) | np.isclose(s["pos"], 0.8)).any()
    assert (s["pos"] != 0.2).all() or (s["pos"] == 0.2).any()
    assert o.current_step == 1


def test_total_momentum_bug_is_kept():
    pos = gen.uniform_rect(64)
    vel = np.arange(128, dtype=float).reshape(64, 2)
    o = orc.Oracle(orc.make_params(), pos, vel)
    s = o.state(sort_by_id=False)
    assert o.total_momentum() == pytest.approx(math.hypot(*s["vel"][-1]))  # `=` not `+=`, sph.go:460


GOLDEN = ["c1_density", "c2_default", "c2_example_config"]


def test_golden_vectors_reproduce():
    """tests/golden/*.npz are what the oracle produces today (make_golden.py); a change of the oracle shows here"""
    import os
    from tests.golden import make_golden as mg  # noqa: F401  (importable = the generating script is committed)
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "c2_default.npz"))
    o = orc.Oracle(orc.make_params(), g["pos0"], None, np.full(len(g["pos0"]), 0.01), None, g["id"])
    o.step(1)
    s = o.state()
    for f in ("pos", "vel", "e", "rho", "h", "vdot", "edot"):
        assert np.array_equal(s[f], g[f + "_1"]), f
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "c1_density.npz"))
    o = orc.Oracle(orc.make_params(hor=(0, 1), ver=(0, 1)), g["pos"], ids=g["id"])
    o.knn(mode=1)  # the exact mode reproduces the faithful golden list
    s = o.state(neighbours=True)
    assert (np.sort(s["nn_id"], 1) == g["nn_id"]).all() and np.array_equal(s["h"], g["h"])


def test_golden_c3_c4_small_reproduce():
    """the reduced-size C3 / C4 fixtures are what the oracle (exact-kNN mode) produces today"""
    import os
    from sphugo_b200 import gen
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "c4_small.npz"))
    pos = gen.shock_tube(8000)
    assert np.array_equal(pos, g["pos0"])
    o = orc.Oracle(orc.make_params(hor=(0.0, 1.0), ver=(0.0, 1.0), dt_half=2e-3), pos, None, np.full(len(pos), 0.01))
    o.step(2, 1)
    st = o.state()
    for f in ("pos", "vel", "e", "rho", "h"):
        assert np.array_equal(st[f], g[f]), f
    o.close()
