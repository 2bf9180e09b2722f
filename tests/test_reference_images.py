"""The reference's own recorded output pins the hot path: doc/density_compare.png, doc/density_test.png and
doc/density_test_periodic.png are what the reference's Go binary drew for examples/density (README.md:201-206) - periodic and
open kNN (k = 32) and Density2D with TopHat / Monaghan / Wendland on BASELINE configs[0].  Here the same scenes are built
from the reconstructed Go math/rand stream, evaluated, drawn with a restatement of the example's drawing code
(tests/gx_restatement.py) and compared PIXEL FOR PIXEL (SHA-256 of the RGB array, tests/golden/reference_images.json; against
the PNG itself where /root/reference is present).  Every disk's colour is a 255-level quantisation of one particle's density
and disks overdraw each other in Root.Particles order, so a match pins the particle stream, the tree permutation, the
neighbour search and the three density kernels of the oracle - and, in the GPU twin of this test
(tests/test_zz_gpu_go_scenes.py::test_cuda_path_reproduces_the_reference_pictures), of the CUDA path."""
import json
import os

import numpy as np

from oracle import oracle as orc
from sphugo_b200 import gorand
from tests import gx_restatement as gx

HERE = os.path.dirname(os.path.abspath(__file__))
with open(os.path.join(HERE, "golden", "reference_images.json")) as _f:
    REF = json.load(_f)
SIDE = 420


def check(name, canvas, look_at_the_png=True):
    r = REF[name]
    assert (canvas.W, canvas.H) == (r["width"], r["height"])
    png = os.path.join("/root/reference", r["source"])
    if look_at_the_png and os.path.exists(png):  # CPU tests in the build container: say how far the pictures differ
        from PIL import Image
        ref = np.array(Image.open(png).convert("RGB"))
        bad = int((ref != canvas.img).any(2).sum())
        assert bad == 0, f"{name}: {bad} of {ref.shape[0] * ref.shape[1]} pixels differ from the reference's picture"
    assert canvas.sha256() == r["sha256_rgb"], name


def density_scene():
    """examples/density main (density.go:50-63): 1000 + 200 particles, each spawner re-seeding Go's math/rand"""
    a, b = gorand.uniform_rect_spawn(1000), gorand.uniform_rect_spawn(200, (0.1, 0.0), (0.3, 0.4))
    return np.concatenate([a["pos"], b["pos"]])


def periodic_visual_scene():
    """periodicVisualTest (density.go:103-118): 10000 + 1200 particles"""
    a, b = gorand.uniform_rect_spawn(10000, (0.1, 0.1), (0.9, 0.9)), gorand.uniform_rect_spawn(1200, (0.85, 0.4), (0.9, 0.9))
    return np.concatenate([a["pos"], b["pos"]])


PANELS = [(0, gx.HeatRamp), (1, gx.HeatRamp), (2, gx.HeatRamp), (0, gx.ParaRamp), (1, gx.ParaRamp), (2, gx.ParaRamp)]  # density.go:81-87


def draw_density_compare(order_pos, rho_of_kernel):
    """density.go:74-94: six panels (three kernels x two ramps) and the white separator lines"""
    w, h = 3 * SIDE, 2 * SIDE
    c = gx.Canvas(w, h)
    for k, (kernel, ramp) in enumerate(PANELS):
        gx.draw_density_panel(c, order_pos, rho_of_kernel[kernel], SIDE, k, ramp)
    c.DrawLine((SIDE, 0), (SIDE, h), gx.WHITE)
    c.DrawLine((2 * SIDE, 0), (2 * SIDE, h), gx.WHITE)
    c.DrawLine((0, SIDE), (w, SIDE), gx.WHITE)
    return c


def test_oracle_reproduces_density_compare_png():
    pos = density_scene()
    o = orc.Oracle(orc.make_params(hor=(0, 1), ver=(0, 1)), pos)  # MakeCells
    o.knn((0.0, 1.0), (0.0, 1.0), mode=0)  # Treebuild + BoundingSpheres + the per-particle periodic search (density.go:65-72)
    rho = {}
    for kernel in (0, 1, 2):
        o.density(kernel)
        st = o.state(sort_by_id=False)  # Root.Particles order: the drawing order
        rho[kernel] = st["rho"]
    check("density_compare", draw_density_compare(st["pos"], rho))
    o.close()


def test_oracle_reproduces_density_test_pngs():
    pos = periodic_visual_scene()
    o = orc.Oracle(orc.make_params(), pos)
    # the example re-makes the tree on the slice the first pass permuted (density.go:122-125): one oracle, two passes
    for name, hor, ver in (("density_test", orc.OPEN, orc.OPEN), ("density_test_periodic", (0.1, 0.9), (0.1, 0.9))):
        o.knn(hor, ver, mode=0)
        o.density(0)
        st = o.state(sort_by_id=False)
        c = gx.Canvas(700, 350)
        gx.draw_density_test(c, st["pos"], st["rho"])
        check(name, c)
    o.close()


def test_a_wrong_density_would_show():
    """sensitivity of the comparison: 0.2 % on the densities changes the picture"""
    pos = density_scene()
    o = orc.Oracle(orc.make_params(hor=(0, 1), ver=(0, 1)), pos)
    o.knn((0.0, 1.0), (0.0, 1.0), mode=0)
    rho = {}
    for kernel in (0, 1, 2):
        o.density(kernel)
        st = o.state(sort_by_id=False)
        rho[kernel] = st["rho"] * 1.002
    assert draw_density_compare(st["pos"], rho).sha256() != REF["density_compare"]["sha256_rgb"]
    o.close()


def test_oracle_reproduces_tree_png():
    """examples/tree-partition (tree-partition.go:33-47): InitUniformly(2200) on the Go stream, Treebuild, MakeTreePlot.  Every
    cell rectangle (coloured by level) and every particle pixel matches: pins Partition / Treebuild (core.go:126-224)"""
    ic = gorand.init_uniformly(2200)
    o = orc.Oracle(orc.make_params(), ic["pos"])
    check("tree", gx.MakeTreePlot(gx.Tree(*o.tree()), o.state(sort_by_id=False)["pos"], 1024, 1024))
    o.close()


def _nearest_neighbours_picture(periodic, batch_knn=None):
    """examples/nearest-neighbors NonPeriodic / Periodic (nearest-neighbors.go:20-102): 220 particles, the tree's cells (open)
    or the bounding circles of its leaves (periodic), all particles, then particle 14 of the tree order, its 32 neighbours
    and the circle of radius NNDists[0] (periodic: in all nine images).  The example queries a COPY of the particle, and
    self-exclusion is by address (nearest-neighbour.go:79): the original is found at distance 0 and takes a slot, so the
    list is the original + the 31 nearest others and NNDists[0] is the 31st.
    batch_knn(pos, hor, ver, particle_id) -> (neighbour ids, distances), both by descending distance: the source of the
    neighbour list (default: the oracle; the GPU twin of this test passes the CUDA path's)"""
    f32 = gx.f32
    ic = gorand.init_uniformly(220)
    o = orc.Oracle(orc.make_params(), ic["pos"])  # MakeCellsUniform; BoundingSpheres again changes nothing
    tree = gx.Tree(*o.tree())
    w = h = 1000
    c = gx.Canvas(w, h)
    # the pictures were drawn by a gx.DrawCircle whose scan box covered radius + border; today's clips the ring at its four
    # extreme points (130 pixels of either picture), everything else is identical
    if periodic:
        tree.PlotBoundingCircles(c, 0, 1, gx.WHITE, box_with_border=True)
    else:
        tree.PlotCells(c, 0, 1, 1)
    hv = ((0.0, 1.0), (0.0, 1.0)) if periodic else (orc.OPEN, orc.OPEN)
    o.knn(hv[0], hv[1], mode=0, rebuild=False)
    st = o.state(neighbours=True, sort_by_id=False)
    for p in st["pos"]:
        c.DrawDisk(f32(p[0] * float(w)), f32(p[1] * float(h)), 3.4, gx.ORANGE)
    p0 = st["pos"][14]
    x, y = p0[0] * float(w), p0[1] * float(h)
    c.DrawDisk(f32(x), f32(y), 10, gx.GREEN)
    by_id = {int(i): p for i, p in zip(st["id"], st["pos"])}
    nn_id, nn_dist = (st["nn_id"][14], st["nn_dist"][14]) if batch_knn is None else batch_knn(ic["pos"], hv[0], hv[1], int(st["id"][14]))
    for i in nn_id[1:]:  # descending distance; slot 0 of the batch result (the 32nd other) is not in the copy's list
        pn = by_id[int(i)]
        c.DrawDisk(f32(pn[0] * float(w)), f32(pn[1] * float(h)), 4.4, gx.GREEN)
    c.DrawDisk(f32(x), f32(y), 4.4, gx.GREEN)  # the original itself, last (distance 0)
    radius = f32(nn_dist[1] * float(w))
    for i in ((-1.0, 0.0, 1.0) if periodic else (0.0,)):
        for j in ((-1.0, 0.0, 1.0) if periodic else (0.0,)):
            c.DrawCircle(f32(x) + f32(float(w) * i), f32(y) + f32(float(h) * j), radius, 2, gx.GREEN, box_with_border=True)
    o.close()
    return c


def nearest_neighbours_pictures(batch_knn=None, look_at_the_png=True):
    check("nearest_neighbours", _nearest_neighbours_picture(False, batch_knn), look_at_the_png)
    check("nearest_neighbours_periodic", _nearest_neighbours_picture(True, batch_knn), look_at_the_png)


def test_oracle_reproduces_nearest_neighbours_pngs():
    """pins, against the Go program's pictures: the leaf bounding circles (core.go:229-298), the open and the periodic
    neighbour search of a particle next to the box edge (its neighbours come through the periodic images) and h"""
    nearest_neighbours_pictures()


def test_colour_ramps_reproduce_custom_colour_maps_png():
    """examples/color-ramp (color-ramp.go:7-23): the four ramps the other pictures are coloured with"""
    c = gx.Canvas(80, 256)
    for j, cmap in enumerate([gx.RainbowRamp, gx.ParaRamp, gx.HeatRamp, gx.ToxicRamp]):
        for i in range(256):
            c.DrawRect((j * 20, 255 - i), ((j + 1) * 20, 255 - i), cmap(i))
    check("customColorMaps", c)
