"""-m gpu: the slab ring inside the library (sphb_ring_set / sphb_ring_step_local, sphb_ring.inc) against the oracle
and the single handle.  All slabs live in one process on one GPU; the exchange between them is a device copy, everything
else - ghosts kept in the sorted arrays across reuse evaluations, fixed-size halos, migration at rebuilds, the schedule -
is the code a multi-GPU run executes (tests/test_multi_gpu.py runs the same over NCCL where two GPUs exist)."""
import numpy as np
import pytest

from tests import util as U
from tests.test_gpu_reuse import env
from tests.test_gpu_slab import FIELDS, _compare
from oracle import oracle as orc
from sphugo_b200 import _lib as L
from sphugo_b200 import gen, slab

pytestmark = pytest.mark.gpu


def _run(ic, world, steps, periodic, period, bounds=None, tol=1e-12, precision=64, **cfg):
    po, pg = U.params_pair(**cfg)
    pg.precision = precision
    lo, hi = (cfg["hor"] if periodic else (float(ic["pos"][:, 0].min()), float(ic["pos"][:, 0].max()) + 1e-9))
    bounds = bounds or slab.equal_count_bounds(ic["pos"][:, 0], world, lo, hi)
    topo = slab.Topology(world, bounds, periodic)
    n = len(ic["pos"])
    area = (hi - lo) * (ic["pos"][:, 1].max() - ic["pos"][:, 1].min())
    with env(SPHB_REUSE_PERIOD=period):
        sim = slab.LocalRingSim(pg, topo, ic["pos"], ic.get("vel"), ic.get("e"), ic["id"],
                                h_max_hint=slab.default_h_hint(n, area))
    o = orc.Oracle(po, ic["pos"], ic.get("vel"), ic.get("e"), None, ic["id"])
    for k in range(steps):
        sim.step(1)
        o.step(1, knn_mode=1)
        got, ref = sim.state(FIELDS), o.state(neighbours=True)
        assert sum(sim.counts()) == n
        _compare(got, ref, po, tol if k == 0 else 1e-9 if precision == 64 else 1e-4, f"world={world} step {k + 1}")
    reuse = [h.counters()["reuse_steps"] for h in sim.handles]
    info = sim.info()
    sim.close(); o.close()
    return reuse, info


def _lattice_ic(nx, ny, v):
    pos = gen.jittered_lattice(nx, ny)
    n = len(pos)
    return dict(pos=pos, vel=np.tile([v], (n, 1)), e=np.full(n, 0.01), id=np.arange(n, dtype=np.int64))


def test_ring_without_reuse_matches_the_oracle():
    _run(_lattice_ic(96, 96, [2.0, -1.0]), 2, 4, True, 1, hor=(0.0, 1.0), ver=(0.0, 1.0), accel=(0.0, 0.2), dt_half=0.002)


def test_two_slab_ring_with_reuse_cycles():
    """ghosts stay in the sorted arrays for the reuse evaluations; halos of fixed size; every step against the oracle"""
    reuse, _ = _run(_lattice_ic(96, 96, [0.5, -0.3]), 2, 9, True, 4, hor=(0.0, 1.0), ver=(0.0, 1.0), accel=(0.0, 0.2),
                    dt_half=0.004)
    assert min(reuse) >= 5


def test_four_slab_ring_bulk_flow_migration_and_reuse():
    """a bulk flow carries particles across slab edges and the periodic seam: migration happens at rebuilds only"""
    reuse, info = _run(_lattice_ic(128, 64, [4.0, 1.0]), 4, 10, True, 3, hor=(0.0, 1.0), ver=(0.0, 1.0), dt_half=0.002)
    assert min(reuse) >= 5 and info[0]["migrations"] >= 1


def test_three_slabs_open_box_with_reuse():
    ic = gen.spawn([(6000, (0, 0), (1, 1))])
    _run(ic, 3, 6, False, 3, accel=(0.0, 0.2), dt_half=0.002)


def test_ring_fp32_build_with_reuse():
    _run(_lattice_ic(96, 96, [1.5, -0.7]), 3, 5, True, 3, bounds=[0.0, 0.35, 0.7, 1.0], tol=1e-5, precision=32,
         hor=(0.0, 1.0), ver=(0.0, 1.0), accel=(0.0, 0.2), dt_half=0.002)


def test_ring_default_schedule_matches_single_handle():
    pos = gen.jittered_lattice(128, 128)
    n = len(pos)
    kw = dict(hor=(0.0, 1.0), ver=(0.0, 1.0), accel=(0.0, 0.2), dt_half=0.002)
    pg = L.make_params(**kw)
    g = L.Handle(pg, pos, None, np.full(n, 0.01))
    with env(SPHB_REUSE=1):
        sim = slab.LocalRingSim(pg, slab.Topology(2, [0.0, 0.5, 1.0], True), pos, None, np.full(n, 0.01),
                                h_max_hint=slab.default_h_hint(n, 1.0))
    g.step(9); sim.step(9)
    a, b = g.state(FIELDS), sim.state(FIELDS)
    assert (a["id"] == b["id"]).all()
    assert U.rel_err(b["h"], a["h"]) <= 1e-11 and np.abs(a["pos"] - b["pos"]).max() <= 1e-11
    assert U.rel_err(b["rho"], a["rho"]) <= 1e-10
    g.close(); sim.close()
