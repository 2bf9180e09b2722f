"""CPU: host-side logic of the slab decomposition incl. a world_size-2 gloo exchange."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
from sphugo_b200 import slab  # noqa: E402


def test_topology_ring_and_open():
    t = slab.Topology(4, [0, .25, .5, .75, 1.0], True)
    assert [t.left(r) for r in range(4)] == [3, 0, 1, 2] and [t.right(r) for r in range(4)] == [1, 2, 3, 0]
    t = slab.Topology(3, [0, 1, 2, 3], False)
    assert t.left(0) is None and t.right(2) is None and t.left(1) == 0
    assert slab.Topology(1, [0, 1], True).left(0) is None  # a single slab uses the wrapping-grid path
    assert t.owner_of(np.array([0.5, 1.0, 2.999])).tolist() == [0, 1, 2]
    with pytest.raises(ValueError):
        slab.Topology(2, [0, 1], True)
    with pytest.raises(ValueError):
        slab.Topology(2, [0, 1, 1], True)


def test_equal_count_bounds():
    rng = np.random.default_rng(0)
    x = np.concatenate([rng.random(8000) * 0.5, 0.5 + rng.random(2000) * 0.5])  # 4:1 density like the shock tube
    b = slab.equal_count_bounds(x, 4, 0.0, 1.0)
    cnt = np.histogram(x, bins=b)[0]
    assert cnt.sum() == 10000 and cnt.max() - cnt.min() <= 2
    assert b[0] == 0.0 and b[-1] == 1.0 and all(b1 > b0 for b0, b1 in zip(b[:-1], b[1:]))


def test_ghost_widths_cover_the_stencil():
    gw, iw = slab.ghost_widths(0.01)
    assert iw >= 0.01 and gw >= iw + 0.01


def test_migration_schedule_bounds_the_excursion():
    """adaptive schedule: migrate before the accumulated movement bound plus the coming step can reach the slack"""
    h = 0.01
    sch = slab.MigrationSchedule(0, dt_half=0.001)
    # slow particles: 1e-3 * h per step -> hundreds of steps between migrations
    due = [sch.after_step(v_max=0.005, h_max=h) for _ in range(700)]
    assert 0 < sum(due) <= 2
    # the bound never exceeded the slack at any evaluation
    sch = slab.MigrationSchedule(0, dt_half=0.001)
    exc, vprev = 0.0, 0.0
    for k in range(200):
        v = 0.2 + 0.1 * (k % 7)
        exc += (vprev + v) * 0.001
        vprev = v
        if sch.after_step(v, h):
            exc = 0.0
        assert exc + 2 * v * 0.001 <= slab.GHOST_SLACK * h
    # fast flow (a third of h per step): every step
    sch = slab.MigrationSchedule(0, dt_half=0.001)
    assert all(sch.after_step(v_max=1.7, h_max=h) for _ in range(5))
    # fixed period
    sch = slab.MigrationSchedule(3, dt_half=0.001)
    assert [sch.after_step(1.0, h) for _ in range(6)] == [False, False, True, False, False, True]


def test_reference_halo_selection_periodic_frame():
    pos = np.array([[0.01, 0.5], [0.26, 0.5], [0.49, 0.5], [0.74, 0.5], [0.99, 0.5]])
    # slab [0, 0.25) of a ring: the particle at 0.99 is owned-looking from the frame (x - 1 = -0.01 < x_lo)
    assert slab.reference_halo(pos[:1], 0.0, 0.25, 0.05, 0, 1.0).tolist() == [0]
    assert slab.reference_halo(pos[4:], 0.75, 1.0, 0.05, 1, 1.0).tolist() == [0]
    # the criterion is one-sided (strays below the edge are sent too); the frame puts 0.99 at -0.01
    assert slab.reference_halo(pos, 0.25, 0.5, 0.05, 0, 1.0).tolist() == [0, 1, 4]


def _worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        topo = slab.Topology(world, [k / world for k in range(world + 1)], True)
        rng = np.random.default_rng(42)
        pos = rng.random((4000, 2))
        owner = topo.owner_of(pos[:, 0])
        mine = pos[owner == rank]
        x_lo, x_hi = topo.interval(rank)
        gw = 0.03
        ex = slab.DistExchange(topo, rank, torch.device("cpu"))
        packs = []
        for side in (0, 1):
            idx = slab.reference_halo(mine, x_lo, x_hi, gw, side, 1.0)
            buf = torch.zeros((len(idx) + 3, slab.HALO_DOUBLES), dtype=torch.float64)
            buf[: len(idx), :2] = torch.from_numpy(mine[idx])
            buf[: len(idx), 5] = float(rank)
            packs += [buf, len(idx)]
        rl, kl, rr, kr = ex.exchange(*packs, slab.HALO_DOUBLES)
        # expected: what the neighbours select towards me
        L, R = topo.left(rank), topo.right(rank)
        exp_l = pos[owner == L][slab.reference_halo(pos[owner == L], *topo.interval(L), gw, 1, 1.0)]
        exp_r = pos[owner == R][slab.reference_halo(pos[owner == R], *topo.interval(R), gw, 0, 1.0)]
        ok = (kl == len(exp_l) and kr == len(exp_r) and np.array_equal(rl[:kl, :2].numpy(), exp_l)
              and np.array_equal(rr[:kr, :2].numpy(), exp_r) and (rl[:kl, 5] == L).all() and (rr[:kr, 5] == R).all())
        hm = ex.allreduce_max(float(rank + 1))
        # the per-evaluation all-reduce of the slab driver: (max h, max speed) in one call
        h2, v2 = ex.allreduce_max(0.01 * (rank + 1), 5.0 - rank)
        ok = ok and h2 == 0.01 * world and v2 == 5.0
        q.put((rank, bool(ok), kl, kr, hm))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_gloo_neighbour_exchange(world):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + world + (os.getpid() % 200)
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert all(r[1] for r in res), res
    assert all(r[2] > 0 and r[3] > 0 for r in res)
    assert all(r[4] == float(world) for r in res)


def test_dist_slab_step_protocol_order_and_profile_ticks():
    """DistSlabSim.step over recording fakes: the per-evaluation call order of the protocol (slab.py header), migration when
    the schedule asks for it, and the per-phase wall-time ticks bench.py --gpus N reports (profile mode)"""
    calls = []

    class H:
        def sync(self):
            pass

    class FakeSlab:
        h = H()

        def __getattr__(self, name):
            def f(*a):
                calls.append(name)
                if name in ("pack_halo", "pack_migrants"):
                    return None, 0, None, 0
                if name == "local_max_h":
                    return 0.01
                if name == "local_max_speed":
                    return 1.0
            return f

    class FakeEx:
        def exchange(self, bl, nl, br, nr, width):
            calls.append(f"exchange{width}")
            return None, 0, None, 0

        def allreduce_max(self, *v):
            calls.append("allreduce")
            return list(v)

    sim = slab.DistSlabSim.__new__(slab.DistSlabSim)
    sim.slab, sim.ex, sim.h_max, sim.v_max, sim.steps_done, sim.h_growth = FakeSlab(), FakeEx(), 0.01, 0.0, 0, 1.0
    sim.schedule = slab.MigrationSchedule(2, 1e-3)  # migrate after every second step
    sim.profile, sim.prof = False, {}
    sim.step(1)
    ev = ["set_widths", "begin", "pack_halo", f"exchange{slab.HALO_DOUBLES}", "add_ghosts", "add_ghosts", "end", "local_max_h",
          "local_max_speed", "allreduce"]
    assert calls == ev + ev  # step 0: initialisation evaluation + the real one (sph.go:89-103)
    calls.clear()
    sim.profile = True
    sim.step(1)
    mig = ["pack_migrants", f"exchange{slab.MIGRANT_DOUBLES}", "add_migrants", "add_migrants", "finish_migration"]
    assert calls == ev + mig
    assert set(sim.prof) == {"begin", "pack_halo", "exchange", "add_ghosts", "end", "allreduce", "migrate"}
    assert all(v >= 0.0 for v in sim.prof.values()) and sim.schedule.migrations == 1


def test_spawned_particles_are_routed_to_the_owning_slab():
    """slab append (sources in a multi-GPU run): every particle goes to exactly one rank, x folded into the period"""
    topo = slab.Topology(3, [0.0, 0.3, 0.55, 1.0], True)
    pos = np.array([[0.1, 0.5], [0.3, 0.1], [0.54, 0.2], [0.99, 0.3], [1.02, 0.4], [-0.01, 0.6]])
    ids = np.arange(100, 106)
    e = np.linspace(1.0, 2.0, 6)
    got = [slab._route_to_owner(topo, r, pos, None, e, None, ids) for r in range(3)]
    assert [g[4].tolist() for g in got] == [[100, 104], [101, 102], [103, 105]]
    assert got[0][1] is None and np.array_equal(got[2][2], e[[3, 5]]) and np.array_equal(got[0][0], pos[[0, 4]])  # positions unchanged
    with pytest.raises(ValueError):
        slab._route_to_owner(topo, 0, pos, None, None, None, None)
