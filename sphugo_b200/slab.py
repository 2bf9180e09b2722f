"""Slab decomposition driver: one process / one libsphb handle per GPU, 1-D slabs along x.

The numeric work is all inside libsphb.so (sphb_slab_* entry points, include/sphb.h); this module is the
plumbing the north_star assigns to `torch.distributed`: who is my neighbour, how wide is the ghost layer,
and the NCCL send/recv of the packed device buffers over NVLink.  Per force evaluation

    begin (drift-1 + predict)  ->  pack halo x2  ->  exchange  ->  add ghosts x2  ->  end (sort, kNN, density,
    force, kick, drift-2, boundaries; ghosts dropped)  ->  [migrate: pack x2 -> exchange -> add x2 -> compact]

There is exactly one exchange per evaluation: the ghost layer is 2 x (safety x max h) wide and the inner half of
it is evaluated redundantly, so neighbours' rho / c / h never have to travel (no "halo-2").

Two exchange back-ends with the same interface:
  DistExchange   torch.distributed (nccl on GPUs; gloo with CPU tensors in the CPU tests)
  LocalExchange  several slabs inside one process (single-GPU tests of the whole slab path)
"""
from __future__ import annotations

import ctypes as C
import math
from dataclasses import dataclass
from typing import List, Optional, Sequence, Tuple

import numpy as np

HALO_DOUBLES = 10
MIGRANT_DOUBLES = 12
OPEN_LO = -1.7976931348623157e308


# ------------------------------------------------------------------------------------------------
# topology (pure Python: covered by the CPU tests)
# ------------------------------------------------------------------------------------------------
@dataclass
class Topology:
    """world slabs along x with edges `bounds` (len world + 1); periodic = the x axis wraps (ring)."""

    world: int
    bounds: Sequence[float]
    periodic: bool

    def __post_init__(self):
        if len(self.bounds) != self.world + 1:
            raise ValueError("bounds must have world + 1 entries")
        if any(b1 <= b0 for b0, b1 in zip(self.bounds[:-1], self.bounds[1:])):
            raise ValueError("slab edges must increase")

    def left(self, rank: int) -> Optional[int]:
        if self.world == 1:
            return None  # a single slab uses the ordinary (wrapping grid) path
        if rank > 0:
            return rank - 1
        return self.world - 1 if self.periodic else None

    def right(self, rank: int) -> Optional[int]:
        if self.world == 1:
            return None
        if rank < self.world - 1:
            return rank + 1
        return 0 if self.periodic else None

    def interval(self, rank: int) -> Tuple[float, float]:
        return float(self.bounds[rank]), float(self.bounds[rank + 1])

    def owner_of(self, x: np.ndarray) -> np.ndarray:
        """rank owning coordinate x (x already inside [bounds[0], bounds[-1]) for a periodic axis)"""
        r = np.searchsorted(np.asarray(self.bounds[1:-1], dtype=np.float64), x, side="right")
        return r.astype(np.int64)


def equal_count_bounds(x: np.ndarray, world: int, lo: float, hi: float) -> List[float]:
    """slab edges with equal particle counts (quantiles of x), outer edges fixed at lo / hi"""
    if world == 1:
        return [lo, hi]
    qs = np.quantile(np.asarray(x, dtype=np.float64), [k / world for k in range(1, world)])
    return [lo] + [float(q) for q in qs] + [hi]


def ghost_widths(h_max: float, safety: float = 1.15, slack: float = 0.5) -> Tuple[float, float]:
    """(ghost_w, inner_w).  An owned particle may sit up to slack * h_max outside its nominal slab (drift-1 of the
    current step; migration restores the interval after every step), so every neighbour of an owned particle lies
    within inner_w = (safety + slack) * h_max of the slab: those ghosts are evaluated.  They in turn need their own
    neighbours, another safety * h_max further out.  Both conditions are verified on the device
    (SPHB_E_GHOST_THIN)."""
    inner = (safety + slack) * h_max
    return inner + safety * h_max, inner


def reference_halo(pos: np.ndarray, x_lo: float, x_hi: float, ghost_w: float, side: int, period: float = 0.0):
    """numpy restatement of k_pack_halo's selection (host-logic tests): indices of owned particles that must be
    sent to the low (side 0) / high (side 1) neighbour"""
    x = np.asarray(pos, dtype=np.float64)[:, 0]
    if period > 0.0:
        mid = 0.5 * (x_lo + x_hi)
        x = x - period * np.round((x - mid) / period)
    return np.nonzero((x - x_lo < ghost_w) if side == 0 else (x_hi - x <= ghost_w))[0]


# ------------------------------------------------------------------------------------------------
# exchange back-ends
# ------------------------------------------------------------------------------------------------
class DistExchange:
    """neighbour exchange over torch.distributed: counts first, then payloads of the exact size."""

    def __init__(self, topo: Topology, rank: int, device):
        import torch
        import torch.distributed as dist
        self.torch, self.dist, self.topo, self.rank, self.device = torch, dist, topo, rank, device

    def _batch(self, ops):
        if ops:
            for r in self.dist.batch_isend_irecv(ops):
                r.wait()

    def exchange(self, send_left, n_left: int, send_right, n_right: int, width: int):
        """send_* : [cap, width] tensors on self.device holding n_* records.  Returns (recv_left, k_left,
        recv_right, k_right): what the left / right neighbour sent towards me."""
        torch, dist = self.torch, self.dist
        L, R = self.topo.left(self.rank), self.topo.right(self.rank)
        cnt_out = torch.tensor([n_left, n_right], dtype=torch.int64, device=self.device)
        cnt_in = torch.zeros(2, dtype=torch.int64, device=self.device)
        ops = []
        # what I pack for side 0 goes to my left neighbour; my right neighbour's side-0 pack comes to me
        if L is not None:
            ops += [dist.P2POp(dist.isend, cnt_out[0:1], L), dist.P2POp(dist.irecv, cnt_in[0:1], L)]
        if R is not None:
            ops += [dist.P2POp(dist.isend, cnt_out[1:2], R), dist.P2POp(dist.irecv, cnt_in[1:2], R)]
        if L is not None and L == R:  # two slabs on a ring: both neighbours are the same peer; order by direction
            ops = [dist.P2POp(dist.isend, cnt_out[0:1], L), dist.P2POp(dist.isend, cnt_out[1:2], L),
                   dist.P2POp(dist.irecv, cnt_in[1:2], L), dist.P2POp(dist.irecv, cnt_in[0:1], L)]
        self._batch(ops)
        k_left, k_right = (int(v) for v in cnt_in.tolist())
        recv_left = torch.empty((max(k_left, 1), width), dtype=torch.float64, device=self.device)
        recv_right = torch.empty((max(k_right, 1), width), dtype=torch.float64, device=self.device)
        ops = []
        if L is not None and L == R:
            if n_left:
                ops.append(dist.P2POp(dist.isend, send_left[:n_left], L))
            if n_right:
                ops.append(dist.P2POp(dist.isend, send_right[:n_right], L))
            if k_right:  # the peer's side-0 pack (sent first) lands on my right
                ops.append(dist.P2POp(dist.irecv, recv_right[:k_right], L))
            if k_left:
                ops.append(dist.P2POp(dist.irecv, recv_left[:k_left], L))
        else:
            if L is not None:
                if n_left:
                    ops.append(dist.P2POp(dist.isend, send_left[:n_left], L))
                if k_left:
                    ops.append(dist.P2POp(dist.irecv, recv_left[:k_left], L))
            if R is not None:
                if n_right:
                    ops.append(dist.P2POp(dist.isend, send_right[:n_right], R))
                if k_right:
                    ops.append(dist.P2POp(dist.irecv, recv_right[:k_right], R))
        self._batch(ops)
        if self.device is not None and getattr(self.device, "type", "cpu") == "cuda":
            torch.cuda.current_stream(self.device).synchronize()  # libsphb reads the buffers on its own stream
        return recv_left, k_left, recv_right, k_right

    def allreduce_max(self, *vals: float):
        t = self.torch.tensor(list(vals), dtype=self.torch.float64, device=self.device)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        out = [float(v) for v in t.tolist()]
        return out[0] if len(out) == 1 else out


class LocalExchange:
    """all slabs live in this process: `post` collects every rank's packs, `collect` hands them to the neighbours"""

    def __init__(self, topo: Topology):
        self.topo = topo
        self.box = {}

    def post(self, rank, send_left, n_left, send_right, n_right):
        self.box[rank] = (send_left, n_left, send_right, n_right)

    def collect(self, rank):
        L, R = self.topo.left(rank), self.topo.right(rank)
        recv_left, k_left, recv_right, k_right = None, 0, None, 0
        if L is not None:  # the left neighbour's side-1 (high edge) pack
            recv_left, k_left = self.box[L][2], self.box[L][3]
        if R is not None:  # the right neighbour's side-0 pack
            recv_right, k_right = self.box[R][0], self.box[R][1]
        return recv_left, k_left, recv_right, k_right


# ------------------------------------------------------------------------------------------------
# one slab = one libsphb handle
# ------------------------------------------------------------------------------------------------
class Slab:
    """phase-wise wrapper of the sphb_slab_* entry points for one rank"""

    def __init__(self, params, topo: Topology, rank: int, pos, vel=None, e=None, ids=None, capacity=None,
                 halo_cap=None, h_max_hint=None, safety=1.15):
        import torch
        from . import _lib as L
        self.L, self.torch, self.topo, self.rank, self.safety = L, torch, topo, rank, safety
        pos = np.ascontiguousarray(pos, dtype=np.float64).reshape(-1, 2)
        n = pos.shape[0]
        self.capacity = int(capacity or (n + max(4096, n // 4)))
        self.h = L.Handle(params, pos, vel, e, None, ids, capacity=self.capacity)
        self.params = params
        self.device = torch.device("cuda", params.device)
        self.halo_cap = int(halo_cap or max(4096, n // 8))
        self.buf = [torch.empty((self.halo_cap, MIGRANT_DOUBLES), dtype=torch.float64, device=self.device) for _ in range(2)]
        self.h_max = float(h_max_hint) if h_max_hint else 0.0
        self.x_lo, self.x_hi = topo.interval(rank)
        self.has_left = topo.left(rank) is not None
        self.has_right = topo.right(rank) is not None
        self.mode_next = 1  # step 0 starts with the initialisation evaluation (sph.go:89-103)

    # -- thin ctypes calls
    def _chk(self, rc):
        self.h._chk(rc)

    def set_widths(self, h_max):
        gw, iw = ghost_widths(h_max, self.safety)
        sl = self.L.Slab(self.x_lo if self.has_left else -1e300, self.x_hi if self.has_right else 1e300,
                         gw, iw, int(self.has_left), int(self.has_right))
        self._chk(self.L.lib().sphb_slab_set(self.h._h, C.byref(sl)))

    def begin(self, mode):
        self._chk(self.L.lib().sphb_slab_step_begin(self.h._h, mode))

    def _pack(self, fn, side, view):
        cnt = C.c_int64()
        self._chk(fn(self.h._h, side, C.c_void_p(view.data_ptr()), view.shape[0], C.byref(cnt)))
        return cnt.value

    def pack_halo(self):
        """-> (buf_left, n_left, buf_right, n_right); buffers are [cap, HALO_DOUBLES] views; one pass, one readback"""
        lib = self.L.lib()
        views = [self.buf[s].view(-1)[: self.halo_cap * HALO_DOUBLES].view(self.halo_cap, HALO_DOUBLES) for s in range(2)]
        cnt = (C.c_int64 * 2)()
        self._chk(lib.sphb_slab_pack_halo(self.h._h, C.c_void_p(views[0].data_ptr()), C.c_void_p(views[1].data_ptr()),
                                          self.halo_cap, cnt))
        return views[0], int(cnt[0]), views[1], int(cnt[1])

    def add_ghosts(self, buf, count):
        if count:
            buf = buf[:count].contiguous()
            self._chk(self.L.lib().sphb_slab_add_ghosts(self.h._h, C.c_void_p(buf.data_ptr()), count))
            self._keep = getattr(self, "_keep", []) + [buf]

    def end(self, integrate):
        self._chk(self.L.lib().sphb_slab_step_end(self.h._h, int(integrate)))
        self._keep = []

    def pack_migrants(self):
        lib = self.L.lib()
        views = [self.buf[s].view(-1)[: self.halo_cap * MIGRANT_DOUBLES].view(self.halo_cap, MIGRANT_DOUBLES) for s in range(2)]
        n0 = self._pack(lib.sphb_slab_pack_migrants, 0, views[0])
        n1 = self._pack(lib.sphb_slab_pack_migrants, 1, views[1])
        return views[0], n0, views[1], n1

    def add_migrants(self, buf, count):
        if count:
            buf = buf[:count].contiguous()
            self._chk(self.L.lib().sphb_slab_add_migrants(self.h._h, C.c_void_p(buf.data_ptr()), count))
            self.h.sync()

    def finish_migration(self):
        self._chk(self.L.lib().sphb_slab_finish_migration(self.h._h))

    def local_max_h(self):
        return self.h.max_h()

    def local_max_speed(self):
        return self.h.max_speed()


GHOST_SLACK = 0.5  # ghost_widths(): an owned particle may sit this many h_max outside its slab


class MigrationSchedule:
    """When to migrate.  `every` = k > 0: after every k-th step (1 = the plain protocol).  `every` = 0: only when
    it is needed - the ghost layer tolerates owned particles up to GHOST_SLACK * h_max outside their slab, and a
    particle moves at most (|v_before| + |v_after|) * dt_half per step, so the run keeps an upper bound of the
    excursion since the last migration (from the all-reduced max speed) and migrates before the bound, plus the
    next step's movement, could reach `fraction` of the slack."""

    def __init__(self, every: int, dt_half: float, fraction: float = 0.8):
        self.every, self.dt_half, self.fraction = int(every or 0), float(dt_half), fraction
        self.excursion = 0.0
        self.v_prev = 0.0
        self.steps = 0
        self.migrations = 0

    def after_step(self, v_max: float, h_max: float) -> bool:
        """called once per completed step with the global max speed / max h; True = migrate now"""
        self.steps += 1
        self.excursion += (self.v_prev + v_max) * self.dt_half
        self.v_prev = v_max
        if self.every > 0:
            due = self.steps % self.every == 0
        else:
            nxt = 2.0 * v_max * self.dt_half * 1.5  # the coming step (speeds may grow a little through the kick)
            due = self.excursion + nxt > self.fraction * GHOST_SLACK * h_max
        if due:
            self.excursion = 0.0
            self.migrations += 1
        return due


def default_h_hint(n_total: int, area: float) -> float:
    """first-evaluation guess of the largest smoothing length: 32 neighbours at the mean density, x1.6"""
    return 1.6 * math.sqrt(33.0 * area / (math.pi * max(n_total, 1)))


def _route_to_owner(topo: "Topology", rank: int, pos, vel, e, rho, ids):
    """the rows of a batch of spawned particles that belong to `rank`'s slab (x folded into the period on a ring)"""
    pos = np.ascontiguousarray(pos, dtype=np.float64).reshape(-1, 2)
    if ids is None:
        raise ValueError("slab append needs explicit global ids (every rank must number the new particles alike)")
    x = pos[:, 0]
    if topo.periodic:
        lo, hi = float(topo.bounds[0]), float(topo.bounds[-1])
        x = lo + np.mod(x - lo, hi - lo)
    m = topo.owner_of(x) == rank
    pick = lambda a, shape: None if a is None else np.ascontiguousarray(a, dtype=np.float64).reshape(shape)[m]  # noqa: E731
    n = len(pos)
    return pos[m], pick(vel, (n, 2)), pick(e, (n,)), pick(rho, (n,)), np.asarray(ids, dtype=np.int64).reshape(n)[m]


class DistSlabSim:
    """one rank of a torch.distributed slab run: same step / state surface as a single libsphb handle"""

    def __init__(self, params, topo: Topology, rank: int, pos, vel=None, e=None, ids=None, h_max_hint=None,
                 capacity=None, halo_cap=None, migrate_every=1, safety=1.15):
        self.slab = Slab(params, topo, rank, pos, vel, e, ids, capacity, halo_cap, h_max_hint, safety)
        self.ex = DistExchange(topo, rank, self.slab.device)
        self.h_max = float(h_max_hint or 0.0)
        self.v_max = 0.0
        self.schedule = MigrationSchedule(migrate_every, params.dt_half)
        self.steps_done = 0
        self.h_growth = 1.0
        import os
        self.profile = os.environ.get("SPHB_SLAB_PROFILE") == "1"
        self.prof = {}

    def _tick(self, name):
        """SPHB_SLAB_PROFILE=1: wall time per protocol phase with a device sync after each (diagnostic only)"""
        import time
        self.slab.h.sync()
        t = time.perf_counter()
        self.prof[name] = self.prof.get(name, 0.0) + (t - self._t0)
        self._t0 = t

    def _evaluate(self, mode, integrate):
        s = self.slab
        if self.profile:
            import time
            s.h.sync()
            self._t0 = time.perf_counter()
        s.set_widths(self.h_max * self.h_growth)
        s.begin(mode)
        if self.profile: self._tick("begin")
        bl, nl, br, nr = s.pack_halo()
        if self.profile: self._tick("pack_halo")
        rl, kl, rr, kr = self.ex.exchange(bl, nl, br, nr, HALO_DOUBLES)
        if self.profile: self._tick("exchange")
        s.add_ghosts(rl, kl)
        s.add_ghosts(rr, kr)
        if self.profile: self._tick("add_ghosts")
        s.end(integrate)
        if self.profile: self._tick("end")
        # one tiny all-reduce per evaluation (also surfaces GHOST_THIN / overflow errors)
        self.h_max, self.v_max = self.ex.allreduce_max(s.local_max_h(), s.local_max_speed())
        if self.profile: self._tick("allreduce")

    def _migrate(self):
        s = self.slab
        bl, nl, br, nr = s.pack_migrants()
        rl, kl, rr, kr = self.ex.exchange(bl, nl, br, nr, MIGRANT_DOUBLES)
        s.add_migrants(rl, kl)
        s.add_migrants(rr, kr)
        s.finish_migration()

    def step(self, nsteps=1):
        for _ in range(nsteps):
            if self.steps_done == 0:
                self._evaluate(1, False)
            self._evaluate(2, True)
            self.steps_done += 1
            if self.schedule.after_step(self.v_max, self.h_max):
                if self.profile:
                    import time
                    self.slab.h.sync()
                    self._t0 = time.perf_counter()
                self._migrate()
                if self.profile: self._tick("migrate")

    def append(self, pos, vel=None, e=None, rho=None, ids=None):
        """sources in a slab run (sph.go:72-86): every rank is handed the same spawned particles between
        two steps and keeps those whose x lies in its slab (sphb_append; no ghosts are attached between steps)"""
        part = _route_to_owner(self.slab.topo, self.slab.rank, pos, vel, e, rho, ids)
        if len(part[0]):
            self.slab.h.append(*part)

    @property
    def handle(self):
        return self.slab.h


class LocalSlabSim:
    """`world` slabs inside one process on one GPU, stepped in lock step: exercises the complete slab path
    (ghost packing, inner/outer ghosts, compaction, migration) without a second GPU"""

    def __init__(self, params, topo: Topology, pos, vel=None, e=None, ids=None, h_max_hint=None, migrate_every=1,
                 safety=1.15, halo_cap=None):
        pos = np.ascontiguousarray(pos, dtype=np.float64).reshape(-1, 2)
        n = len(pos)
        ids = np.arange(n, dtype=np.int64) if ids is None else np.asarray(ids, dtype=np.int64)
        vel = np.zeros((n, 2)) if vel is None else np.asarray(vel, dtype=np.float64).reshape(n, 2)
        e = np.zeros(n) if e is None else np.asarray(e, dtype=np.float64)
        owner = topo.owner_of(pos[:, 0])
        self.topo, self.ex = topo, LocalExchange(topo)
        self.slabs = []
        for r in range(topo.world):
            m = owner == r
            self.slabs.append(Slab(params, topo, r, pos[m], vel[m], e[m], ids[m], capacity=n, halo_cap=halo_cap or n,
                                   h_max_hint=h_max_hint, safety=safety))
        self.h_max = float(h_max_hint or 0.0)
        self.v_max = 0.0
        self.steps_done = 0
        self.schedule = MigrationSchedule(migrate_every, params.dt_half)

    def _evaluate(self, mode, integrate):
        for s in self.slabs:
            s.set_widths(self.h_max)
            s.begin(mode)
            self.ex.post(s.rank, *s.pack_halo())
        for s in self.slabs:
            rl, kl, rr, kr = self.ex.collect(s.rank)
            s.add_ghosts(rl, kl)
            s.add_ghosts(rr, kr)
        for s in self.slabs:
            s.end(integrate)
        self.h_max = max(s.local_max_h() for s in self.slabs)
        self.v_max = max(s.local_max_speed() for s in self.slabs)

    def _migrate(self):
        for s in self.slabs:
            self.ex.post(s.rank, *s.pack_migrants())
        for s in self.slabs:
            rl, kl, rr, kr = self.ex.collect(s.rank)
            s.add_migrants(rl, kl)
            s.add_migrants(rr, kr)
        for s in self.slabs:
            s.finish_migration()

    def step(self, nsteps=1):
        for _ in range(nsteps):
            if self.steps_done == 0:
                self._evaluate(1, False)
            self._evaluate(2, True)
            self.steps_done += 1
            if self.schedule.after_step(self.v_max, self.h_max):
                self._migrate()

    def append(self, pos, vel=None, e=None, rho=None, ids=None):
        """sources: the new particles go to the slab that owns their x (see DistSlabSim.append)"""
        for sl in self.slabs:
            part = _route_to_owner(self.topo, sl.rank, pos, vel, e, rho, ids)
            if len(part[0]):
                sl.h.append(*part)

    def state(self, fields):
        """concatenated over the slabs, sorted by id"""
        parts = [s.h.download(list(dict.fromkeys(list(fields) + ["id"]))) for s in self.slabs]
        d = {k: np.concatenate([p[k] for p in parts], axis=0) for k in parts[0]}
        o = np.argsort(d["id"], kind="stable")
        return {k: v[o] for k, v in d.items()}

    def counts(self):
        return [s.h.n for s in self.slabs]

    def close(self):
        for s in self.slabs:
            s.h.close()


# ------------------------------------------------------------------------------------------------
# the ring inside the library (include/sphb.h "slab ring"): the protocol above, in C++, behind sphb_ring_step
# ------------------------------------------------------------------------------------------------
class LocalRingSim:
    """all slabs of a ring in this process, stepped by sphb_ring_step_local (the exchange is a device copy): the whole
    in-library protocol - halo, ghosts kept across reuse evaluations, migration, schedule - on one GPU"""

    def __init__(self, params, topo: Topology, pos, vel=None, e=None, ids=None, h_max_hint=None, migrate_every=0,
                 safety=0.0, halo_cap=None):
        from . import _lib as L
        self.L = L
        pos = np.ascontiguousarray(pos, dtype=np.float64).reshape(-1, 2)
        n = len(pos)
        ids = np.arange(n, dtype=np.int64) if ids is None else np.asarray(ids, dtype=np.int64)
        vel = np.zeros((n, 2)) if vel is None else np.asarray(vel, dtype=np.float64).reshape(n, 2)
        e = np.zeros(n) if e is None else np.asarray(e, dtype=np.float64)
        owner = topo.owner_of(pos[:, 0])
        self.topo, self.handles = topo, []
        for r in range(topo.world):
            m = owner == r
            h = L.Handle(params, pos[m], vel[m], e[m], None, ids[m], capacity=n)
            lo, hi = topo.interval(r)
            h.ring_set(r, topo.world, topo.periodic, lo, hi, float(h_max_hint or default_h_hint(n, 1.0)), halo_cap or n,
                       safety, migrate_every)
            self.handles.append(h)

    def step(self, nsteps=1):
        self.L.ring_step_local(self.handles, nsteps)

    def append(self, pos, vel=None, e=None, rho=None, ids=None):
        for r, h in enumerate(self.handles):
            part = _route_to_owner(self.topo, r, pos, vel, e, rho, ids)
            if len(part[0]):
                h.append(*part)

    def state(self, fields):
        parts = [h.download(list(dict.fromkeys(list(fields) + ["id"]))) for h in self.handles]
        d = {k: np.concatenate([p[k] for p in parts], axis=0) for k in parts[0]}
        o = np.argsort(d["id"], kind="stable")
        return {k: v[o] for k, v in d.items()}

    def counts(self):
        return [h.n for h in self.handles]

    def info(self):
        return [h.ring_info() for h in self.handles]

    def close(self):
        for h in self.handles:
            h.close()


class RingSim:
    """one rank of a multi-GPU ring: everything happens inside libsphb (NCCL); torch.distributed is used once, to hand
    the NCCL unique id from rank 0 to the others"""

    def __init__(self, params, topo: Topology, rank: int, pos, vel=None, e=None, ids=None, h_max_hint=None,
                 capacity=None, halo_cap=None, migrate_every=0, safety=0.0):
        import torch
        import torch.distributed as dist
        from . import _lib as L
        pos = np.ascontiguousarray(pos, dtype=np.float64).reshape(-1, 2)
        n = len(pos)
        self.h = L.Handle(params, pos, vel, e, None, ids, capacity=int(capacity or (n + max(4096, n // 4))))
        lo, hi = topo.interval(rank)
        self.h.ring_set(rank, topo.world, topo.periodic, lo, hi, float(h_max_hint or 0.0), int(halo_cap or max(4096, n // 8)),
                        safety, migrate_every)
        dev = torch.device("cuda", params.device) if dist.get_backend() == "nccl" else torch.device("cpu")
        idt = torch.zeros(L.NCCL_ID_BYTES, dtype=torch.uint8, device=dev)
        if rank == 0:
            idt.copy_(torch.frombuffer(bytearray(L.comm_unique_id()), dtype=torch.uint8))
        dist.broadcast(idt, 0)
        self.h.comm_init(bytes(idt.cpu().numpy().tobytes()), rank, topo.world)
        self.topo, self.rank = topo, rank

    def step(self, nsteps=1):
        self.h.ring_step(nsteps)

    def append(self, pos, vel=None, e=None, rho=None, ids=None):
        part = _route_to_owner(self.topo, self.rank, pos, vel, e, rho, ids)
        if len(part[0]):
            self.h.append(*part)

    @property
    def handle(self):
        return self.h


# ------------------------------------------------------------------------------------------------
# bench.py --gpus N (N > 1): weak scaling, one rank per GPU
# ------------------------------------------------------------------------------------------------
def bench(args, nx, ny, box, phys, desc, rank, world, local):
    """one rank of `bench.py --gpus N`: the ring inside the library (sphb_ring_step over NCCL)"""
    import time
    import torch
    import torch.distributed as dist
    from . import _lib as L
    import bench as B  # repo-root bench.py (helpers: make_ic, ClockSampler, measured_peak)

    if args.workload in ("c4", "c4dam"):  # strong scaling, non-uniform density: equal-count slabs
        pos_all, kw, desc4 = B.leg_ic(args.workload)
        periodic = args.workload == "c4"
        lo, hi = (0.0, 1.0) if periodic else (float(pos_all[:, 0].min()), float(pos_all[:, 0].max()) + 1e-9)
        bounds = equal_count_bounds(pos_all[:, 0], world, lo, hi)
        m = (pos_all[:, 0] >= bounds[rank]) & (pos_all[:, 0] < bounds[rank + 1])
        pos, ids = np.ascontiguousarray(pos_all[m]), np.nonzero(m)[0].astype(np.int64)
        n_local, n_total = len(pos), len(pos_all)
        del pos_all
        h_hint = 2.0 * default_h_hint(n_total, 1.0 if periodic else 0.125)  # (the dilute half has twice the mean spacing)
        prm = L.make_params(precision=args.precision, device=local, **kw)
        desc = desc4 + f"; {world} GPU(s), equal-count x-slabs, strong scaling"
    else:
        pos, (x0, x1) = B.make_ic(nx, ny, box, rank, world)
        n_local = len(pos)
        n_total = nx * ny
        bounds = [box[0] * k / world for k in range(world + 1)]
        ids = np.arange(n_local, dtype=np.int64) + rank * n_local
        h_hint = default_h_hint(n_total, box[0] * box[1])
        periodic = True
        prm = L.make_params(hor=(0.0, box[0]), ver=(0.0, box[1]), device=local, **phys)
    topo = Topology(world, bounds, periodic=periodic)
    sim = RingSim(prm, topo, rank, pos, None, np.full(n_local, 0.01), ids, h_max_hint=h_hint,
                  capacity=n_local + max(1 << 20, n_local // 8), halo_cap=max(1 << 18, n_local // 16))
    del pos
    K, W = args.steps, max(args.warmup, 3)
    sim.step(1 + W)
    sim.handle.sync()
    c0 = sim.handle.counters()
    sampler = B.ClockSampler(local)
    sampler.start()
    ext = torch.cuda.ExternalStream(sim.handle.stream, device=torch.device("cuda", local))
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    ev0.record(ext)
    sim.step(K)
    ev1.record(ext)
    sim.handle.sync()
    torch.cuda.synchronize()
    dist.barrier()
    wall = time.perf_counter() - t0
    dev_ms = ev0.elapsed_time(ev1)
    clocks = sampler.stop()
    c1 = sim.handle.counters()
    info = sim.handle.ring_info()
    t = torch.tensor([dev_ms, wall * 1e3], dtype=torch.float64, device=torch.device("cuda", local))
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    cnt = torch.tensor([sim.handle.n, c1["knn_fallback"] - c0["knn_fallback"]], dtype=torch.int64, device=torch.device("cuda", local))
    dist.all_reduce(cnt, op=dist.ReduceOp.SUM)
    dev_ms, wall_ms = float(t[0]), float(t[1])
    # device phases of separate untimed steps on rank 0 (library CUDA-event timers), by kind of evaluation
    KP = max(2, min(K, 8))
    ph = {"rebuild": [], "reuse": []}
    for _ in range(KP):
        r0 = sim.handle.counters()["reuse_steps"]
        sim.step(1)
        kind = "reuse" if sim.handle.counters()["reuse_steps"] > r0 else "rebuild"
        ph[kind].append(sim.handle.phase_times())
    device_phases = {k: {f: float(np.mean([p[f] for p in v])) for f in v[0]} for k, v in ph.items() if v}
    if rank == 0:
        peak, peak_src = B.measured_peak()
        ms_per_step = dev_ms / K
        value = n_total * K / (dev_ms * 1e-3)
        achieved = n_total * B.B_ALG_TOTAL / (ms_per_step * 1e-3) / 1e9 / world  # per GPU
        reuse_steps = c1["reuse_steps"] - c0["reuse_steps"]
        line = {
            "metric": B.METRIC, "value": value, "unit": B.UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "strong" if args.workload.startswith("c4") else "weak", "vs_baseline": None,
            "dtype": "f64" if args.precision == 64 else "f32", "data": "synthetic",
            "config": {"workload": desc, "particles": n_total, "particles_after": int(cnt[0]),
                       "decomposition": f"{world} x-slabs, {'periodic ring' if periodic else 'open chain'}, inside libsphb (sphb_ring_step): NCCL send/recv of the halo "
                                        f"on the library stream, ghosts kept across reuse evaluations, migration at rebuilds when the excursion bound nears "
                                        f"the ghost slack ({int(info['migrations'])} so far)",
                       "timing": "CUDA events on each rank's library stream around K steps, max over ranks",
                       "wall_ms_per_step": wall_ms / K},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": None, "peak_source": peak_src, "kernel": "whole step, per GPU",
                         "alg_bytes_per_particle": B.B_ALG_TOTAL},
            "cpu_baseline": None,
            "e2e": {"value": n_total * K / (wall_ms * 1e-3), "unit": B.UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0,
                    "what": "wall clock of the same K steps including host orchestration and NCCL exchange; state stays device resident"},
            "gpu_launches": c1["kernel_launches"] - c0["kernel_launches"],
            "knn_fallback_particles": int(cnt[1]),
            "reuse": {"reuse_steps": reuse_steps, "rebuild_steps": K - reuse_steps, "fixed_period": int(info["period"]),
                      "refused_fraction": int(cnt[1]) / (n_total * K), "ghosts_rank0": int(info["ghosts"])},
            "clocks": clocks,
            "phases": {"device_ms": device_phases,
                       "what": f"rank 0, {KP} separate untimed steps: library CUDA-event timers per kind of evaluation; for reuse evaluations "
                               "'keys' = predict + halo gather / exchange / scatter, for rebuilds the exchange precedes the timers"},
        }
        B.emit(line)
    dist.barrier()
    sim.handle.close()
    dist.destroy_process_group()
