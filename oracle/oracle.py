"""ctypes binding of oracle/liborc.so — TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module (see the header of oracle/sph_oracle.c).  The product package
sphugo_b200 never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

NN = 32
OPEN = (-1.7976931348623157e308, 1.7976931348623157e308)


class OrcParams(C.Structure):
    """Mirror of sphb_params (include/sphb.h) == numeric part of SphConfig (config-parser.go:111-128)."""

    _fields_ = [
        ("dt_half", C.c_double), ("gamma", C.c_double), ("particle_mass", C.c_double),
        ("accel", C.c_double * 2), ("hor", C.c_double * 2), ("ver", C.c_double * 2),
        ("refl_L", C.c_double), ("refl_R", C.c_double), ("refl_U", C.c_double), ("refl_D", C.c_double),
        ("kernel", C.c_int32), ("precision", C.c_int32), ("device", C.c_int32), ("flags", C.c_int32),
    ]


def build(force: bool = False) -> str:
    so = os.path.join(_HERE, "liborc.so")
    src = os.path.join(_HERE, "sph_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s", "liborc.so"])
    return so


def lib():
    global _LIB
    if _LIB is None:
        L = C.CDLL(build())
        dp, ip = C.POINTER(C.c_double), C.POINTER(C.c_int64)
        L.orc_create.restype = C.c_void_p
        L.orc_create.argtypes = [C.POINTER(OrcParams), C.c_int64, C.c_int64, dp, dp, dp, dp, ip]
        L.orc_destroy.argtypes = [C.c_void_p]
        L.orc_set_params.argtypes = [C.c_void_p, C.POINTER(OrcParams)]
        for f in ("orc_count", "orc_current_step", "orc_underfull"):
            getattr(L, f).restype = C.c_int64
            getattr(L, f).argtypes = [C.c_void_p]
        for f in ("orc_status", "orc_stale_root_child", "orc_calc_forces", "orc_step"):
            getattr(L, f).restype = C.c_int
            getattr(L, f).argtypes = [C.c_void_p]
        L.orc_sizeof_particle.restype = C.c_int64
        L.orc_append.restype = C.c_int
        L.orc_append.argtypes = [C.c_void_p, C.c_int64, dp, dp, dp, dp, ip]
        L.orc_knn.restype = C.c_int
        L.orc_knn.argtypes = [C.c_void_p, dp, dp, C.c_int, C.c_int]
        L.orc_density.restype = C.c_int
        L.orc_density.argtypes = [C.c_void_p, C.c_int]
        L.orc_calc_forces_mode.restype = C.c_int
        L.orc_calc_forces_mode.argtypes = [C.c_void_p, C.c_int]
        L.orc_step_mode.restype = C.c_int
        L.orc_step_mode.argtypes = [C.c_void_p, C.c_int]
        L.orc_run.restype = C.c_int
        L.orc_run.argtypes = [C.c_void_p, C.c_int, C.c_int]
        for f in ("orc_total_energy", "orc_total_density", "orc_total_momentum"):
            getattr(L, f).restype = C.c_double
            getattr(L, f).argtypes = [C.c_void_p]
        L.orc_get.argtypes = [C.c_void_p] + [dp] * 10 + [ip, ip, dp, dp]
        L.orc_set.argtypes = [C.c_void_p, dp, dp, dp, dp]
        L.orc_partition.restype = C.c_int64
        L.orc_partition.argtypes = [dp, C.c_int64, C.c_int, C.c_double]
        L.orc_tree_count_outside_all.restype = C.c_int64
        L.orc_tree_count_outside_all.argtypes = [C.c_void_p]
        L.orc_tree_stats.argtypes = [C.c_void_p, ip]
        L.orc_heap_build.argtypes = [ip, C.c_int64]
        L.orc_heap_insert.restype = C.c_int64
        L.orc_heap_insert.argtypes = [ip, C.c_int64, C.c_int64]
        L.orc_heap_extract_min.restype = C.c_int64
        L.orc_heap_extract_min.argtypes = [ip, C.c_int64, ip]
        L.orc_heap_replace.restype = C.c_int64
        L.orc_heap_replace.argtypes = [ip, C.c_int64, C.c_int64, ip]
        _LIB = L
    return _LIB


def _dp(a):
    return None if a is None else a.ctypes.data_as(C.POINTER(C.c_double))


def _ip(a):
    return None if a is None else a.ctypes.data_as(C.POINTER(C.c_int64))


def _f64(a, shape=None):
    if a is None:
        return None
    a = np.ascontiguousarray(a, dtype=np.float64)
    if shape is not None:
        a = a.reshape(shape)
    return a


def make_params(dt_half=0.001, gamma=1.66666, particle_mass=1.0, accel=(0.0, 0.0), hor=OPEN, ver=OPEN,
                refl=(OPEN[0], OPEN[1], OPEN[0], OPEN[1]), kernel=1, precision=64, device=0, flags=0):
    """Defaults == sim.MakeConfig() (config-parser.go:131-149)."""
    p = OrcParams()
    p.dt_half, p.gamma, p.particle_mass = dt_half, gamma, particle_mass
    p.accel[0], p.accel[1] = accel
    p.hor[0], p.hor[1] = hor
    p.ver[0], p.ver[1] = ver
    p.refl_L, p.refl_R, p.refl_U, p.refl_D = refl
    p.kernel, p.precision, p.device, p.flags = kernel, precision, device, flags
    return p


class Oracle:
    """CPU reference simulation. knn_mode 0 = faithful tree walk, 1 = exact brute force."""

    def __init__(self, params, pos, vel=None, e=None, rho=None, ids=None, capacity=None):
        L = lib()
        pos = _f64(pos, (-1, 2))
        n = pos.shape[0]
        vel, e, rho = _f64(vel, (-1, 2)), _f64(e), _f64(rho)
        ids = None if ids is None else np.ascontiguousarray(ids, dtype=np.int64)
        self._p = OrcParams()
        C.memmove(C.byref(self._p), C.byref(params), C.sizeof(OrcParams))
        self._h = L.orc_create(C.byref(self._p), n, capacity or n, _dp(pos), _dp(vel), _dp(e), _dp(rho), _ip(ids))
        if not self._h:
            raise MemoryError("orc_create failed")

    def close(self):
        if getattr(self, "_h", None):
            lib().orc_destroy(self._h)
            self._h = None

    __del__ = close

    @property
    def n(self):
        return lib().orc_count(self._h)

    @property
    def current_step(self):
        return lib().orc_current_step(self._h)

    @property
    def underfull(self):
        return lib().orc_underfull(self._h)

    @property
    def stale_root_child(self):
        return bool(lib().orc_stale_root_child(self._h))

    def set_params(self, params):
        C.memmove(C.byref(self._p), C.byref(params), C.sizeof(OrcParams))
        lib().orc_set_params(self._h, C.byref(self._p))

    def _chk(self, rc, what):
        if rc:
            raise RuntimeError(f"oracle {what}: reference would panic / fail (code {rc})")

    def knn(self, hor=None, ver=None, mode=0, rebuild=True):
        hor = np.asarray(hor if hor is not None else list(self._p.hor), dtype=np.float64)
        ver = np.asarray(ver if ver is not None else list(self._p.ver), dtype=np.float64)
        self._chk(lib().orc_knn(self._h, _dp(hor), _dp(ver), mode, int(rebuild)), "knn")

    def density(self, kernel=None):
        self._chk(lib().orc_density(self._h, self._p.kernel if kernel is None else kernel), "density")

    def calc_forces(self, knn_mode=0):
        self._chk(lib().orc_calc_forces_mode(self._h, knn_mode), "calc_forces")

    def step(self, nsteps=1, knn_mode=0):
        self._chk(lib().orc_run(self._h, nsteps, knn_mode), "step")

    def append(self, pos, vel=None, e=None, rho=None, ids=None):
        pos = _f64(pos, (-1, 2))
        ids = None if ids is None else np.ascontiguousarray(ids, dtype=np.int64)
        self._chk(lib().orc_append(self._h, pos.shape[0], _dp(pos), _dp(_f64(vel)), _dp(_f64(e)), _dp(_f64(rho)), _ip(ids)), "append")

    def total_energy(self):
        return lib().orc_total_energy(self._h)

    def total_density(self):
        return lib().orc_total_density(self._h)

    def total_momentum(self):
        return lib().orc_total_momentum(self._h)

    def outside_all_circles(self):
        return lib().orc_tree_count_outside_all(self._h)

    def tree_stats(self):
        out = np.zeros(4, dtype=np.int64)
        lib().orc_tree_stats(self._h, _ip(out))
        return dict(nodes=int(out[0]), leaves=int(out[1]), max_leaf=int(out[2]), depth=int(out[3]))

    def tree(self):
        """the tree as arrays in pre-order: geo [nodes, 7] = LowerLeft, UpperRight, BCenter, BRadius; link [nodes, 4] =
        index of Lower, index of Upper (-1 = nil), first particle (current order), particle count"""
        L = lib()
        L.orc_tree_dump.restype = C.c_int64
        L.orc_tree_dump.argtypes = [C.c_void_p, C.c_int64, C.POINTER(C.c_double), C.POINTER(C.c_int64)]
        n = L.orc_tree_dump(self._h, 0, None, None)
        geo, link = np.zeros((n, 7)), np.zeros((n, 4), dtype=np.int64)
        L.orc_tree_dump(self._h, n, _dp(geo), _ip(link))
        return geo, link

    def state(self, neighbours=False, sort_by_id=True):
        """Dict of numpy arrays. With sort_by_id rows are ordered by particle id (ids must be unique)."""
        n = self.n
        d = dict(pos=np.empty((n, 2)), vel=np.empty((n, 2)), rho=np.empty(n), c=np.empty(n), e=np.empty(n),
                 edot=np.empty(n), vdot=np.empty((n, 2)), epred=np.empty(n), vpred=np.empty((n, 2)), h=np.empty(n),
                 id=np.empty(n, dtype=np.int64))
        nn_id = nn_dist = nn_pos = None
        if neighbours:
            nn_id = np.empty((n, NN), dtype=np.int64)
            nn_dist = np.empty((n, NN))
            nn_pos = np.empty((n, NN, 2))
        lib().orc_get(self._h, _dp(d["pos"]), _dp(d["vel"]), _dp(d["rho"]), _dp(d["c"]), _dp(d["e"]), _dp(d["edot"]),
                      _dp(d["vdot"]), _dp(d["epred"]), _dp(d["vpred"]), _dp(d["h"]), _ip(d["id"]), _ip(nn_id),
                      _dp(nn_dist), _dp(nn_pos))
        if neighbours:
            d.update(nn_id=nn_id, nn_dist=nn_dist, nn_pos=nn_pos)
        if sort_by_id:
            o = np.argsort(d["id"], kind="stable")
            d = {k: v[o] for k, v in d.items()}
        return d


# ---- KAT helpers -------------------------------------------------------------------------------
def partition(points, orientation, middle):
    """Partition (core.go:126-164) on a list of (x, y); returns (len(a), len(b), permuted points)."""
    a = _f64(np.array(points, dtype=np.float64).reshape(-1, 2)).copy()
    n = a.shape[0]
    la = lib().orc_partition(_dp(a), n, orientation, float(middle))
    return int(la), int(n - la), a


def heap_build(arr):
    a = np.array(arr, dtype=np.int64)
    lib().orc_heap_build(_ip(a), len(a))
    return a.tolist()


def heap_insert(arr, x):
    a = np.array(list(arr) + [0], dtype=np.int64)
    n = lib().orc_heap_insert(_ip(a), len(arr), x)
    return a[:n].tolist()


def heap_extract_min(arr):
    a = np.array(arr, dtype=np.int64)
    m = C.c_int64()
    n = lib().orc_heap_extract_min(_ip(a), len(a), C.byref(m))
    return a[:n].tolist(), m.value


def heap_replace(arr, x):
    a = np.array(arr, dtype=np.int64)
    m = C.c_int64()
    n = lib().orc_heap_replace(_ip(a), len(a), x, C.byref(m))
    return a[:n].tolist(), m.value
