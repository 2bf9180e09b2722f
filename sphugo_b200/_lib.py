"""ctypes binding of libsphb.so — exactly the C ABI of include/sphb.h (what the Go cgo shim binds).

No numeric work happens in Python; without the built CUDA library every call fails loudly
(there is no CPU fallback anywhere in this package).
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libsphb.so")

NN = 32
OPEN_LO, OPEN_HI = -1.7976931348623157e308, 1.7976931348623157e308
OPEN = (OPEN_LO, OPEN_HI)

OK, E_INVALID, E_CUDA, E_NOMEM, E_KNN_UNDERFULL, E_KERNEL, E_STATE, E_GHOST_THIN = 0, -1, -2, -3, -4, -5, -6, -7
KERNEL_TOPHAT, KERNEL_MONAGHAN, KERNEL_WENDLAND = 0, 1, 2

FIELDS = ["pos", "vel", "rho", "c", "e", "edot", "vdot", "epred", "vpred", "h", "id", "nn_idx", "nn_dist", "nn_pos"]
FIELD_BIT = {name: i for i, name in enumerate(FIELDS)}
FIELD_SHAPE = {  # trailing shape, dtype
    "pos": ((2,), np.float64), "vel": ((2,), np.float64), "rho": ((), np.float64), "c": ((), np.float64),
    "e": ((), np.float64), "edot": ((), np.float64), "vdot": ((2,), np.float64), "epred": ((), np.float64),
    "vpred": ((2,), np.float64), "h": ((), np.float64), "id": ((), np.int64), "nn_idx": ((NN,), np.int32),
    "nn_dist": ((NN,), np.float64), "nn_pos": ((NN, 2), np.float64),
}
SUM_E, SUM_RHO, LAST_VEL_NORM = 0, 1, 2
PHASES = ["keys", "sort", "reorder", "knn", "force", "total"]
COUNTERS = ["steps", "kernel_launches", "knn_fallback", "regrids", "reuse_steps"]
HALO_RECORD_DOUBLES, MIGRANT_RECORD_DOUBLES = 10, 12

EXPORTS = [
    "sphb_create", "sphb_destroy", "sphb_last_error", "sphb_set_params", "sphb_get_params", "sphb_count",
    "sphb_current_step", "sphb_set_current_step", "sphb_append", "sphb_step", "sphb_calc_forces", "sphb_knn", "sphb_density", "sphb_stream", "sphb_sync",
    "sphb_download", "sphb_upload", "sphb_upload_by_id", "sphb_reduce", "sphb_frame", "sphb_phase_times", "sphb_counters", "sphb_create_device",
    "sphb_slab_set", "sphb_max_h", "sphb_max_speed", "sphb_slab_step_begin", "sphb_slab_pack_halo", "sphb_slab_add_ghosts",
    "sphb_slab_step_end", "sphb_slab_pack_migrants", "sphb_slab_add_migrants", "sphb_slab_finish_migration",
    "sphb_upload_by_id_begin", "sphb_upload_by_id_end",
    "sphb_comm_unique_id", "sphb_comm_init", "sphb_ring_set", "sphb_ring_step", "sphb_ring_step_local", "sphb_ring_info",
]
NCCL_ID_BYTES = 128
RING_INFO = ["h_max", "v_max", "migrations", "ghosts", "period", "excursion"]


class Params(C.Structure):
    """sphb_params == numeric part of sim.SphConfig (config-parser.go:111-128)."""

    _fields_ = [
        ("dt_half", C.c_double), ("gamma", C.c_double), ("particle_mass", C.c_double),
        ("accel", C.c_double * 2), ("hor", C.c_double * 2), ("ver", C.c_double * 2),
        ("refl_L", C.c_double), ("refl_R", C.c_double), ("refl_U", C.c_double), ("refl_D", C.c_double),
        ("kernel", C.c_int32), ("precision", C.c_int32), ("device", C.c_int32), ("flags", C.c_int32),
    ]


class Slab(C.Structure):
    """sphb_slab"""

    _fields_ = [("x_lo", C.c_double), ("x_hi", C.c_double), ("ghost_w", C.c_double), ("inner_w", C.c_double),
                ("has_left", C.c_int32), ("has_right", C.c_int32)]


class Ring(C.Structure):
    """sphb_ring"""

    _fields_ = [("rank", C.c_int32), ("nranks", C.c_int32), ("periodic", C.c_int32), ("migrate_every", C.c_int32),
                ("x_lo", C.c_double), ("x_hi", C.c_double), ("h_hint", C.c_double), ("safety", C.c_double),
                ("halo_cap", C.c_int64)]


class SphbError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"libsphb error {code}: {msg}")
        self.code = code


_LIB = None


def lib():
    """Load libsphb.so; raises if it has not been built (python -m sphugo_b200.build)."""
    global _LIB
    if _LIB is not None:
        return _LIB
    if not os.path.exists(LIB_PATH):
        raise ImportError(f"{LIB_PATH} is missing: build it with `python -m sphugo_b200.build` "
                          "(the CUDA library is the only implementation; there is no CPU fallback)")
    L = C.CDLL(LIB_PATH)
    vp, dp, ip = C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_int64)
    pp = C.POINTER(Params)
    L.sphb_create.restype = C.c_int
    L.sphb_create.argtypes = [pp, C.c_int64, C.c_int64, vp, vp, vp, vp, vp, C.POINTER(vp)]
    L.sphb_create_device.restype = C.c_int
    L.sphb_create_device.argtypes = [pp, C.c_int64, C.c_int64, vp, vp, vp, vp, C.POINTER(vp)]
    L.sphb_destroy.restype = None
    L.sphb_destroy.argtypes = [vp]
    L.sphb_last_error.restype = C.c_char_p
    L.sphb_last_error.argtypes = [vp]
    L.sphb_set_params.restype = C.c_int
    L.sphb_set_params.argtypes = [vp, pp]
    L.sphb_get_params.restype = C.c_int
    L.sphb_get_params.argtypes = [vp, pp]
    L.sphb_count.restype = C.c_int64
    L.sphb_count.argtypes = [vp]
    L.sphb_current_step.restype = C.c_int64
    L.sphb_current_step.argtypes = [vp]
    L.sphb_set_current_step.restype = C.c_int
    L.sphb_set_current_step.argtypes = [vp, C.c_int64]
    L.sphb_append.restype = C.c_int
    L.sphb_append.argtypes = [vp, C.c_int64, vp, vp, vp, vp, vp]
    L.sphb_step.restype = C.c_int
    L.sphb_step.argtypes = [vp, C.c_int32]
    L.sphb_calc_forces.restype = C.c_int
    L.sphb_calc_forces.argtypes = [vp]
    L.sphb_knn.restype = C.c_int
    L.sphb_knn.argtypes = [vp, dp, dp]
    L.sphb_density.restype = C.c_int
    L.sphb_density.argtypes = [vp, C.c_int32]
    L.sphb_stream.restype = C.c_void_p
    L.sphb_stream.argtypes = [vp]
    L.sphb_sync.restype = C.c_int
    L.sphb_sync.argtypes = [vp]
    L.sphb_download.restype = C.c_int
    L.sphb_download.argtypes = [vp, C.c_uint32, C.POINTER(vp), C.c_int64, ip]
    L.sphb_upload.restype = C.c_int
    L.sphb_upload.argtypes = [vp, C.c_uint32, C.POINTER(vp), C.c_int64]
    L.sphb_reduce.restype = C.c_int
    L.sphb_reduce.argtypes = [vp, C.c_int32, dp]
    L.sphb_upload_by_id.restype = C.c_int
    L.sphb_upload_by_id.argtypes = L.sphb_upload.argtypes
    L.sphb_upload_by_id_begin.restype = C.c_int
    L.sphb_upload_by_id_begin.argtypes = L.sphb_upload.argtypes
    L.sphb_upload_by_id_end.restype = C.c_int
    L.sphb_upload_by_id_end.argtypes = [vp]
    L.sphb_frame.restype = C.c_int
    L.sphb_frame.argtypes = [vp, C.c_int32, C.c_int32, vp, vp, vp, C.c_int64, ip]
    L.sphb_phase_times.restype = C.c_int
    L.sphb_phase_times.argtypes = [vp, dp, C.c_int32]
    L.sphb_counters.restype = C.c_int
    L.sphb_counters.argtypes = [vp, ip, C.c_int32]
    L.sphb_slab_set.restype = C.c_int
    L.sphb_slab_set.argtypes = [vp, C.POINTER(Slab)]
    L.sphb_max_h.restype = C.c_int
    L.sphb_max_h.argtypes = [vp, dp]
    L.sphb_max_speed.restype = C.c_int
    L.sphb_max_speed.argtypes = [vp, dp]
    L.sphb_slab_step_begin.restype = C.c_int
    L.sphb_slab_step_begin.argtypes = [vp, C.c_int32]
    L.sphb_slab_pack_halo.restype = C.c_int
    L.sphb_slab_pack_halo.argtypes = [vp, vp, vp, C.c_int64, ip]
    L.sphb_slab_add_ghosts.restype = C.c_int
    L.sphb_slab_add_ghosts.argtypes = [vp, vp, C.c_int64]
    L.sphb_slab_step_end.restype = C.c_int
    L.sphb_slab_step_end.argtypes = [vp, C.c_int32]
    L.sphb_slab_pack_migrants.restype = C.c_int
    L.sphb_slab_pack_migrants.argtypes = [vp, C.c_int32, vp, C.c_int64, ip]
    L.sphb_slab_add_migrants.restype = C.c_int
    L.sphb_slab_add_migrants.argtypes = [vp, vp, C.c_int64]
    L.sphb_slab_finish_migration.restype = C.c_int
    L.sphb_slab_finish_migration.argtypes = [vp]
    L.sphb_comm_unique_id.restype = C.c_int
    L.sphb_comm_unique_id.argtypes = [vp]
    L.sphb_comm_init.restype = C.c_int
    L.sphb_comm_init.argtypes = [vp, vp, C.c_int32, C.c_int32]
    L.sphb_ring_set.restype = C.c_int
    L.sphb_ring_set.argtypes = [vp, C.POINTER(Ring)]
    L.sphb_ring_step.restype = C.c_int
    L.sphb_ring_step.argtypes = [vp, C.c_int32]
    L.sphb_ring_step_local.restype = C.c_int
    L.sphb_ring_step_local.argtypes = [C.POINTER(vp), C.c_int32, C.c_int32]
    L.sphb_ring_info.restype = C.c_int
    L.sphb_ring_info.argtypes = [vp, dp, C.c_int32]
    _LIB = L
    return L


def make_params(dt_half=0.001, gamma=1.66666, particle_mass=1.0, accel=(0.0, 0.0), hor=OPEN, ver=OPEN,
                refl=(OPEN_LO, OPEN_HI, OPEN_LO, OPEN_HI), kernel=KERNEL_MONAGHAN, precision=64, device=0, flags=0):
    """Defaults == sim.MakeConfig() (config-parser.go:131-149)."""
    p = Params()
    p.dt_half, p.gamma, p.particle_mass = dt_half, gamma, particle_mass
    p.accel[0], p.accel[1] = accel
    p.hor[0], p.hor[1] = hor
    p.ver[0], p.ver[1] = ver
    p.refl_L, p.refl_R, p.refl_U, p.refl_D = refl
    p.kernel, p.precision, p.device, p.flags = kernel, precision, device, flags
    return p


def _ptr(a):
    return None if a is None else C.c_void_p(a.ctypes.data)


class Handle:
    """Thin RAII wrapper of an sphb_sim*; every method is one C-ABI call."""

    def __init__(self, params: Params, pos, vel=None, e=None, rho=None, ids=None, capacity=None):
        L = lib()
        pos = np.ascontiguousarray(pos, dtype=np.float64).reshape(-1, 2)
        n = pos.shape[0]
        vel = None if vel is None else np.ascontiguousarray(vel, dtype=np.float64).reshape(n, 2)
        e = None if e is None else np.ascontiguousarray(e, dtype=np.float64).reshape(n)
        rho = None if rho is None else np.ascontiguousarray(rho, dtype=np.float64).reshape(n)
        ids = None if ids is None else np.ascontiguousarray(ids, dtype=np.int64).reshape(n)
        self._h = C.c_void_p()
        rc = L.sphb_create(C.byref(params), n, capacity or n, _ptr(pos), _ptr(vel), _ptr(e), _ptr(rho), _ptr(ids),
                           C.byref(self._h))
        if rc:
            raise SphbError(rc, L.sphb_last_error(None).decode())

    @classmethod
    def from_device(cls, params: Params, n, d_pos, d_vel=None, d_e=None, d_id=None, capacity=None):
        """d_* are CUDA device pointers (ints) on params.device."""
        L = lib()
        self = cls.__new__(cls)
        self._h = C.c_void_p()
        rc = L.sphb_create_device(C.byref(params), n, capacity or n, C.c_void_p(d_pos),
                                  C.c_void_p(d_vel) if d_vel else None, C.c_void_p(d_e) if d_e else None,
                                  C.c_void_p(d_id) if d_id else None, C.byref(self._h))
        if rc:
            raise SphbError(rc, L.sphb_last_error(None).decode())
        return self

    def close(self):
        if getattr(self, "_h", None):
            lib().sphb_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _chk(self, rc):
        if rc:
            raise SphbError(rc, lib().sphb_last_error(self._h).decode())

    # -- one method per ABI entry point
    @property
    def n(self):
        return int(lib().sphb_count(self._h))

    @property
    def current_step(self):
        return int(lib().sphb_current_step(self._h))

    def set_params(self, p: Params):
        self._chk(lib().sphb_set_params(self._h, C.byref(p)))

    def get_params(self) -> Params:
        p = Params()
        self._chk(lib().sphb_get_params(self._h, C.byref(p)))
        return p

    def append(self, pos, vel=None, e=None, rho=None, ids=None):
        pos = np.ascontiguousarray(pos, dtype=np.float64).reshape(-1, 2)
        n = pos.shape[0]
        vel = None if vel is None else np.ascontiguousarray(vel, dtype=np.float64).reshape(n, 2)
        e = None if e is None else np.ascontiguousarray(e, dtype=np.float64).reshape(n)
        rho = None if rho is None else np.ascontiguousarray(rho, dtype=np.float64).reshape(n)
        ids = None if ids is None else np.ascontiguousarray(ids, dtype=np.int64).reshape(n)
        self._chk(lib().sphb_append(self._h, n, _ptr(pos), _ptr(vel), _ptr(e), _ptr(rho), _ptr(ids)))

    def step(self, nsteps=1):
        self._chk(lib().sphb_step(self._h, nsteps))

    def calc_forces(self):
        self._chk(lib().sphb_calc_forces(self._h))

    def knn(self, hor=OPEN, ver=OPEN):
        h = (C.c_double * 2)(*hor)
        v = (C.c_double * 2)(*ver)
        self._chk(lib().sphb_knn(self._h, h, v))

    def density(self, kernel):
        self._chk(lib().sphb_density(self._h, kernel))

    @property
    def stream(self):
        """cudaStream_t of the handle as an int (torch.cuda.ExternalStream(handle.stream))"""
        return int(lib().sphb_stream(self._h) or 0)

    def sync(self):
        self._chk(lib().sphb_sync(self._h))

    def download(self, fields, out=None):
        """dict name -> array in current device order. `out` may hold preallocated (e.g. pinned) arrays."""
        n = self.n
        arrs, ptrs, mask = {}, (C.c_void_p * len(FIELDS))(), 0
        for name in fields:
            shp, dt = FIELD_SHAPE[name]
            a = out[name] if out is not None and name in out else np.empty((n,) + shp, dtype=dt)
            arrs[name] = a
            ptrs[FIELD_BIT[name]] = a.ctypes.data
            mask |= 1 << FIELD_BIT[name]
        nout = C.c_int64()
        self._chk(lib().sphb_download(self._h, mask, ptrs, n, C.byref(nout)))
        return arrs

    def upload(self, **fields):
        n = self.n
        ptrs, mask, keep = (C.c_void_p * len(FIELDS))(), 0, []
        for name, a in fields.items():
            shp, dt = FIELD_SHAPE[name]
            a = np.ascontiguousarray(a, dtype=dt).reshape((n,) + shp)
            keep.append(a)
            ptrs[FIELD_BIT[name]] = a.ctypes.data
            mask |= 1 << FIELD_BIT[name]
        self._chk(lib().sphb_upload(self._h, mask, ptrs, n))

    def upload_by_id(self, **fields):
        """pos / vel / e arrays indexed by particle id (dense ids): the caller's own fixed order"""
        n = self.n
        ptrs, mask, keep = (C.c_void_p * len(FIELDS))(), 0, []
        for name, a in fields.items():
            shp, dt = FIELD_SHAPE[name]
            a = np.ascontiguousarray(a, dtype=dt).reshape((n,) + shp)
            keep.append(a)
            ptrs[FIELD_BIT[name]] = a.ctypes.data
            mask |= 1 << FIELD_BIT[name]
        self._chk(lib().sphb_upload_by_id(self._h, mask, ptrs, n))

    def upload_by_id_begin(self, **fields):
        """first half of upload_by_id: starts the copies and returns; the arrays must stay alive (and unchanged) until
        upload_by_id_end"""
        n = self.n
        ptrs, mask, keep = (C.c_void_p * len(FIELDS))(), 0, []
        for name, a in fields.items():
            shp, dt = FIELD_SHAPE[name]
            a = np.ascontiguousarray(a, dtype=dt).reshape((n,) + shp)
            keep.append(a)
            ptrs[FIELD_BIT[name]] = a.ctypes.data
            mask |= 1 << FIELD_BIT[name]
        self._up_keep = keep
        self._chk(lib().sphb_upload_by_id_begin(self._h, mask, ptrs, n))

    def upload_by_id_end(self):
        self._chk(lib().sphb_upload_by_id_end(self._h))
        self._up_keep = None

    def reduce(self, which):
        out = C.c_double()
        self._chk(lib().sphb_reduce(self._h, which, C.byref(out)))
        return out.value

    def frame(self, width=1280, height=720, ids=True, out=None):
        """frame data of (*Animator).CurrentFrame (animator.go:75-101): dict xy float32 [n, 2], colour uint8 [n],
        id int64 [n] in current device order; `out` may hold preallocated (pinned) arrays"""
        n = self.n
        out = out or {}
        xy = out.get("xy") if out.get("xy") is not None else np.empty((n, 2), np.float32)
        col = out.get("colour") if out.get("colour") is not None else np.empty(n, np.uint8)
        idv = (out.get("id") if out.get("id") is not None else np.empty(n, np.int64)) if ids else None
        nout = C.c_int64()
        self._chk(lib().sphb_frame(self._h, width, height, xy.ctypes.data, col.ctypes.data,
                                   idv.ctypes.data if ids else None, n, C.byref(nout)))
        return {"xy": xy, "colour": col, "id": idv}

    def max_h(self):
        out = C.c_double()
        self._chk(lib().sphb_max_h(self._h, C.byref(out)))
        return out.value

    def max_speed(self):
        out = C.c_double()
        self._chk(lib().sphb_max_speed(self._h, C.byref(out)))
        return out.value

    def phase_times(self):
        ms = (C.c_double * len(PHASES))()
        self._chk(lib().sphb_phase_times(self._h, ms, len(PHASES)))
        return dict(zip(PHASES, list(ms)))

    # ---- the slab ring inside the library (include/sphb.h "slab ring")
    def ring_set(self, rank, nranks, periodic, x_lo, x_hi, h_hint, halo_cap, safety=0.0, migrate_every=0):
        r = Ring(rank, nranks, int(bool(periodic)), int(migrate_every), x_lo, x_hi, h_hint, safety, int(halo_cap))
        self._chk(lib().sphb_ring_set(self._h, C.byref(r)))

    def comm_init(self, unique_id: bytes, rank: int, nranks: int):
        buf = C.create_string_buffer(bytes(unique_id), NCCL_ID_BYTES)
        self._chk(lib().sphb_comm_init(self._h, buf, rank, nranks))

    def ring_step(self, nsteps=1):
        self._chk(lib().sphb_ring_step(self._h, nsteps))

    def ring_info(self):
        out = (C.c_double * len(RING_INFO))()
        self._chk(lib().sphb_ring_info(self._h, out, len(RING_INFO)))
        return dict(zip(RING_INFO, [float(x) for x in out]))

    def counters(self):
        out = (C.c_int64 * len(COUNTERS))()
        self._chk(lib().sphb_counters(self._h, out, len(COUNTERS)))
        return dict(zip(COUNTERS, [int(x) for x in out]))

    def state(self, fields=("pos", "vel", "rho", "c", "e", "edot", "vdot", "epred", "vpred", "h", "id"),
              sort_by_id=True):
        d = self.download(list(dict.fromkeys(list(fields) + ["id"])))
        if "nn_idx" in d:  # neighbour indices -> neighbour ids (order independent)
            idx = d["nn_idx"].astype(np.int64)
            d["nn_id"] = np.where(idx >= 0, d["id"][np.clip(idx, 0, None)], -1)
        if sort_by_id:
            o = np.argsort(d["id"], kind="stable")
            d = {k: v[o] for k, v in d.items()}
        return d


def comm_unique_id() -> bytes:
    """ncclGetUniqueId through the library (rank 0 calls it and distributes the 128 bytes)"""
    buf = C.create_string_buffer(NCCL_ID_BYTES)
    rc = lib().sphb_comm_unique_id(buf)
    if rc:
        raise SphbError(rc, lib().sphb_last_error(None).decode())
    return buf.raw


def ring_step_local(handles, nsteps=1):
    """Step() of a ring whose slabs all live in this process (handles[k] = rank k)"""
    arr = (C.c_void_p * len(handles))(*[h._h for h in handles])
    rc = lib().sphb_ring_step_local(arr, len(handles), nsteps)
    if rc:
        for h in handles:
            msg = lib().sphb_last_error(h._h).decode()
            if msg:
                raise SphbError(rc, msg)
        raise SphbError(rc, "sphb_ring_step_local failed")
