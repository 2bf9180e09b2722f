"""sphugo_b200 — B200-native (sm_100a) replacement for the per-step SPH hot path of bbeni/sphugo.

Product = libsphb.so (CUDA kernels + C ABI, include/sphb.h).  This package only holds the build recipe,
the ctypes binding of that ABI and a host-side mirror of the reference's Go `sim` API (sim.py) so that
parity tests read like the reference's own call sites.  Nothing here computes on the CPU.
"""
from . import _lib  # noqa: F401
from ._lib import Handle, Params, SphbError, make_params, OPEN  # noqa: F401
