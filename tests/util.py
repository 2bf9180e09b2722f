"""Shared helpers of the parity tests: workloads, oracle/GPU runners and error measures."""
from __future__ import annotations

import numpy as np

from oracle import oracle as orc
from sphugo_b200 import _lib as L
from sphugo_b200 import gen

OPEN = L.OPEN
TOL64 = 1e-12  # north_star: fp64 build relative tolerance


def params_pair(**kw):
    """the same configuration for the oracle and for libsphb"""
    return orc.make_params(**kw), L.make_params(**kw)


def c1_density(n_bg=1000, n_blob=200):
    """examples/density main (density.go:50-63): 1000 U([0,1]^2) + 200 U([0.1,0.3]x[0,0.4])."""
    return gen.spawn([(n_bg, (0.0, 0.0), (1.0, 1.0)), (n_blob, (0.1, 0.0), (0.3, 0.4))])


def c1_periodic_visual():
    """periodicVisualTest (density.go:99-141): 10000 U([0.1,0.9]^2) + 1200 U([0.85,0.9]x[0.4,0.9])."""
    return gen.spawn([(10000, (0.1, 0.1), (0.9, 0.9)), (1200, (0.85, 0.4), (0.9, 0.9))])


def by_id(d):
    o = np.argsort(d["id"], kind="stable")
    return {k: v[o] for k, v in d.items()}


def rel_err(a, b, scale=None):
    """max |a-b| / max(|b|, scale) (scale broadcast per particle)"""
    a, b = np.asarray(a, float), np.asarray(b, float)
    den = np.abs(b)
    if scale is not None:
        s = np.asarray(scale, float)
        while s.ndim < den.ndim:
            s = s[..., None]
        den = np.maximum(den, s)
    den = np.maximum(den, 1e-300)
    return float(np.max(np.abs(a - b) / den)) if a.size else 0.0


def neighbour_sets_equal(gpu, ref, tie_rel=0.0):
    """Compare neighbour id multisets per particle (both sorted by id).

    Exact distance ties at the k-th place are excluded (north_star): a mismatch is tolerated only if the
    differing neighbours sit at exactly the same distance as h in both lists (|d - h| <= tie_rel * h).
    Returns (n_mismatch_particles, n_tie_excused)."""
    g = np.sort(gpu["nn_id"], axis=1)
    r = np.sort(ref["nn_id"], axis=1)
    bad = np.nonzero((g != r).any(axis=1))[0]
    excused = 0
    hard = 0
    for i in bad:
        gs, rs = set(gpu["nn_id"][i].tolist()), set(ref["nn_id"][i].tolist())
        dg = gpu["nn_dist"][i][[k for k, v in enumerate(gpu["nn_id"][i]) if v not in rs]]
        dr = ref["nn_dist"][i][[k for k, v in enumerate(ref["nn_id"][i]) if v not in gs]]
        h = ref["h"][i]
        if len(dg) and len(dr) and np.all(np.abs(dg - h) <= tie_rel * h) and np.all(np.abs(dr - h) <= tie_rel * h):
            excused += 1
        else:
            hard += 1
    return hard, excused


def force_scales(ref, prm):
    """Per-particle magnitude sums of the terms of AccelerationAndEDot2D (sph.go:327-401) from an oracle
    state with neighbours (sorted by id): the 32-term sums nearly cancel on near-uniform data, so errors
    are measured relative to sum |term| (SURVEY §7 hard parts)."""
    from oracle.oracle import NN  # noqa
    ids = ref["id"]
    pos_of = {int(z): k for k, z in enumerate(ids)}
    nn = np.vectorize(lambda z: pos_of.get(int(z), 0))(ref["nn_id"])
    valid = ref["nn_id"] >= 0
    gamma, m = prm.gamma, prm.particle_mass
    A = ref["c"] ** 2 / (gamma * ref["rho"])
    B = A[nn]
    h = ref["h"]
    d = ref["nn_dist"]
    q = d / h[:, None]
    if prm.kernel == 1:
        df = np.where(q < 0.5, 3 * q * q - 2 * q, -(1 - q) ** 2)
        pref = 6 * 40 / (np.pi * 7)
    else:
        df = -10 * q * (1 - q) ** 3
        pref = 8 * 7 / (np.pi * 4)
    rab = ref["nn_pos"] - ref["pos"][:, None, :]
    vab = ref["vpred"][nn] - ref["vpred"][:, None, :]
    dot = (rab * vab).sum(-1)
    cab = 0.5 * (ref["c"][:, None] + ref["c"][nn])
    rhoab = 0.5 * (ref["rho"][:, None] + ref["rho"][nn])
    hab = 0.5 * (h[:, None] + h[nn])
    mu = dot * hab / ((rab ** 2).sum(-1) + 0.01)
    pi = np.where(dot < 0, (np.abs(0.75 * cab * mu) + 1.5 * mu * mu) / rhoab, 0.0)
    w = (pi + A[:, None] + B) * np.abs(df) / np.maximum(d, 1e-300) * valid
    f = m * pref / h ** 3
    acc_scale = (np.abs(rab) * w[..., None]).sum(1) * f[:, None] + np.abs(np.array([prm.accel[0], prm.accel[1]]))
    edot_scale = np.abs(A) * (np.abs(dot) * np.abs(df) * valid).sum(1) * m
    return acc_scale, edot_scale


def run_oracle(po, ic, steps=0, knn_mode=0, forces=False, neighbours=True):
    o = orc.Oracle(po, ic["pos"], ic.get("vel"), ic.get("e"), ic.get("rho"), ic.get("id"))
    if forces:
        # CalculateForces on a fresh simulation needs VPred/EPred like step 0 sets them (sph.go:97-100);
        # orc_step_mode does that itself, so a plain force call is only used after steps.
        o.calc_forces(knn_mode)
    if steps:
        o.step(steps, knn_mode)
    return o


FIELDS_STATE = ["pos", "vel", "rho", "c", "e", "edot", "vdot", "epred", "vpred", "h", "id"]
FIELDS_NN = ["nn_idx", "nn_dist", "nn_pos"]
