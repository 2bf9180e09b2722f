#!/usr/bin/env python
"""CPU study for a round-2 kernel: exact kNN from the PREVIOUS neighbour list with an exactness certificate.

Today k_knn_tile scans ~105 staged candidates per particle and selects 32 of ~34 survivors every step (8.4 of 16.1 ms at
2^25 particles, issue-bound).  Between two steps the neighbour sets barely change.  Idea: when the full search runs, keep
the K_ext > 32 nearest per particle and the distance d_excl of the first one NOT kept.  At a later step evaluate only those
K_ext candidates at their new positions, take h' = the 32nd smallest distance, and accept if

    h' < d_excl - D,     D >= |disp_i - disp_j| for every pair since the list was built

(every particle outside the list was at least d_excl away and can have come closer by at most D), otherwise fall back to
the full search.  D comes from a global reduction: 2 * max_k |disp_k - mean disp| accumulated over the steps.  Accepted
results are EXACT kNN (the reference's contract), so parity is unaffected; only the fallback rate decides the speed-up.

This script measures that rate on bench.py's workload in scaled units (jittered lattice, dt_half = spacing, E = 0.01,
g = (0, 0.2): the Euler equations are scale-free, so 128^2 particles behave like the 16384^2 of C5), stepping with the CPU
oracle.  Output: per (K_ext, steps since rebuild) the fraction of particles and of 32-particle tiles (cell order) certified.

    python tools/verlet_knn_study.py [--n 128] [--steps 24] [--kext 36,40,48]
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
from scipy.spatial import cKDTree  # noqa: E402

from oracle import oracle as orc  # noqa: E402
from sphugo_b200 import gen  # noqa: E402


def min_image(d, L=1.0):
    return d - L * np.round(d / L)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=128)
    ap.add_argument("--steps", type=int, default=24)
    ap.add_argument("--kext", default="36,40,48")
    ap.add_argument("--rebuild-at", type=int, default=4, help="step whose evaluation builds the lists (after warm-up)")
    a = ap.parse_args()
    n, s = a.n, 1.0 / a.n
    pos = gen.jittered_lattice(n, n)
    N = len(pos)
    dtH = 0.98 * s  # C5: 6e-5 at spacing 2^-14 = 6.1e-5
    o = orc.Oracle(orc.make_params(hor=(0.0, 1.0), ver=(0.0, 1.0), accel=(0.0, 0.2), dt_half=dtH), pos, None, np.full(N, 0.01))
    kexts = [int(k) for k in a.kext.split(",")]
    kmax = max(kexts) + 2
    evals = []  # positions at the force evaluation of each step (after drift-1), unwrapped increments tracked separately
    st = o.state()
    for step in range(a.steps):
        ev = st["pos"] + st["vel"] * dtH  # drift-1 of the coming step (sph.go:108-112)
        evals.append(ev.copy())
        o.step(1, 1)
        st = o.state()
    t0 = a.rebuild_at
    p0 = np.mod(evals[t0], 1.0)
    d0, j0 = cKDTree(p0, boxsize=1.0).query(p0, k=kmax + 1)
    d0, j0 = d0[:, 1:], j0[:, 1:]  # drop self
    # tile = 32 consecutive particles in cell order (row-major cells of height 1.15 h, width 0.32 * that)
    h_mean = d0[:, 31].mean()
    dy = 1.15 * h_mean
    dx = 0.32 * dy
    key = np.floor(p0[:, 1] / dy).astype(np.int64) * int(np.ceil(1.0 / dx)) + np.floor(p0[:, 0] / dx).astype(np.int64)
    order = np.argsort(key, kind="stable")
    tile_of = np.empty(N, dtype=np.int64)
    tile_of[order] = np.arange(N) // 32
    ntiles = tile_of.max() + 1
    print(f"N = {N}, spacing {s:.4g}, dt_half {dtH:.4g}, mean h {h_mean:.4g} = {h_mean / s:.3f} spacings; lists built at step {t0}")
    print("steps since rebuild | D / h | " + " | ".join(f"K_ext={k}: particles, tiles, particles with the local bound" for k in kexts))
    # local bound: D_i = largest |cum_j - cum_i| over the 3x3 block of coarse cells (edge >= 2 h) around particle i, from
    # per-cell bounding boxes of the cumulative displacement (valid while d_excl < one coarse cell)
    nc = int(1.0 / (2.0 * h_mean))
    cs = 1.0 / nc
    cid = np.minimum((p0[:, 1] / cs).astype(int), nc - 1) * nc + np.minimum((p0[:, 0] / cs).astype(int), nc - 1)
    cum = np.zeros((N, 2))
    D = 0.0
    for t in range(t0 + 1, a.steps):
        disp = min_image(evals[t] - evals[t - 1])
        dev = disp - disp.mean(0)
        D += 2.0 * np.sqrt((dev ** 2).sum(1)).max()
        cum += disp
        mn, mx = np.full((nc * nc, 2), np.inf), np.full((nc * nc, 2), -np.inf)
        np.minimum.at(mn, cid, cum)
        np.maximum.at(mx, cid, cum)
        mn, mx = mn.reshape(nc, nc, 2), mx.reshape(nc, nc, 2)
        bmn, bmx = mn.copy(), mx.copy()
        for sy in (-1, 0, 1):
            for sx in (-1, 0, 1):
                bmn = np.minimum(bmn, np.roll(np.roll(mn, sy, 0), sx, 1))
                bmx = np.maximum(bmx, np.roll(np.roll(mx, sy, 0), sx, 1))
        ext = np.maximum(np.abs(cum - bmn.reshape(-1, 2)[cid]), np.abs(bmx.reshape(-1, 2)[cid] - cum))
        d_loc = np.sqrt((ext ** 2).sum(1))
        pt = evals[t]
        row = [f"{t - t0:19d}", f"{D / h_mean:.4f}"]
        # exact answer for reference: the true 32nd-neighbour distance now
        ptm = np.mod(pt, 1.0)
        dtrue, _ = cKDTree(ptm, boxsize=1.0).query(ptm, k=33)
        for k in kexts:
            cand = j0[:, :k]
            dvec = min_image(pt[cand] - pt[:, None, :])
            dn = np.sqrt((dvec ** 2).sum(-1))
            hnew = np.sort(dn, axis=1)[:, 31]
            ok = hnew < d0[:, k] - D  # d0[:, k] = the first neighbour not kept
            assert np.allclose(hnew[ok], dtrue[ok, 32], rtol=1e-12), "certified result is not the exact kNN"
            tiles_ok = np.bincount(tile_of, weights=(~ok).astype(float), minlength=ntiles) == 0
            ok_loc = (hnew < d0[:, k] - d_loc) & (d0[:, k] < cs)
            assert np.allclose(hnew[ok_loc], dtrue[ok_loc, 32], rtol=1e-12), "locally certified result is not the exact kNN"
            row.append(f"{ok.mean():.4f}, {tiles_ok.mean():.4f}, {ok_loc.mean():.4f}")
        print(" | ".join(row))


if __name__ == "__main__":
    main()
