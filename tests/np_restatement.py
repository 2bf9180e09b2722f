"""A second, independent restatement of the reference's step in numpy, used ONLY to cross-check oracle/sph_oracle.c.

The reference has no test for kNN, density, force or the integrator, and Go cannot run here; the C oracle is a
line-by-line restatement with the reference's tree walk.  This file restates the same Go functions a second time, from
the Go source and not from the C: brute-force kNN over the periodic images (no tree), every particle at once, each
floating-point operation in the order the Go expression tree gives it (numpy rounds every elementwise operation
separately, like a GOAMD64=v1 build without FMA), the 32-slot sums run slot by slot in the list's order.  Where the two
restatements agree BIT FOR BIT (tests/test_oracle_crosscheck.py), a transcription mistake would have to be made twice.

    kNN      sim/nearest-neighbour.go:28-67, 70-83, 139-153   (exact: sorted by descending d^2; sentinel cap d^2 < 0.4 asserted)
    density  sim/sph.go:306-323, kernels :244-304
    force    sim/sph.go:327-401, sound speed :423-429
    step     sim/sph.go:89-193
"""
from __future__ import annotations

from decimal import Decimal, getcontext

import numpy as np

MAXF = 1.7976931348623157e308
NN = 32
getcontext().prec = 60
_PI = Decimal("3.14159265358979323846264338327950288419716939937510582097494459")  # Go's math.Pi literal
# Go untyped-constant expressions are evaluated exactly and rounded once (sph.go:249,261,266,275,293,303)
TOPHAT_F = float(Decimal(1) / _PI)
MONAGHAN = float(Decimal(6 * 40) / (_PI * 7))
WEND_F = float(Decimal(4 * 7) / (_PI * 4))
WEND_DF = float(Decimal(8 * 7) / (_PI * 4))
SIXTH = float(Decimal(1) / Decimal(6))  # the constant 1.0/6


def F(kernel, q):
    if kernel == 0:
        return np.ones_like(q)
    if kernel == 1:
        a = q * q * q - q * q + SIXTH
        b = (1 - q) * (1 - q) * (1 - q) / 3
        return np.where(q < 0.5, a, b)
    return (1 - q) * (1 - q) * (1 - q) * (1 - q) * (1 + 4 * q)


def DF(kernel, q):
    if kernel == 1:
        return np.where(q < 0.5, 3 * q * q - 2 * q, -(1 - q) * (1 - q))
    if kernel == 2:
        return -10 * q * (1 - q) * (1 - q) * (1 - q)
    raise RuntimeError("not defined. derivative is delta distribution!")


def prefactors(kernel):
    return {0: (TOPHAT_F, 1.0), 1: (MONAGHAN, MONAGHAN), 2: (WEND_F, WEND_DF)}[kernel]


def knn(pos, hor, ver):
    """-> idx [N,32], d [N,32] (sqrt-ed), npos [N,32,2]; slot 0 = farthest"""
    n = len(pos)
    irange, dX = ([0], 0.0) if hor[0] == -MAXF else ([-1, 0, 1], hor[1] - hor[0])
    jrange, dY = ([0], 0.0) if ver[0] == -MAXF else ([-1, 0, 1], ver[1] - ver[0])
    d2s, idxs, npos = [], [], []
    for i in irange:
        for j in jrange:
            ox, oy = float(i) * dX, float(j) * dY
            qx, qy = pos[:, 0] + ox, pos[:, 1] + oy  # pos := particle.Pos.Add(&offset)
            dx, dy = qx[:, None] - pos[None, :, 0], qy[:, None] - pos[None, :, 1]
            d2 = dx * dx + dy * dy
            np.fill_diagonal(d2, np.inf)  # particle != &root.Particles[i], in every image
            d2s.append(d2)
            idxs.append(np.broadcast_to(np.arange(n), (n, n)))
            npos.append(np.broadcast_to(np.stack([pos[:, 0] - ox, pos[:, 1] - oy], 1), (n, n, 2)))  # b.Pos.Sub(&offset)
    d2, idx, npos = np.concatenate(d2s, 1), np.concatenate(idxs, 1), np.concatenate(npos, 1)
    sel = np.argsort(d2, axis=1, kind="stable")[:, :NN][:, ::-1]  # descending distance, slot 0 = farthest
    rows = np.arange(n)[:, None]
    d2 = d2[rows, sel]
    assert d2[:, 0].max() < 0.4, "the reference's sentinel cap would bind (nearest-neighbour.go:155-165)"
    return idx[rows, sel], np.sqrt(d2), npos[rows, sel]


def density(d, kernel, mass):
    maxR = d[:, 0]
    acc = np.zeros(len(d))
    for i in range(NN):
        acc = acc + F(kernel, d[:, i] / maxR)
    return prefactors(kernel)[0] * mass * acc / (maxR * maxR)


def forces(st, cfg):
    """CalculateForces (sph.go:403-435) on the state dict; fills rho, c, h, vdot, edot"""
    pos, kernel, gamma, mass = st["pos"], cfg["kernel"], cfg["gamma"], cfg["particle_mass"]
    idx, d, npos = knn(pos, cfg["hor"], cfg["ver"])
    rho = density(d, kernel, mass)
    c = np.sqrt(gamma * (gamma - 1) * st["epred"])
    maxR = d[:, 0]
    A = c * c / (gamma * rho)
    ax, ay, ae = np.zeros(len(pos)), np.zeros(len(pos)), np.zeros(len(pos))
    for i in range(NN):
        nb = idx[:, i]
        q = d[:, i] / maxR
        dR = DF(kernel, q)
        B = c[nb] * c[nb] / (gamma * rho[nb])
        vx, vy = st["vpred"][nb, 0] - st["vpred"][:, 0], st["vpred"][nb, 1] - st["vpred"][:, 1]
        rx, ry = npos[:, i, 0] - pos[:, 0], npos[:, i, 1] - pos[:, 1]
        dot = vx * rx + vy * ry
        cAB, rhoAB, hAB = 0.5 * (c + c[nb]), 0.5 * (rho + rho[nb]), 0.5 * (maxR + maxR[nb])
        mu = dot * hAB / ((rx * rx + ry * ry) + 0.01)
        pi = np.where(dot < 0, (-0.75 * cAB * mu + 1.5 * mu * mu) / rhoAB, 0.0)
        ax = ax + rx * (pi + A + B) * dR / d[:, i]
        ay = ay + ry * (pi + A + B) * dR / d[:, i]
        ae = ae + dot * dR
    f = mass * prefactors(kernel)[1] / (maxR * maxR * maxR)
    st.update(rho=rho, c=c, h=maxR, vdot=np.stack([ax * f + cfg["accel"][0], ay * f + cfg["accel"][1]], 1), edot=A * ae * mass)


def make_state(pos, vel=None, e=None):
    n = len(pos)
    z2, z1 = np.zeros((n, 2)), np.zeros(n)
    return dict(pos=np.array(pos, float), vel=z2.copy() if vel is None else np.array(vel, float),
                e=z1.copy() if e is None else np.array(e, float), vdot=z2.copy(), edot=z1.copy(), vpred=z2.copy(),
                epred=z1.copy(), step=0)


def step(st, cfg):
    """(*Simulation).Step, sph.go:64-198 (no sources)"""
    dtH = cfg["dt_half"]
    if st["step"] == 0:  # sph.go:89-103
        st["vpred"], st["epred"] = st["vel"].copy(), st["e"].copy()
        forces(st, cfg)
    st["pos"] = st["pos"] + st["vel"] * dtH
    st["vpred"] = st["vel"] + st["vdot"] * dtH
    st["epred"] = st["e"] + st["edot"] * dtH
    forces(st, cfg)
    st["vel"] = st["vel"] + st["vdot"] * (2 * dtH)
    st["e"] = st["e"] + st["edot"] * 2 * dtH
    st["pos"] = st["pos"] + st["vel"] * dtH
    x, y = st["pos"][:, 0].copy(), st["pos"][:, 1].copy()
    (h0, h1), (v0, v1) = cfg["hor"], cfg["ver"]
    with np.errstate(over="ignore"):  # open axes: the period overflows but no particle is beyond +-MaxFloat64
        lx, ly = h1 - h0, v1 - v0
        c1 = x < h0  # sph.go:147-167: one shift at most, each `continue` skips the later tests
        c2 = ~c1 & (x > h1)
        c3 = ~c1 & ~c2 & (y < v0)
        c4 = ~c1 & ~c2 & ~c3 & (y > v1)
        x = np.where(c1, x + lx, np.where(c2, x - lx, x))
        y = np.where(c3, y + ly, np.where(c4, y - ly, y))
    vx, vy = st["vel"][:, 0].copy(), st["vel"][:, 1].copy()
    rL, rR, rU, rD = cfg["refl"]
    m = x < rL; x = np.where(m, x - (x - rL), x); vx = np.where(m, -vx, vx)  # noqa: E702  sph.go:170-193, in this order
    m = x > rR; x = np.where(m, x - (x - rR), x); vx = np.where(m, -vx, vx)  # noqa: E702
    m = y < rU; y = np.where(m, y - (y - rU), y); vy = np.where(m, -vy, vy)  # noqa: E702
    m = y > rD; y = np.where(m, y - (y - rD), y); vy = np.where(m, -vy, vy)  # noqa: E702
    st["pos"], st["vel"] = np.stack([x, y], 1), np.stack([vx, vy], 1)
    st["step"] += 1
