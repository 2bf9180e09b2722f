"""Synthetic initial conditions.

The reference draws its particles from Go's math/rand (config-parser.go:58-80); gorand.py reconstructs that
stream bit for bit and the mirrored spawners of sim.py use it.  The benchmark workloads and most parity cases
(large N, lattices, shock tubes: shapes the reference has no generator for) use this documented, vectorised
generator instead; initial conditions always cross the boundary as explicit arrays either way:

    splitmix64, state0 = seed;  u = (next() >> 11) * 2**-53  in [0, 1)

Per particle x is drawn first, then y (the order of config-parser.go:68-72).  Like the reference, which
re-seeds on every Spawn, each rectangle restarts the stream, so rectangles share the same uniforms.
"""
from __future__ import annotations

import numpy as np

DEFAULT_SEED = 12345678  # config-parser.go:63
_GOLD = np.uint64(0x9E3779B97F4A7C15)
_M1 = np.uint64(0xBF58476D1CE4E5B9)
_M2 = np.uint64(0x94D049BB133111EB)


def splitmix64(seed: int, count: int, offset: int = 0) -> np.ndarray:
    """outputs offset+1 .. offset+count of splitmix64 seeded with `seed` (uint64)."""
    with np.errstate(over="ignore"):
        k = np.arange(offset + 1, offset + count + 1, dtype=np.uint64)
        z = np.uint64(seed & 0xFFFFFFFFFFFFFFFF) + k * _GOLD
        z = (z ^ (z >> np.uint64(30))) * _M1
        z = (z ^ (z >> np.uint64(27))) * _M2
        return z ^ (z >> np.uint64(31))


def uniform01(seed: int, count: int, offset: int = 0) -> np.ndarray:
    return (splitmix64(seed, count, offset) >> np.uint64(11)).astype(np.float64) * (2.0 ** -53)


def uniform_rect(n: int, upper_left=(0.0, 0.0), lower_right=(1.0, 1.0), seed: int = DEFAULT_SEED) -> np.ndarray:
    """UniformRectSpawner.Spawn (config-parser.go:58-80): n points, x then y per particle."""
    u = uniform01(seed, 2 * n).reshape(n, 2)
    ul, lr = np.asarray(upper_left, float), np.asarray(lower_right, float)
    return ul + u * (lr - ul)


def spawn(rects, seed: int = DEFAULT_SEED, e0: float = 0.01):
    """Start spawners of a config: list of (n, upper_left, lower_right) -> dict(pos, vel, e, id).
    Spawner defaults: E = 0.01, everything else zero (config-parser.go:74-77); id = running index."""
    pos = np.concatenate([uniform_rect(n, ul, lr, seed) for (n, ul, lr) in rects], axis=0)
    n = pos.shape[0]
    return dict(pos=pos, vel=np.zeros((n, 2)), e=np.full(n, e0), id=np.arange(n, dtype=np.int64))


def jittered_lattice(nx: int, ny: int, lo=(0.0, 0.0), hi=(1.0, 1.0), jitter: float = 0.25,
                     seed: int = DEFAULT_SEED, offset: int = 0) -> np.ndarray:
    """nx x ny lattice over [lo, hi), each coordinate displaced by U(-jitter, jitter) * spacing (C3-P / C5)."""
    lo, hi = np.asarray(lo, float), np.asarray(hi, float)
    sx, sy = (hi - lo) / np.array([nx, ny], float)
    iy, ix = np.divmod(np.arange(nx * ny, dtype=np.int64) + offset, nx)
    u = uniform01(seed, 2 * nx * ny, 2 * offset).reshape(-1, 2)
    x = lo[0] + (ix + 0.5 + (2.0 * u[:, 0] - 1.0) * jitter) * sx
    y = lo[1] + (iy + 0.5 + (2.0 * u[:, 1] - 1.0) * jitter) * sy
    return np.stack([x, y], axis=1)


def shock_tube(n_total: int, ratio: int = 4, seed: int = DEFAULT_SEED) -> np.ndarray:
    """C4: periodic [0,1]^2, jittered lattices with number-density ratio `ratio`:1 left/right of x = 0.5."""
    n_left = int(round(n_total * ratio / (ratio + 1.0)))
    n_right = n_total - n_left

    def half(n, x0, x1, sd):
        ny = max(1, int(round(np.sqrt(n / (x1 - x0)))))
        nx = max(1, int(np.ceil(n / ny)))
        p = jittered_lattice(nx, ny, (x0, 0.0), (x1, 1.0), 0.25, sd)
        return p[:n]

    return np.concatenate([half(n_left, 0.0, 0.5, seed), half(n_right, 0.5, 1.0, seed + 1)], axis=0)


def dam_break(n_total: int, seed: int = DEFAULT_SEED) -> np.ndarray:
    """C4 variant: column [0, 0.25] x [0.5, 1] filled with a jittered lattice (y points down)."""
    ny = max(1, int(round(np.sqrt(n_total * 2.0))))
    nx = max(1, int(np.ceil(n_total / ny)))
    return jittered_lattice(nx, ny, (0.0, 0.5), (0.25, 1.0), 0.25, seed)[:n_total]
