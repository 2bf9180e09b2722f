"""Go `math/rand` (the Go 1 additive lagged-Fibonacci source), bit for bit, for the reference's initial conditions.

Every particle the reference spawns comes from the top-level functions of Go's math/rand after an explicit
`rand.Seed(12345678)` (config-parser.go:58-80 `UniformRectSpawner.Spawn`, core.go:76-91 `InitUniformly`,
config-parser.go:82-102 `PointSource.Spawn`; `examples/heap/heap.go:27` seeds 101).  An explicit Seed selects the
Go 1 source whatever the Go version (go.mod:3 pins 1.22.2), so reproducing that source reproduces the scenes of
`examples/density`, `examples/sph-simulation`, `examples/speed-test` and the generated `.sph-config` files particle
for particle (SURVEY §8c "third-party arithmetic", §8f-4).

The generator is x[n] = x[n-607] + x[n-273] mod 2^64.  Its seeding XORs a 607-entry table (`rngCooked` in Go's
rng.go) into an LCG-filled vector; Go's sources are not in this container, so the table is *recomputed* here from
its published definition (Go's gen_cooked.go: fill the vector from the same LCG with seed 1, run the generator
7.8e12 times, print the vector).  7.8e12 steps are taken by jump-ahead: the recurrence is linear over Z/2^64, so
the state after n steps is x^n modulo the characteristic polynomial x^607 - x^334 - 1 applied to the initial
sequence (43 polynomial squarings, 0.2 s).  Pinned by the reference's own recorded output - README.md:89 prints the
26 numbers examples/heap/heap.go:27-33 drew with rand.Seed(101), and this generator reproduces them - and by known
answers from Go's documentation and playground (tests/test_gorand.py): rngCooked[0] = -4181792142133755926, and after rand.Seed(1) rand.Int() =
5577006791947779410, 8674665223082153551, ..., rand.Float64() = 0.6046602879796196, 0.9405090880450124, ...,
rand.Intn(100) = 81, 87, 47, 59, 81, 18, 25, 40, 56, 0.

Host-side input generation only: nothing here is on the step path.
"""
from __future__ import annotations

import functools

import numpy as np

RNG_LEN, RNG_TAP = 607, 273
_FEED0 = RNG_LEN - RNG_TAP  # 334
_M64 = (1 << 64) - 1
_MASK63 = (1 << 63) - 1
_INT32_MAX = (1 << 31) - 1


def _seedrand(x: int) -> int:
    """x[n+1] = 48271 * x[n] mod (2^31 - 1), Schrage's form (Go rng.go seedrand)"""
    hi, lo = divmod(x, 44488)
    x = 48271 * lo - 3399 * hi
    return x + _INT32_MAX if x < 0 else x


def _lcg_fill(seed: int, shift_hi: int, shift_mid: int, xor_table=None):
    seed %= _INT32_MAX
    if seed == 0:
        seed = 89482311
    x = seed
    vec = [0] * RNG_LEN
    for i in range(-20, RNG_LEN):
        x = _seedrand(x)
        if i >= 0:
            u = (x << shift_hi) & _M64
            x = _seedrand(x)
            u ^= (x << shift_mid) & _M64
            x = _seedrand(x)
            u ^= x
            if xor_table is not None:
                u ^= xor_table[i]
            vec[i] = u
    return vec


def _mulmod(a: np.ndarray, b: np.ndarray) -> np.ndarray:
    """a*b mod (x^607 - x^334 - 1), coefficients mod 2^64"""
    with np.errstate(over="ignore"):
        prod = np.zeros(2 * RNG_LEN - 1, dtype=np.uint64)
        for i in np.flatnonzero(b):
            prod[i:i + RNG_LEN] += a * b[i]
        for k in range(2 * RNG_LEN - 2, RNG_LEN - 1, -1):  # x^k = x^(k-607+334) + x^(k-607)
            c = prod[k]
            prod[k - RNG_LEN + _FEED0] += c
            prod[k - RNG_LEN] += c
    return prod[:RNG_LEN].copy()


def _x_pow(n: int) -> np.ndarray:
    result = np.zeros(RNG_LEN, dtype=np.uint64)
    result[0] = 1
    base = np.zeros(RNG_LEN, dtype=np.uint64)
    base[1] = 1
    while n:
        if n & 1:
            result = _mulmod(result, base)
        base = _mulmod(base, base)
        n >>= 1
    return result


def _advance_state(vec, n: int):
    """the vector after n generator calls started with tap = 0, feed = 334 (positions as stored, like gen_cooked prints)

    With y_k the k-th output, y_k = y_{k-607} + y_{k-273}; call k overwrites vec[(334 - k) mod 607], so the initial
    vector is y_j = vec[(334 - j) mod 607] for j = -606..0."""
    y = np.zeros(2 * RNG_LEN, dtype=np.uint64)  # y[j + 606] = y_j for j = -606 .. 607
    for j in range(-606, 1):
        y[j + 606] = vec[(_FEED0 - j) % RNG_LEN]
    with np.errstate(over="ignore"):
        for k in range(1, RNG_LEN + 1):
            y[k + 606] = y[k - RNG_LEN + 606] + y[k - RNG_TAP + 606]
        r = _x_pow(n)
        out = [0] * RNG_LEN
        for j0 in range(-606, 1):  # y_{j0+n} = sum_j r_j y_{j0+j}
            v = int((r * y[j0 + 606:j0 + 606 + RNG_LEN]).sum(dtype=np.uint64))
            out[(_FEED0 - (j0 + n)) % RNG_LEN] = v
    return out


@functools.lru_cache(maxsize=None)
def rng_cooked():
    """Go's rngCooked table as unsigned 64-bit ints: "the state of the generator after 780e10 iterations" """
    return tuple(_advance_state(_lcg_fill(1, 20, 10), 7_800_000_000_000))


class Rand:
    """the top-level math/rand functions the reference calls, on one explicitly seeded Go 1 source"""

    def __init__(self, seed: int = 1):
        self.Seed(seed)

    def Seed(self, seed: int):
        vec = _lcg_fill(seed, 40, 20, rng_cooked())  # rngSource.Seed
        # outputs are produced in blocks: hist[-607:] always holds the last 607 outputs (= the state)
        self._hist = np.array([vec[(_FEED0 - j) % RNG_LEN] for j in range(-606, 1)], dtype=np.uint64)
        self._buf = np.zeros(0, dtype=np.uint64)

    def _raw(self, n: int) -> np.ndarray:
        """the next n Uint64 outputs"""
        while len(self._buf) < n:
            want = max(n - len(self._buf), 4096)
            y = np.concatenate([self._hist, np.zeros(want, dtype=np.uint64)])
            with np.errstate(over="ignore"):
                for a in range(RNG_LEN, RNG_LEN + want, RNG_TAP):  # a block of <= 273 depends on earlier blocks only
                    b = min(a + RNG_TAP, RNG_LEN + want)
                    y[a:b] = y[a - RNG_LEN:b - RNG_LEN] + y[a - RNG_TAP:b - RNG_TAP]
            self._buf = np.concatenate([self._buf, y[RNG_LEN:]])
            self._hist = y[-RNG_LEN:].copy()
        out, self._buf = self._buf[:n], self._buf[n:]
        return out

    # --- scalar API (names as in Go)
    def Uint64(self) -> int:
        return int(self._raw(1)[0])

    def Int63(self) -> int:
        return self.Uint64() & _MASK63

    def Int(self) -> int:  # 64-bit platforms: uint(Int63())
        return self.Int63()

    def Int31(self) -> int:
        return self.Int63() >> 32

    def Int31n(self, n: int) -> int:
        if n <= 0:
            raise ValueError("invalid argument to Int31n")
        if n & (n - 1) == 0:
            return self.Int31() & (n - 1)
        mx = (1 << 31) - 1 - (1 << 31) % n
        v = self.Int31()
        while v > mx:
            v = self.Int31()
        return v % n

    def Int63n(self, n: int) -> int:
        if n <= 0:
            raise ValueError("invalid argument to Int63n")
        if n & (n - 1) == 0:
            return self.Int63() & (n - 1)
        mx = (1 << 63) - 1 - (1 << 63) % n
        v = self.Int63()
        while v > mx:
            v = self.Int63()
        return v % n

    def Intn(self, n: int) -> int:
        return self.Int31n(n) if n <= _INT32_MAX else self.Int63n(n)

    def Float64(self) -> float:
        return float(self.Float64s(1)[0])

    # --- vector forms (same stream as calling the scalar function n times)
    def Ints(self, n: int) -> np.ndarray:
        return (self._raw(n) & np.uint64(_MASK63)).astype(np.int64)

    def Float64s(self, n: int) -> np.ndarray:
        """float64(Int63()) / (1<<63), drawn again when the division rounds up to 1 (Go's `again:` loop)"""
        out = np.zeros(0)
        while len(out) < n:
            f = (self._raw(n - len(out)) & np.uint64(_MASK63)).astype(np.int64).astype(np.float64) * 2.0 ** -63
            out = np.concatenate([out, f[f != 1.0]])
        return out


def uniform_rect_spawn(n: int, upper_left=(0.0, 0.0), lower_right=(1.0, 1.0), seed: int = 12345678, rand: "Rand" = None):
    """UniformRectSpawner.Spawn (config-parser.go:58-80): re-seed, n x (x, y) positions, then n Z values; E = 0.01.
    `rand`: the stream to re-seed and draw from (Go's spawners share the package-level source)"""
    if rand is None:
        r = Rand(seed)
    else:
        r = rand
        r.Seed(seed)
    u = r.Float64s(2 * n).reshape(n, 2)
    ul, lr = np.asarray(upper_left, float), np.asarray(lower_right, float)
    pos = ul + u * (lr - ul)  # UpperLeft.X + rand.Float64()*(LowerRight.X-UpperLeft.X)
    return dict(pos=pos, vel=np.zeros((n, 2)), e=np.full(n, 0.01), z=r.Ints(n))


def init_uniformly(n: int, seed: int = 12345678):
    """InitUniformly (core.go:76-91): two (x, y) draws per particle, the second kept; Rho = 1; then n Z values"""
    r = Rand(seed)
    u = r.Float64s(4 * n).reshape(n, 4)
    return dict(pos=u[:, 2:4].copy(), vel=np.zeros((n, 2)), e=np.zeros(n), rho=np.ones(n), z=r.Ints(n))


def point_source_spawn(r: Rand, n: int, origin):
    """PointSource.Spawn's particle loop (config-parser.go:89-99) on the caller's running stream (no re-seed): per
    particle dy, dx jitters then Z; Rho = 100, E = 0.002"""
    pos, z = np.zeros((n, 2)), np.zeros(n, dtype=np.int64)
    for i in range(n):
        dy = 0.01 * (-1 + 2 * r.Float64())
        dx = 0.01 * (-1 + 2 * r.Float64())
        pos[i] = (origin[0] + dx, origin[1] + dy)
        z[i] = r.Int()
    return dict(pos=pos, vel=np.zeros((n, 2)), e=np.full(n, 0.002), rho=np.full(n, 100.0), z=z)
