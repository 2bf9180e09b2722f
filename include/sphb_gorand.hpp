// sphb_gorand.hpp — Go's math/rand (the Go 1 additive lagged-Fibonacci source) bit for bit, header only.
//
// Every particle the reference spawns is drawn from the package-level functions of Go's math/rand after an explicit
// rand.Seed(12345678) (sim/config-parser.go:58-80 UniformRectSpawner.Spawn, :82-102 PointSource.Spawn,
// sim/core.go:76-91 InitUniformly).  The C++ host side (sphb_sim.hpp) needs the same stream to build the same scenes.
//
// x[n] = x[n-607] + x[n-273] mod 2^64; seeding XORs the 607-entry table rngCooked (Go's rng.go) into a vector filled by
// the LCG x' = 48271 x mod (2^31-1).  Go's sources are not available where this repository is built, so the table is
// recomputed from its definition (Go's gen_cooked.go: LCG fill with seed 1, then 7.8e12 generator steps) by jump-ahead:
// the recurrence is linear over Z/2^64, so 7.8e12 steps are x^n mod (x^607 - x^334 - 1), 43 polynomial squarings.
// Pinned by the reference's own recorded output (README.md:89 prints what examples/heap/heap.go:27-33 drew after
// rand.Seed(101); reproduced here) and by known answers (tests/c/gorand_kat.cpp, tests/test_gorand.py):
// rngCooked[0] = -4181792142133755926; after Seed(1): Int() = 5577006791947779410, 8674665223082153551, ...;
// Float64() = 0.6046602879796196, ...; Intn(100) = 81, 87, 47, ...
// The Python twin is sphugo_b200/gorand.py.  Input generation only: nothing here is on the step path.
#pragma once
#include <array>
#include <cstdint>
#include <stdexcept>
#include <vector>

namespace gorand {

constexpr int kLen = 607, kTap = 273, kFeed0 = kLen - kTap;
using Vec = std::array<uint64_t, kLen>;

inline int32_t seedrand(int32_t x) {  // Schrage's form of 48271 * x mod (2^31 - 1)
  const int32_t hi = x / 44488, lo = x % 44488;
  x = 48271 * lo - 3399 * hi;
  return x < 0 ? x + 2147483647 : x;
}

inline Vec lcg_fill(int64_t seed, int shift_hi, int shift_mid, const Vec* xor_table) {
  seed %= 2147483647;
  if (seed < 0) seed += 2147483647;
  if (seed == 0) seed = 89482311;
  int32_t x = (int32_t)seed;
  Vec v{};
  for (int i = -20; i < kLen; ++i) {
    x = seedrand(x);
    if (i >= 0) {
      uint64_t u = (uint64_t)x << shift_hi;
      x = seedrand(x);
      u ^= (uint64_t)x << shift_mid;
      x = seedrand(x);
      u ^= (uint64_t)x;
      if (xor_table) u ^= (*xor_table)[i];
      v[i] = u;
    }
  }
  return v;
}

inline Vec mulmod(const Vec& a, const Vec& b) {  // a * b mod (x^607 - x^334 - 1), coefficients mod 2^64
  std::vector<uint64_t> p(2 * kLen - 1, 0);
  for (int i = 0; i < kLen; ++i)
    if (b[i])
      for (int j = 0; j < kLen; ++j) p[i + j] += a[j] * b[i];
  for (int k = 2 * kLen - 2; k >= kLen; --k) {
    p[k - kLen + kFeed0] += p[k];
    p[k - kLen] += p[k];
  }
  Vec r;
  for (int i = 0; i < kLen; ++i) r[i] = p[i];
  return r;
}

// the vector after n generator calls started with tap = 0, feed = 334, in storage order (what gen_cooked prints).
// With y_k the k-th output, y_k = y_{k-607} + y_{k-273}; call k overwrites vec[(334 - k) mod 607].
inline Vec advance_state(const Vec& vec, uint64_t n) {
  std::vector<uint64_t> y(2 * kLen);  // y[j + 606] = y_j, j = -606 .. 607
  for (int j = -606; j <= 0; ++j) y[j + 606] = vec[((kFeed0 - j) % kLen + kLen) % kLen];
  for (int k = 1; k <= kLen; ++k) y[k + 606] = y[k - kLen + 606] + y[k - kTap + 606];
  Vec result{}, base{};
  result[0] = 1;
  base[1] = 1;
  for (uint64_t e = n; e; e >>= 1) {
    if (e & 1) result = mulmod(result, base);
    base = mulmod(base, base);
  }
  Vec out{};
  const int nm = (int)(n % kLen);
  for (int j0 = -606; j0 <= 0; ++j0) {  // y_{j0+n} = sum_j r_j y_{j0+j}
    uint64_t s = 0;
    for (int j = 0; j < kLen; ++j) s += result[j] * y[j0 + 606 + j];
    out[(((kFeed0 - j0 - nm) % kLen) + kLen) % kLen] = s;
  }
  return out;
}

inline const Vec& rng_cooked() {  // "the state of the generator after 780e10 iterations"
  static const Vec table = advance_state(lcg_fill(1, 20, 10, nullptr), 7800000000000ull);
  return table;
}

class Rand {  // the package-level math/rand functions the reference calls, on one explicitly seeded Go 1 source
 public:
  explicit Rand(int64_t seed = 1) { Seed(seed); }
  void Seed(int64_t seed) {
    vec_ = lcg_fill(seed, 40, 20, &rng_cooked());
    tap_ = 0;
    feed_ = kFeed0;
  }
  uint64_t Uint64() {
    if (--tap_ < 0) tap_ += kLen;
    if (--feed_ < 0) feed_ += kLen;
    return vec_[feed_] += vec_[tap_];
  }
  int64_t Int63() { return (int64_t)(Uint64() & 0x7fffffffffffffffull); }
  int64_t Int() { return Int63(); }  // 64-bit platforms
  int32_t Int31() { return (int32_t)(Int63() >> 32); }
  int32_t Int31n(int32_t n) {
    if (n <= 0) throw std::invalid_argument("invalid argument to Int31n");
    if ((n & (n - 1)) == 0) return Int31() & (n - 1);
    const int32_t mx = (int32_t)((1u << 31) - 1 - (1u << 31) % (uint32_t)n);
    int32_t v = Int31();
    while (v > mx) v = Int31();
    return v % n;
  }
  int64_t Int63n(int64_t n) {
    if (n <= 0) throw std::invalid_argument("invalid argument to Int63n");
    if ((n & (n - 1)) == 0) return Int63() & (n - 1);
    const int64_t mx = (int64_t)((1ull << 63) - 1 - (1ull << 63) % (uint64_t)n);
    int64_t v = Int63();
    while (v > mx) v = Int63();
    return v % n;
  }
  int64_t Intn(int64_t n) { return n <= 2147483647 ? (int64_t)Int31n((int32_t)n) : Int63n(n); }
  double Float64() {  // float64(Int63()) / (1 << 63), drawn again when the division rounds up to 1
    for (;;) {
      const double f = (double)Int63() / 9223372036854775808.0;
      if (f != 1.0) return f;
    }
  }

 private:
  Vec vec_{};
  int tap_ = 0, feed_ = kFeed0;
};

}  // namespace gorand
