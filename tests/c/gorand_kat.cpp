// Known answers for include/sphb_gorand.hpp (Go's math/rand, Go 1 source): rngCooked as printed in Go's rng.go and the
// rand.Seed(1) streams Go's documentation and playground show.  Exit code 0 = all match.
#include <cstdio>
#include "sphb_gorand.hpp"

int main() {
  const gorand::Vec& c = gorand::rng_cooked();
  if ((int64_t)c[0] != -4181792142133755926ll || (int64_t)c[1] != -4576982950128230565ll || c[606] != 4152330101494654406ull) return 1;
  {  // jump-ahead == stepping
    gorand::Vec v = gorand::lcg_fill(1, 20, 10, nullptr), w = v;
    int tap = 0, feed = gorand::kFeed0;
    for (int k = 0; k < 3000; ++k) {
      if (--tap < 0) tap += gorand::kLen;
      if (--feed < 0) feed += gorand::kLen;
      w[feed] += w[tap];
    }
    if (gorand::advance_state(v, 3000) != w) return 2;
  }
  const int64_t ints[10] = {5577006791947779410ll, 8674665223082153551ll, 6129484611666145821ll, 4037200794235010051ll, 3916589616287113937ll,
                            6334824724549167320ll, 605394647632969758ll,  1443635317331776148ll, 894385949183117216ll,  2775422040480279449ll};
  gorand::Rand r(1);
  for (int64_t want : ints)
    if (r.Int() != want) return 3;
  const double floats[5] = {0.6046602879796196, 0.9405090880450124, 0.6645600532184904, 0.4377141871869802, 0.4246374970712657};
  r.Seed(1);
  for (double want : floats)
    if (r.Float64() != want) return 4;
  const int intn[10] = {81, 87, 47, 59, 81, 18, 25, 40, 56, 0};
  r.Seed(1);
  for (int want : intn)
    if (r.Intn(100) != want) return 5;
  {  // the reference's own recorded output: examples/heap/heap.go:27-33 with rand.Seed(101), printed in README.md:89
    const int readme[26] = {31, 37, 82, 83, 33, 54, 39, 42, 62, 49, 84, 59, 88, 26, 27, 21, 92, 97, 87, 49, 33, 9, 42, 49, 88, 67};
    r.Seed(101);
    for (int want : readme)
      if ((int)(r.Int() % 90 + 9) != want) return 6;
  }
  r.Seed(12345678);  // the reference's seed (config-parser.go:63): printed for the Python twin to compare
  const double f0 = r.Float64(), f1 = r.Float64();
  const long long z = (long long)r.Int();
  std::printf("%.17g %.17g %lld\n", f0, f1, z);
  return 0;
}
