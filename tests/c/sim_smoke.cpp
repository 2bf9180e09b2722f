// C++ client of include/sphb_sim.hpp (the compiled-language mirror of the Go `sim` step API).  Without a GPU the
// constructor must throw sim::Panic with SPHB_E_CUDA; with one, the speed-test shape (examples/speed-test/
// speed-test.go:22-43 at reduced N) steps and conserves what it should.  Exit code 0 = behaved as specified.
#include <cmath>
#include <cstdio>
#include "sphb_sim.hpp"

int main() {
  const int nx = 48;
  std::vector<double> pos, e(nx * nx, 0.01);
  uint64_t s = 12345678;  // splitmix64 jitter, like sphugo_b200/gen.py
  auto next = [&]() { s += 0x9E3779B97F4A7C15ull; uint64_t z = s; z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull; z = (z ^ (z >> 27)) * 0x94D049BB133111EBull; return ((z ^ (z >> 31)) >> 11) * (1.0 / 9007199254740992.0); };
  for (int j = 0; j < nx; ++j)
    for (int i = 0; i < nx; ++i) { pos.push_back((i + 0.5 + 0.5 * (next() - 0.5)) / nx); pos.push_back((j + 0.5 + 0.5 * (next() - 0.5)) / nx); }
  sim::SphConfig conf = sim::MakeConfig();
  conf.HorPeriodicity[0] = 0; conf.HorPeriodicity[1] = 1; conf.VertPeriodicity[0] = 0; conf.VertPeriodicity[1] = 1;
  conf.Acceleration = {0, 0.2};
  conf.DeltaTHalf = 0.002;
  // host-side pieces that need no device: the spawners draw Go's math/rand stream (config-parser.go:58-102)
  {
    const std::vector<sim::Particle> a = sim::MakeUniformRectSpawner().Spawn(0);
    gorand::Rand r(12345678);
    const double x0 = r.Float64(), y0 = r.Float64();
    if (a.size() != 1000 || a[0].Pos.X != x0 || a[0].Pos.Y != y0 || a[0].E != 0.01) return 10;
    sim::PointSource src{{0.2, 0.2}, 100.0};
    if (!src.Spawn(0.0).empty()) return 11;
    const std::vector<sim::Particle> b = src.Spawn(0.035);  // int(0.035 * 100) = 3, continuing the stream after the Z draws
    if (b.size() != 3 || std::fabs(src.LastSpwned - 0.03) > 1e-15 || b[0].Rho != 100 || b[0].E != 0.002) return 12;
    if (std::fabs(b[0].Pos.X - 0.2) > 0.01 || std::fabs(b[0].Pos.Y - 0.2) > 0.01) return 13;
  }
  try {
    {  // sim.MakeSimulation() (sph.go:23-30) and a config with a source: particles join before each step
      sim::Simulation d = sim::MakeSimulation();
      d.Step();
      if (d.Len() != 1000 || d.Z().size() != 1000 || d.Particles()[0].Z != d.Z()[0]) return 20;
      sim::SphConfig c2 = sim::MakeConfig();
      c2.Start.push_back(sim::UniformRectSpawner{{0.2, 0.25}, {0.5, 0.5}, 1500});
      c2.Sources.push_back(sim::PointSource{{0.35, 0.3}, 1000.0});
      c2.DeltaTHalf = 0.002;
      c2.kernel = sim::Kernel::Wendtland2D;
      sim::Simulation q = sim::MakeSimulationFromConf(c2);
      for (int k = 0; k < 3; ++k) q.Step();
      const sim::FrameData f = q.Frame(1280, 720);
      if (q.Len() != 1508 || q.Z().size() != 1508 || f.colour.size() != 1508) return 21;
    }
    sim::Simulation simu = sim::MakeSimulationFromParticles(conf, pos, {}, e);
    const double e0 = 0.01 * nx * nx;
    for (int k = 0; k < 5; ++k) simu.Step();
    const double e5 = simu.TotalEnergy();
    auto ps = simu.Particles();
    double rho = 0;
    for (auto& p : ps) rho += p.Rho;
    std::printf("gpu path: %lld particles, step %d, sum E %.12g (start %.12g), mean rho %.6g\n", (long long)simu.Len(), simu.CurrentStep,
                e5, e0, rho / ps.size());
    // a near-uniform box: the thermal energy barely moves in 5 steps; the mean density sits about 17 % under N m because
    // Density2D has no self term (sph.go:306-323; the oracle gives 1912.7 for this input, N m = 2304)
    const double mean = rho / ps.size(), nm = double(nx) * nx;
    if (!(std::fabs(e5 - e0) < 1e-2 * e0) || !(mean > 0.78 * nm && mean < 0.88 * nm) || ps.size() != (size_t)nx * nx) return 2;
    bool thrown = false;
    try { simu.Density2D(sim::Kernel::TopHat2D); simu.Config.kernel = sim::Kernel::TopHat2D; simu.Step(); } catch (const sim::Panic& p) { thrown = p.code == SPHB_E_KERNEL; }
    return thrown ? 0 : 3;  // TopHat2D.DF panics in the reference (sph.go:251-253)
  } catch (const sim::Panic& p) {
    std::printf("no device: Panic(%d): %s\n", p.code, p.what());
    return p.code == SPHB_E_CUDA ? 0 : 1;
  }
}
