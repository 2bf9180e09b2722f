// The numeric part of the reference's examples/density (density.go:41-141) over the C++ host side: the 1000 + 200 particle
// scene with periodic kNN on [0,1]^2 and the three kernels, then periodicVisualTest's 10000 + 1200 particles, open and
// periodic on [0.1,0.9]^2, TopHat.  Prints the density statistics the pictures are coloured by instead of drawing them
// (the gx canvas is outside the path).
#include <algorithm>
#include <cstdio>
#include "sphb_sim.hpp"

static void report(const char* what, sim::Simulation& sph) {
  const std::vector<sim::Particle> ps = sph.Particles();
  double lo = 1e300, hi = 0, sum = 0;
  for (const auto& p : ps) { lo = std::min(lo, p.Rho); hi = std::max(hi, p.Rho); sum += p.Rho; }
  std::printf("%-28s N %zu  rho min %.6g mean %.6g max %.6g\n", what, ps.size(), lo, sum / ps.size(), hi);
}

int main() {
  try {
    {  // main, density.go:41-97
      sim::SphConfig conf = sim::MakeConfig();
      conf.Start.push_back(sim::UniformRectSpawner{{0, 0}, {1, 1}, 1000});
      conf.Start.push_back(sim::UniformRectSpawner{{0.1, 0}, {0.3, 0.4}, 200});
      sim::Simulation sph = sim::MakeSimulationFromConf(conf);
      const double box[2] = {0, 1};
      sph.FindNearestNeighboursPeriodic(box, box);
      sph.Density2D(sim::Kernel::TopHat2D); report("TopHat2D", sph);
      sph.Density2D(sim::Kernel::Monahan2D); report("Monahan2D", sph);
      sph.Density2D(sim::Kernel::Wendtland2D); report("Wendtland2D", sph);
    }
    {  // periodicVisualTest, density.go:99-141
      sim::SphConfig conf = sim::MakeConfig();
      conf.Start.push_back(sim::UniformRectSpawner{{0.1, 0.1}, {0.9, 0.9}, 10000});
      conf.Start.push_back(sim::UniformRectSpawner{{0.85, 0.4}, {0.9, 0.9}, 1200});
      sim::Simulation sph = sim::MakeSimulationFromConf(conf);
      sph.FindNearestNeighbours();
      sph.Density2D(sim::Kernel::TopHat2D); report("open, TopHat2D", sph);
      const double box[2] = {0.1, 0.9};
      sph.FindNearestNeighboursPeriodic(box, box);
      sph.Density2D(sim::Kernel::TopHat2D); report("periodic [0.1,0.9]^2, TopHat2D", sph);
    }
  } catch (const sim::Panic& p) {
    std::fprintf(stderr, "panic: %s\n", p.what());
    return 2;
  }
  return 0;
}
