// The reference's examples/speed-test (speed-test.go:22-45) over the C++ host side (include/sphb_sim.hpp): 100000
// particles of the default spawner (Go's math/rand stream), DeltaTHalf 0.02, g = (0, 0.2), 20 steps, FPS per step.
//   make -C examples && ./examples/speed_test [nparticles]
// Like a Go panic, a failure (no CUDA device: libsphb has no CPU fallback) prints "panic: ..." and exits with status 2.
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include "sphb_sim.hpp"

int main(int argc, char** argv) {
  try {
    sim::UniformRectSpawner spwn = sim::MakeUniformRectSpawner();
    spwn.NParticles = argc > 1 ? std::atoi(argv[1]) : 100000;
    sim::SphConfig conf = sim::MakeConfig();
    conf.Start.push_back(spwn);
    conf.DeltaTHalf = 0.02;
    conf.Acceleration = {0, 0.2};
    sim::Simulation sph = sim::MakeSimulationFromConf(conf);
    auto previous = std::chrono::steady_clock::now();
    double total = 0;
    for (int i = 0; i < 20; ++i) {
      sph.Step();
      sph.TotalEnergy();  // the step is asynchronous: a reduction waits for it (simviewer reads it every step too)
      const auto now = std::chrono::steady_clock::now();
      const double elapsed = std::chrono::duration<double>(now - previous).count();
      previous = now;
      total += elapsed;
      std::printf("Step %d FPS %g\n", i, 1 / elapsed);
    }
    std::printf("Took %.4g seconds, and got an average FPS of %.4g\n", total, 20 / total);
  } catch (const sim::Panic& p) {
    std::fprintf(stderr, "panic: %s\n", p.what());
    return 2;
  }
  return 0;
}
