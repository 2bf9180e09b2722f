// The reference's examples/sph-simulation (sph-simulation.go:49-58) over the C++ host side: sim.MakeSimulation(), then per
// step Step() and the frame the animator would draw.  The per-particle frame data (pixel coordinates and colour-ramp index,
// animator.go:75-101) comes from the device, 9 bytes per particle; rasterising it to a PNG stays with the caller (gx).
//   ./examples/sph_simulation [nsteps]        (the reference runs 10000)
#include <cstdio>
#include <cstdlib>
#include "sphb_sim.hpp"

int main(int argc, char** argv) {
  const int nsteps = argc > 1 ? std::atoi(argv[1]) : 100;
  try {
    sim::Simulation sph = sim::MakeSimulation();
    for (int i = 0; i < nsteps; ++i) {
      sph.Step();
      const sim::FrameData frame = sph.Frame(1280, 720);
      if (i % 10 == 0 || i == nsteps - 1) {
        unsigned hist[4] = {0, 0, 0, 0};
        for (uint8_t c : frame.colour) ++hist[c >> 6];
        std::printf("step %4d  sum E %.12g  colour index quartiles %u %u %u %u\n", sph.CurrentStep, sph.TotalEnergy(), hist[0], hist[1],
                    hist[2], hist[3]);
      }
    }
  } catch (const sim::Panic& p) {
    std::fprintf(stderr, "panic: %s\n", p.what());
    return 2;
  }
  return 0;
}
