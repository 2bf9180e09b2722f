// sphb_sim.hpp — C++ host-side mirror of the step API of the reference's Go package `sim`, over the C ABI of
// libsphb.so (include/sphb.h).  Header only; no CUDA or torch types.
//
// The reference is compiled Go and no Go toolchain exists where this repository is built, so next to the cgo shim
// (go/sim/backend_cuda.go) this is the compiled-language host side: same names, argument meaning
// and error behaviour as the Go API that simviewer and the examples use:
//
//   Go (reference)                                                     here (namespace sim)
//   MakeConfig()                            config-parser.go:131-149   MakeConfig()
//   SphConfig{NSteps, DeltaTHalf, ...}      config-parser.go:111-128   SphConfig
//   MakeUniformRectSpawner / Spawn          config-parser.go:50-80     same (Go's math/rand stream: sphb_gorand.hpp)
//   PointSource.Spawn                       config-parser.go:82-102    same
//   MakeSimulation()                        sph.go:23-30               MakeSimulation()
//   MakeSimulationFromConf(conf)            sph.go:40-54               MakeSimulationFromConf(conf) (Start spawners, Sources);
//                                                                      MakeSimulationFromParticles(conf, arrays) for own particles
//   (*Simulation).Step / Run / CalculateForces        sph.go:56-64,403 Simulation::Step / Run / CalculateForces
//   TotalEnergy / TotalDensity / TotalMomentum        sph.go:441-463   same
//   p.FindNearestNeighbours[Periodic](root, hor, ver) nearest-neighbour.go:15-67   FindNearestNeighbours[Periodic] (batch)
//   Density2D(p, sim, kernel)               sph.go:306-323             Density2D(kernel) (batch)
//   sim.Root.Particles[i].{Pos, Vel, Rho, ...}        core.go:17-42    Particles() (download, in spawn order)
//   (*Animator).CurrentFrame's per-particle arithmetic animator.go:75-101      FrameData(width, height)
//
// Reference panics (sph.go:93,251,317,354; nearest-neighbour.go:44,53) become sim::Panic carrying the library's
// message; there is no CPU fallback: without a CUDA device construction throws Panic with code SPHB_E_CUDA.
// Initial particles always cross the boundary as explicit arrays.  The device id of a particle is its spawn index
// (dense, which the by-id transfers need); Particle.Z, the random draw the renderer orders by, stays on the host (Z()).
#pragma once
#include <algorithm>
#include <cstdint>
#include <numeric>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "sphb.h"
#include "sphb_gorand.hpp"

namespace sim {

constexpr double MaxFloat64 = SPHB_OPEN_HI;

struct Vec2 { double X = 0, Y = 0; };  // linear-algebra.go:36-38

struct Panic : std::runtime_error {  // what is a Go panic in the reference
  int code;
  Panic(int c, const std::string& m) : std::runtime_error(m), code(c) {}
};

enum class Kernel : int32_t { TopHat2D = SPHB_KERNEL_TOPHAT, Monahan2D = SPHB_KERNEL_MONAGHAN, Wendtland2D = SPHB_KERNEL_WENDLAND };

struct Reflections { double L = -MaxFloat64, R = MaxFloat64, U = -MaxFloat64, D = MaxFloat64; };  // config-parser.go:104-109

struct Particle {  // the fields of core.go:17-42 the step path consumes and produces
  Vec2 Pos, Vel, VDot;
  double Rho = 0, C = 0, E = 0, EDot = 0, H = 0;  // H = NNDists[0]
  int64_t Z = 0;
};

// the package-level math/rand source all spawners share.  A Go >= 1.20 process that never calls rand.Seed starts it
// from a random seed; every scene of the reference seeds it (each UniformRectSpawner.Spawn does), so this only
// matters for a config with sources and no Start rectangle: seed 1 (Go's documented pre-1.20 default) is used then.
inline gorand::Rand& GlobalRand() {
  static gorand::Rand r(1);
  return r;
}

struct UniformRectSpawner {  // config-parser.go:37-41
  Vec2 UpperLeft{0, 0}, LowerRight{1, 1};
  int NParticles = 1000;
  std::vector<Particle> Spawn(double /*t*/) const {  // config-parser.go:58-80: re-seed, positions, then Z; E = 0.01
    gorand::Rand& r = GlobalRand();
    r.Seed(12345678);
    std::vector<Particle> ps(NParticles);
    for (auto& p : ps) {
      const double x = UpperLeft.X + r.Float64() * (LowerRight.X - UpperLeft.X);
      const double y = UpperLeft.Y + r.Float64() * (LowerRight.Y - UpperLeft.Y);
      p.Pos = {x, y};
    }
    for (auto& p : ps) { p.Z = r.Int(); p.E = 0.01; }
    return ps;
  }
};
inline UniformRectSpawner MakeUniformRectSpawner() { return UniformRectSpawner{}; }  // config-parser.go:50-55

struct PointSource {  // config-parser.go:43-47
  Vec2 origin;
  double rate = 1;
  double LastSpwned = 0;
  std::vector<Particle> Spawn(double t) {  // config-parser.go:82-102: the running stream, no re-seed
    const double cooldown = 1 / rate;
    const int n = (int)((t - LastSpwned) / cooldown);
    LastSpwned += double(n) * cooldown;
    gorand::Rand& r = GlobalRand();
    std::vector<Particle> ps(n > 0 ? n : 0);
    for (auto& p : ps) {
      const double dy = 0.01 * (-1 + 2 * r.Float64());
      const double dx = 0.01 * (-1 + 2 * r.Float64());
      p.Pos = {origin.X + dx, origin.Y + dy};
      p.Rho = 100;
      p.Z = r.Int();
      p.E = 0.002;
    }
    return ps;
  }
};

struct SphConfig {  // config-parser.go:111-128 (Viewport stays with the caller)
  int NSteps = 10000;
  double DeltaTHalf = 0.001, Gamma = 1.66666, ParticleMass = 1;
  Vec2 Acceleration;
  Kernel kernel = Kernel::Monahan2D;
  double HorPeriodicity[2] = {-MaxFloat64, MaxFloat64};   // open
  double VertPeriodicity[2] = {-MaxFloat64, MaxFloat64};
  Reflections reflections;
  std::vector<PointSource> Sources;
  std::vector<UniformRectSpawner> Start;
};
inline SphConfig MakeConfig() { return SphConfig{}; }  // config-parser.go:131-149

struct FrameData {  // what (*Animator).CurrentFrame derives per particle (animator.go:75-101), in spawn order
  std::vector<float> xy;        // float32(Pos) * float32(size), interleaved
  std::vector<uint8_t> colour;  // ramp index uint8(min(Rho / (m N 10) * 256, 255))
};

class Simulation {
 public:
  SphConfig Config;  // public and mutable like sim.Config (sph.go:15): pushed to the device before every call
  int CurrentStep = 0;

  // MakeSimulationFromConf with the spawned particles passed in: pos/vel interleaved x,y; vel, e, rho, z may be empty.
  // The device ids are the spawn indices 0..n-1; z (Particle.Z) is kept on the host.
  Simulation(const SphConfig& conf, const std::vector<double>& pos_xy, const std::vector<double>& vel_xy = {},
             const std::vector<double>& e = {}, const std::vector<int64_t>& z = {}, int device = 0, int precision = 64,
             const std::vector<double>& rho = {})
      : Config(conf), z_(z), device_(device), precision_(precision) {
    const int64_t n = (int64_t)pos_xy.size() / 2;
    z_.resize(n, 0);
    const sphb_params p = params();
    const int64_t capacity = n + (conf.Sources.empty() ? 0 : 100000);  // sph.go:45 reserves 100000 for the sources
    const int rc = sphb_create(&p, n, capacity, pos_xy.data(), vel_xy.empty() ? nullptr : vel_xy.data(), e.empty() ? nullptr : e.data(),
                               rho.empty() ? nullptr : rho.data(), nullptr, &h_);
    if (rc != SPHB_OK) throw Panic(rc, sphb_last_error(nullptr));
  }
  Simulation(const Simulation&) = delete;
  Simulation& operator=(const Simulation&) = delete;
  Simulation(Simulation&& o) noexcept
      : Config(std::move(o.Config)), CurrentStep(o.CurrentStep), h_(o.h_), z_(std::move(o.z_)), device_(o.device_), precision_(o.precision_) {
    o.h_ = nullptr;
  }
  ~Simulation() { sphb_destroy(h_); }

  void Step() {  // sph.go:64-198 (the step-0 double force evaluation is inside the library)
    push();
    const double t = double(CurrentStep) * Config.DeltaTHalf * 2;  // sources spawn first, sph.go:72-86
    for (auto& src : Config.Sources) Append(src.Spawn(t));
    if (Len() == 0) throw Panic(SPHB_E_STATE, "int Run(): Simulation not initialized!");  // sph.go:92-94
    check(sphb_step(h_, 1));
    ++CurrentStep;
  }
  // append(sim.Root.Particles, newParticles...) + MakeCells (sph.go:75-86): ids continue the spawn index
  void Append(const std::vector<Particle>& ps) {
    if (ps.empty()) return;
    const int64_t n0 = Len(), k = (int64_t)ps.size();
    std::vector<double> pos(2 * k), vel(2 * k), e(k), rho(k);
    std::vector<int64_t> id(k);
    for (int64_t i = 0; i < k; ++i) {
      pos[2 * i] = ps[i].Pos.X; pos[2 * i + 1] = ps[i].Pos.Y; vel[2 * i] = ps[i].Vel.X; vel[2 * i + 1] = ps[i].Vel.Y;
      e[i] = ps[i].E; rho[i] = ps[i].Rho; id[i] = n0 + i;
      z_.push_back(ps[i].Z);
    }
    check(sphb_append(h_, k, pos.data(), vel.data(), e.data(), rho.data(), id.data()));
  }
  void Run() { for (int s = 0; s < Config.NSteps; ++s) Step(); }  // sph.go:56-61
  void CalculateForces() { push(); check(sphb_calc_forces(h_)); }  // sph.go:403-435
  double TotalEnergy() { return reduce(SPHB_SUM_E); }              // sph.go:441-447
  double TotalDensity() { return reduce(SPHB_SUM_RHO); }           // sph.go:449-455
  double TotalMomentum() { return reduce(SPHB_LAST_VEL_NORM); }    // sph.go:457-463 (keeps the `=` of the reference)

  void FindNearestNeighboursPeriodic(const double hor[2], const double ver[2]) { push(); check(sphb_knn(h_, hor, ver)); }
  void FindNearestNeighbours() {
    const double open[2] = {-MaxFloat64, MaxFloat64};
    FindNearestNeighboursPeriodic(open, open);
  }
  void Density2D(Kernel k) { push(); check(sphb_density(h_, (int32_t)k)); }

  int64_t Len() const { return sphb_count(h_); }

  const std::vector<int64_t>& Z() const { return z_; }  // Particle.Z by spawn index

  FrameData Frame(int width, int height) {  // per-particle frame data in spawn order, 9 bytes per particle over the bus
    const int64_t n = Len();
    FrameData f;
    f.xy.resize(2 * n);
    f.colour.resize(n);
    int64_t n_out = 0;
    if (n) check(sphb_frame(h_, width, height, f.xy.data(), f.colour.data(), nullptr, n, &n_out));
    return f;
  }

  // sim.Root.Particles in spawn order (the device order is cell order and changes with every step)
  std::vector<Particle> Particles() {
    const int64_t n = Len();
    std::vector<double> pos(2 * n), vel(2 * n), vdot(2 * n), rho(n), c(n), e(n), edot(n), h(n);
    std::vector<int64_t> id(n);
    void* ptrs[SPHB_F_COUNT] = {};
    ptrs[SPHB_F_POS] = pos.data(); ptrs[SPHB_F_VEL] = vel.data(); ptrs[SPHB_F_VDOT] = vdot.data();
    ptrs[SPHB_F_RHO] = rho.data(); ptrs[SPHB_F_C] = c.data(); ptrs[SPHB_F_E] = e.data(); ptrs[SPHB_F_EDOT] = edot.data();
    ptrs[SPHB_F_H] = h.data(); ptrs[SPHB_F_ID] = id.data();
    const uint32_t mask = SPHB_MASK(SPHB_F_POS) | SPHB_MASK(SPHB_F_VEL) | SPHB_MASK(SPHB_F_VDOT) | SPHB_MASK(SPHB_F_RHO) |
                          SPHB_MASK(SPHB_F_C) | SPHB_MASK(SPHB_F_E) | SPHB_MASK(SPHB_F_EDOT) | SPHB_MASK(SPHB_F_H) | SPHB_MASK(SPHB_F_ID);
    int64_t n_out = 0;
    check(sphb_download(h_, mask, ptrs, n, &n_out));
    std::vector<int64_t> order(n);
    std::iota(order.begin(), order.end(), 0);
    std::sort(order.begin(), order.end(), [&](int64_t a, int64_t b) { return id[a] < id[b]; });
    std::vector<Particle> out(n);
    for (int64_t k = 0; k < n; ++k) {
      const int64_t j = order[k];
      Particle& p = out[k];
      p.Pos = {pos[2 * j], pos[2 * j + 1]}; p.Vel = {vel[2 * j], vel[2 * j + 1]}; p.VDot = {vdot[2 * j], vdot[2 * j + 1]};
      p.Rho = rho[j]; p.C = c[j]; p.E = e[j]; p.EDot = edot[j]; p.H = h[j];
      p.Z = (size_t)id[j] < z_.size() ? z_[id[j]] : id[j];
    }
    return out;
  }

 private:
  sphb_sim* h_ = nullptr;
  std::vector<int64_t> z_;
  int device_ = 0, precision_ = 64;

  sphb_params params() const {
    sphb_params p{};
    p.dt_half = Config.DeltaTHalf; p.gamma = Config.Gamma; p.particle_mass = Config.ParticleMass;
    p.accel[0] = Config.Acceleration.X; p.accel[1] = Config.Acceleration.Y;
    p.hor[0] = Config.HorPeriodicity[0]; p.hor[1] = Config.HorPeriodicity[1];
    p.ver[0] = Config.VertPeriodicity[0]; p.ver[1] = Config.VertPeriodicity[1];
    p.refl_L = Config.reflections.L; p.refl_R = Config.reflections.R; p.refl_U = Config.reflections.U; p.refl_D = Config.reflections.D;
    p.kernel = (int32_t)Config.kernel; p.precision = precision_; p.device = device_; p.flags = 0;
    return p;
  }
  void push() {  // callers edit sim.Config between steps
    const sphb_params p = params();
    check(sphb_set_params(h_, &p));
  }
  void check(int rc) const {
    if (rc != SPHB_OK) throw Panic(rc, sphb_last_error(h_));
  }
  double reduce(int32_t which) {
    double out = 0;
    check(sphb_reduce(h_, which, &out));
    return out;
  }
};

inline Simulation MakeSimulationFromParticles(const SphConfig& conf, const std::vector<double>& pos_xy,
                                              const std::vector<double>& vel_xy = {}, const std::vector<double>& e = {},
                                              const std::vector<int64_t>& z = {}, int device = 0) {
  return Simulation(conf, pos_xy, vel_xy, e, z, device);
}

inline Simulation MakeSimulationFromParticleList(const SphConfig& conf, const std::vector<Particle>& ps, int device = 0) {
  const size_t n = ps.size();
  std::vector<double> pos(2 * n), vel(2 * n), e(n), rho(n);
  std::vector<int64_t> z(n);
  for (size_t i = 0; i < n; ++i) {
    pos[2 * i] = ps[i].Pos.X; pos[2 * i + 1] = ps[i].Pos.Y; vel[2 * i] = ps[i].Vel.X; vel[2 * i + 1] = ps[i].Vel.Y;
    e[i] = ps[i].E; rho[i] = ps[i].Rho; z[i] = ps[i].Z;
  }
  return Simulation(conf, pos, vel, e, z, device, 64, rho);
}

// sph.go:40-54: every Start spawner spawns at t = 0 (each re-seeds, so rectangles share their uniforms)
inline Simulation MakeSimulationFromConf(const SphConfig& conf, int device = 0) {
  std::vector<Particle> ps;
  for (const auto& sp : conf.Start) {
    const std::vector<Particle> part = sp.Spawn(0);
    ps.insert(ps.end(), part.begin(), part.end());
  }
  return MakeSimulationFromParticleList(conf, ps, device);
}

// sph.go:23-30: MakeConfig defaults and one default spawner (1000 particles in the unit square)
inline Simulation MakeSimulation(int device = 0) {
  SphConfig conf = MakeConfig();
  return MakeSimulationFromParticleList(conf, MakeUniformRectSpawner().Spawn(0), device);
}

}  // namespace sim
