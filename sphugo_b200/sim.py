"""Host-side mirror of the reference's Go package `sim` for the step path, over the C ABI of libsphb.so.

Same names, argument meaning and error behaviour as the Go API that simviewer and the examples use
(SURVEY §8b), so that the parity tests read like the reference's own call sites:

    Go (reference)                                   here
    sim.MakeConfig()                 config-parser.go:131   MakeConfig()
    sim.MakeUniformRectSpawner()     config-parser.go:50    MakeUniformRectSpawner()
    sim.MakeSimulation()             sph.go:23              MakeSimulation()
    sim.MakeSimulationFromConf(c)    sph.go:40              MakeSimulationFromConf(c)
    (*Simulation).Step()             sph.go:64              Simulation.Step()
    (*Simulation).Run()              sph.go:56              Simulation.Run()
    (*Simulation).CalculateForces()  sph.go:403             Simulation.CalculateForces()
    TotalEnergy/TotalDensity/TotalMomentum  sph.go:441-463  same
    p.FindNearestNeighboursPeriodic(root, hor, ver)  nearest-neighbour.go:28   Simulation.FindNearestNeighboursPeriodic(hor, ver)  (batch)
    Density2D(p, sim, kernel)        sph.go:306             Simulation.Density2D(kernel)                          (batch)
    sim.Root.Particles[i].{Pos,...}  core.go:17             Simulation.Particles()  (lazy, field-masked download)
    (*Animator).CurrentFrame's per-particle arithmetic  animator.go:75-101   Simulation.FrameData()

Reference panics become SimPanic (the Go shim re-panics); config errors stay ValueError.
All numerics run in the CUDA library; this file only marshals SphConfig into sphb_params.
"""
from __future__ import annotations

import dataclasses
import math
from typing import List, Optional, Sequence, Tuple

import numpy as np

from . import _lib as L
from . import gen
from . import gorand

MaxFloat64 = 1.7976931348623157e308


class SimPanic(RuntimeError):
    """what would be a Go panic in the reference (sph.go:93,251,317,354; nearest-neighbour.go:44,53)"""


@dataclasses.dataclass(frozen=True)
class Kernel:
    """sim.Kernel (sph.go:237-242) identified by value: closures cannot cross a C ABI."""

    name: str
    id: int
    FPrefactor: float
    DFPrefactor: float


TopHat2D = Kernel("TopHat2D", L.KERNEL_TOPHAT, 1 / math.pi, 1.0)
Monahan2D = Kernel("Monahan2D", L.KERNEL_MONAGHAN, 6 * 40 / (math.pi * 7), 6 * 40 / (math.pi * 7))
Wendtland2D = Kernel("Wendtland2D", L.KERNEL_WENDLAND, 4 * 7 / (math.pi * 4), 8 * 7 / (math.pi * 4))
KERNELS_BY_CONFIG_NAME = {"Monahan": Monahan2D, "Wendtland": Wendtland2D}  # config-parser.go:273-281


@dataclasses.dataclass
class Reflections:  # config-parser.go:104-109
    L: float = -MaxFloat64
    R: float = MaxFloat64
    U: float = -MaxFloat64
    D: float = MaxFloat64


@dataclasses.dataclass
class UniformRectSpawner:  # config-parser.go:37-41
    UpperLeft: Tuple[float, float] = (0.0, 0.0)
    LowerRight: Tuple[float, float] = (1.0, 1.0)
    NParticles: int = 1000

    def Spawn(self, t: float = 0.0, seed: int = gen.DEFAULT_SEED):
        """positions (x then y per particle), then Z per particle; E = 0.01, everything else zero
        (config-parser.go:58-80).  The stream is Go's math/rand after rand.Seed(12345678), reconstructed bit for
        bit in gorand.py, so the particles are the ones the Go binary spawns."""
        return gorand.uniform_rect_spawn(self.NParticles, self.UpperLeft, self.LowerRight, seed, rand=GlobalRand)


def MakeUniformRectSpawner() -> UniformRectSpawner:
    return UniformRectSpawner()


# the package-level math/rand source every spawner draws from.  A Go >= 1.20 process that never calls rand.Seed starts
# it from a random seed; every scene of the reference seeds it (each UniformRectSpawner.Spawn does), so it only matters
# for a config with sources and no Start rectangle: seed 1 (Go's documented pre-1.20 default) is used then.
GlobalRand = gorand.Rand(1)


@dataclasses.dataclass
class PointSource:  # config-parser.go:43-47
    origin: Tuple[float, float] = (0.0, 0.0)
    rate: float = 1.0
    LastSpwned: float = 0.0

    def Spawn(self, t: float):
        """n = int((t - LastSpwned) * rate) particles jittered by +-0.01 around origin, Rho = 100, E = 0.002, drawn
        from the running package-level stream without re-seeding (config-parser.go:82-102)"""
        cooldown = 1 / self.rate
        n = int((t - self.LastSpwned) / cooldown)
        self.LastSpwned += float(n) * cooldown
        return gorand.point_source_spawn(GlobalRand, n, self.origin)


@dataclasses.dataclass
class SphConfig:  # config-parser.go:111-128
    NSteps: int = 10000
    DeltaTHalf: float = 0.001
    Gamma: float = 1.66666
    ParticleMass: float = 1.0
    Acceleration: Tuple[float, float] = (0.0, 0.0)
    Kernel: Kernel = Monahan2D
    HorPeriodicity: Tuple[float, float] = (-MaxFloat64, MaxFloat64)
    VertPeriodicity: Tuple[float, float] = (-MaxFloat64, MaxFloat64)
    Reflections: Reflections = dataclasses.field(default_factory=Reflections)
    Sources: list = dataclasses.field(default_factory=list)
    Start: List[UniformRectSpawner] = dataclasses.field(default_factory=list)
    Viewport: Tuple[Tuple[float, float], Tuple[float, float]] = ((0.0, 0.0), (1.0, 1.0))

    def to_params(self, device: int = 0, precision: int = 64) -> L.Params:
        r = self.Reflections
        return L.make_params(dt_half=self.DeltaTHalf, gamma=self.Gamma, particle_mass=self.ParticleMass,
                             accel=tuple(self.Acceleration), hor=tuple(self.HorPeriodicity),
                             ver=tuple(self.VertPeriodicity), refl=(r.L, r.R, r.U, r.D), kernel=self.Kernel.id,
                             precision=precision, device=device)


def MakeConfig() -> SphConfig:
    return SphConfig()


# ---- .sph-config reader: just enough of the reference grammar (config-parser.go:486-679) to run the two generated
# example files; the full tokenizer with line:col diagnostics stays in Go (out of scope) ----------------
def MakeConfigFromText(text: str) -> SphConfig:
    conf = MakeConfig()
    title = sub = None
    pending = {}
    src_pending = {}

    def flush_rect():
        nonlocal pending
        if pending:
            if set(pending) != {"NParticles", "UpperLeft", "LowerRight"}:
                raise ValueError(f"[UniformRect] needs NParticles, UpperLeft, LowerRight; got {sorted(pending)}")
            conf.Start.append(UniformRectSpawner(pending["UpperLeft"], pending["LowerRight"], int(pending["NParticles"])))
            pending = {}

    for ln, raw in enumerate(text.splitlines(), 1):
        line = raw.split("//")[0].strip()
        if not line:
            continue
        if line.startswith("[["):
            flush_rect()
            title, sub = line.strip("[]").strip(), None
            if title not in ("Simulation", "Start", "Sources", "Boundaries"):  # config-parser.go:26-31
                raise ValueError(f"line {ln}: `{title}` is not a valid title")
            continue
        if line.startswith("["):
            flush_rect()
            sub = line.strip("[]").strip()
            continue
        tok = line.split()
        name, vals = tok[0], tok[1:]
        num = [float(v) for v in vals] if name != "Kernel" else vals
        key = (title, sub, name)
        if key == ("Simulation", "Config", "NSteps"): conf.NSteps = int(num[0])
        elif key == ("Simulation", "Config", "Gamma"): conf.Gamma = num[0]
        elif key == ("Simulation", "Config", "ParticleMass"): conf.ParticleMass = num[0]
        elif key == ("Simulation", "Config", "DeltaTHalf"): conf.DeltaTHalf = num[0]
        elif key == ("Simulation", "Config", "Acceleration"): conf.Acceleration = (num[0], num[1])
        elif key == ("Simulation", "Config", "Kernel"):
            if vals[0] not in KERNELS_BY_CONFIG_NAME:
                raise ValueError(f"line {ln}: Kernel `{vals[0]}` is not implemented")
            conf.Kernel = KERNELS_BY_CONFIG_NAME[vals[0]]
        elif title == "Simulation" and sub == "Viewport": pass
        elif key == ("Boundaries", "Periodic", "Horizontal"): conf.HorPeriodicity = (num[0], num[1])
        elif key == ("Boundaries", "Periodic", "Vertical"): conf.VertPeriodicity = (num[0], num[1])
        elif title == "Boundaries" and sub == "Reflection" and name in ("Left", "Right", "Up", "Down"):
            setattr(conf.Reflections, name[0], num[0])
        elif title == "Start" and sub == "UniformRect":
            pending[name] = num[0] if name == "NParticles" else (num[0], num[1])
            if len(pending) == 3:
                flush_rect()
        elif title == "Sources" and sub == "Point" and name in ("Pos", "Rate"):  # config-parser.go:384-423: a Pos/Rate pair
            src_pending[name] = (num[0], num[1]) if name == "Pos" else num[0]
            if len(src_pending) == 2:
                conf.Sources.append(PointSource(origin=src_pending["Pos"], rate=src_pending["Rate"]))
                src_pending.clear()
        else:
            raise ValueError(f"line {ln}: unknown parameter {key}")
    flush_rect()
    if src_pending:
        raise ValueError(f"[Point] needs both Pos and Rate; got {sorted(src_pending)}")
    return conf


class Simulation:
    """sim.Simulation (sph.go:14-21): Config is public and mutable; particle state lives on the GPU."""

    def __init__(self, conf: SphConfig, particles: Optional[dict] = None, device: int = 0, capacity: Optional[int] = None,
                 precision: int = 64):
        self.Config = conf
        self.device = device
        if particles is None:
            parts = [s.Spawn(0) for s in conf.Start]  # every spawner re-seeds: same stream per rectangle (config-parser.go:60-64)
            if parts:
                particles = {k: np.concatenate([p[k] for p in parts], axis=0) for k in parts[0]}
            else:
                particles = dict(pos=np.zeros((0, 2)), vel=np.zeros((0, 2)), e=np.zeros(0))
        n = len(particles["pos"])
        # the device id is the spawn index (dense, so the by-id transfers work); Particle.Z (core.go:41, a random
        # 63-bit draw that two spawners can repeat because each re-seeds) is kept on the host: Z[id]
        ids = particles.get("id", np.arange(n, dtype=np.int64))
        self.Z = particles.get("z")
        self._pushed = self._snapshot()
        self.precision = precision  # 64: the reference's arithmetic; 32: the fp32 build (results within 1e-5)
        if capacity is None:  # sph.go:45 reserves 100000 more particles for the sources
            capacity = max(n, 1) + (100000 if conf.Sources else 0)
        self._h = L.Handle(conf.to_params(device, precision), particles["pos"], particles.get("vel"), particles.get("e"),
                           particles.get("rho"), ids, capacity=capacity)

    # --- plumbing
    def _snapshot(self):
        c = self.Config
        r = c.Reflections
        return (c.DeltaTHalf, c.Gamma, c.ParticleMass, tuple(c.Acceleration), c.Kernel.id, tuple(c.HorPeriodicity),
                tuple(c.VertPeriodicity), (r.L, r.R, r.U, r.D))

    def _push_config(self):
        snap = self._snapshot()
        if snap != self._pushed:  # sim.Config is a public field: callers edit it between steps
            self._call(self._h.set_params, self.Config.to_params(self.device, self.precision))
            self._pushed = snap

    @staticmethod
    def _call(fn, *a):
        try:
            return fn(*a)
        except L.SphbError as e:
            if e.code in (L.E_KERNEL, L.E_STATE, L.E_KNN_UNDERFULL) or "open and periodic" in str(e):
                raise SimPanic(str(e)) from e
            raise

    # --- the reference API
    @property
    def CurrentStep(self) -> int:
        return self._h.current_step

    def _spawn_sources(self):
        """sph.go:72-86: every source spawns at t = CurrentStep * dtHalf * 2; the new particles join the state (the
        reference re-makes its tree; here the next evaluation sorts them in).  Ids continue the spawn index."""
        t = float(self.CurrentStep) * self.Config.DeltaTHalf * 2
        for src in self.Config.Sources:
            new = src.Spawn(t)
            k = len(new["pos"])
            if k:
                n = len(self)
                self._call(self._h.append, new["pos"], new["vel"], new["e"], new.get("rho"), np.arange(n, n + k, dtype=np.int64))
                if self.Z is not None:
                    self.Z = np.concatenate([self.Z, new["z"]])

    def Step(self):
        self._push_config()
        self._spawn_sources()
        if len(self) == 0:
            raise SimPanic("int Run(): Simulation not initialized!")  # sph.go:92-94
        self._call(self._h.step, 1)

    def Run(self):
        if self.Config.Sources:  # sources spawn between steps (sph.go:56-61 calls Step NSteps times)
            for _ in range(self.Config.NSteps):
                self.Step()
            return
        self._push_config()
        self._call(self._h.step, self.Config.NSteps)

    def CalculateForces(self):
        self._push_config()
        self._call(self._h.calc_forces)

    def TotalEnergy(self) -> float:
        return self._call(self._h.reduce, L.SUM_E)

    def TotalDensity(self) -> float:
        return self._call(self._h.reduce, L.SUM_RHO)

    def TotalMomentum(self) -> float:
        return self._call(self._h.reduce, L.LAST_VEL_NORM)  # the `=` instead of `+=` of sph.go:460 is kept

    def FindNearestNeighboursPeriodic(self, HorPeriodic: Sequence[float], VertPeriodic: Sequence[float]):
        """batch form of `for i: Particles[i].FindNearestNeighboursPeriodic(root, hor, ver)` (density.go:67-69)"""
        self._call(self._h.knn, tuple(HorPeriodic), tuple(VertPeriodic))

    def FindNearestNeighbours(self):
        self._call(self._h.knn, L.OPEN, L.OPEN)

    def Density2D(self, kernel: Kernel):
        """batch form of `for i: Particles[i].Rho = Density2D(&p, sim, kernel)` (density.go:71-72)"""
        self._push_config()
        self._call(self._h.density, kernel.id)

    def Particles(self, fields=("pos", "vel", "rho", "c", "e", "h", "id"), sort_by_id=True) -> dict:
        """sim.Root.Particles as SoA numpy arrays: Pos, Vel, Rho, C, E, ..., NNDists[0] (= h); `nn_idx`, `nn_dist`,
        `nn_pos` fill NearestNeighbours / NNDists / NNPos on request only (descending distance, slot 0 = h)."""
        return self._call(self._h.state, tuple(fields), sort_by_id)

    def FrameData(self, width: int = 1280, height: int = 720, by_id: bool = False) -> dict:
        """what (*Animator).CurrentFrame derives per particle (animator.go:75-101), computed on the device: `xy`
        float32 pixel coordinates, `colour` uint8 ramp index, `id` (Z; None with by_id, where element k belongs
        to the particle with id k)"""
        return self._call(self._h.frame, width, height, not by_id)

    def __len__(self):
        return self._h.n

    def Close(self):
        self._h.close()


def MakeSimulation(device: int = 0) -> Simulation:
    conf = MakeConfig()
    sim = Simulation(conf, MakeUniformRectSpawner().Spawn(0), device)
    return sim


def MakeSimulationFromConf(conf: SphConfig, device: int = 0) -> Simulation:
    return Simulation(conf, None, device)


def MakeSimulationFromConfig(path: str, device: int = 0) -> Simulation:
    with open(path) as f:
        return MakeSimulationFromConf(MakeConfigFromText(f.read()), device)
