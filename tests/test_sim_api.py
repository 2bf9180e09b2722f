"""The host-side mirror of the Go `sim` API (sphugo_b200/sim.py): config marshalling on CPU, behaviour on GPU."""
import numpy as np
import pytest

from sphugo_b200 import sim, _lib as L

# same parameter values as the reference's generated example.sph-config (config-parser.go:872-924), own text
EXAMPLE = """
// dam-break-like column, periodic in x, floor at the bottom
[[Simulation]]
[Config]
NSteps        1000
Gamma         4.666
ParticleMass  1000000.0
Acceleration  0  0.55
DeltaTHalf    0.00324
Kernel        Wendtland
[[Start]]
[UniformRect]
NParticles  260
UpperLeft   0.6   0.2
LowerRight  0.79  0.3
[UniformRect]
NParticles  700
UpperLeft   0.27  0.3
LowerRight  0.4   0.9
[[Boundaries]]
[Periodic]
Horizontal  0.2   0.8
Vertical    -100  100
[Reflection]
Down        0.99
[[Simulation]]
[Viewport]
UpperLeft   0 0
LowerRight  1 1
"""


def test_makeconfig_defaults_match_reference():
    c = sim.MakeConfig()  # config-parser.go:131-149
    assert (c.Gamma, c.NSteps, c.DeltaTHalf, c.ParticleMass, c.Kernel) == (1.66666, 10000, 0.001, 1.0, sim.Monahan2D)
    assert c.HorPeriodicity == (-sim.MaxFloat64, sim.MaxFloat64) and c.Reflections.D == sim.MaxFloat64
    p = c.to_params()
    assert p.kernel == L.KERNEL_MONAGHAN and p.hor[0] == L.OPEN_LO and p.refl_R == L.OPEN_HI
    s = sim.MakeUniformRectSpawner()
    assert s.NParticles == 1000 and s.LowerRight == (1.0, 1.0)
    sp = s.Spawn(0)
    assert sp["pos"].shape == (1000, 2) and (sp["e"] == 0.01).all() and (sp["vel"] == 0).all()


def test_sph_config_text():
    c = sim.MakeConfigFromText(EXAMPLE)
    assert c.Kernel == sim.Wendtland2D and c.Gamma == 4.666 and c.ParticleMass == 1e6
    assert c.Acceleration == (0.0, 0.55) and c.DeltaTHalf == 0.00324
    assert c.HorPeriodicity == (0.2, 0.8) and c.VertPeriodicity == (-100.0, 100.0)
    assert c.Reflections.D == 0.99 and c.Reflections.L == -sim.MaxFloat64
    assert [(s.NParticles, s.UpperLeft, s.LowerRight) for s in c.Start] == [(260, (0.6, 0.2), (0.79, 0.3)), (700, (0.27, 0.3), (0.4, 0.9))]
    with pytest.raises(ValueError):
        sim.MakeConfigFromText("[[Simulation]]\n[Config]\nKernel Gauss\n")
    with pytest.raises(ValueError):
        sim.MakeConfigFromText("[[Nonsense]]\n")


@pytest.mark.gpu
def test_simulation_api_against_oracle():
    from oracle import oracle as orc
    from tests import util as U
    from sphugo_b200 import gen
    conf = sim.MakeConfigFromText(EXAMPLE)
    s = sim.Simulation(conf, gen.spawn([(sp.NParticles, sp.UpperLeft, sp.LowerRight) for sp in conf.Start]))
    assert len(s) == 960
    p0 = s.Particles(("pos", "vel", "e", "id"))
    kw = dict(gamma=4.666, particle_mass=1e6, accel=(0.0, 0.55), dt_half=0.00324, kernel=2, hor=(0.2, 0.8),
              ver=(-100.0, 100.0), refl=(orc.OPEN[0], orc.OPEN[1], orc.OPEN[0], 0.99))
    o = orc.Oracle(orc.make_params(**kw), p0["pos"], p0["vel"], p0["e"], None, p0["id"])
    for _ in range(3):
        s.Step()
    # exact-kNN mode: at step 3 the reference's own tree walk returns one non-nearest neighbour on this input
    # (non-enclosing circle merge, core.go:300-311); the GPU computes exact kNN
    o.step(3, knn_mode=1)
    a, b = s.Particles(), o.state()
    assert s.CurrentStep == 3
    assert np.abs(a["pos"] - b["pos"]).max() <= 1e-9
    assert U.rel_err(a["rho"], b["rho"]) <= 1e-9
    assert abs(s.TotalEnergy() - o.total_energy()) <= 1e-9 * abs(o.total_energy())
    # Config is a public mutable field (sph.go:15): change gravity between steps like a caller would
    s.Config.Acceleration = (0.0, 0.1)
    kw["accel"] = (0.0, 0.1)
    o.set_params(orc.make_params(**kw))
    s.Step(); o.step(1, knn_mode=1)
    assert np.abs(s.Particles()["pos"] - o.state()["pos"]).max() <= 1e-9
    # TopHat has no derivative: Step panics in the reference (sph.go:251-253)
    s.Config.Kernel = sim.TopHat2D
    with pytest.raises(sim.SimPanic):
        s.Step()
    s.Close()


@pytest.mark.gpu
def test_density_example_flow():
    """examples/density main (density.go:41-97) through the mirrored API"""
    from sphugo_b200 import gen
    ic = gen.spawn([(1000, (0.0, 0.0), (1.0, 1.0)), (200, (0.1, 0.0), (0.3, 0.4))])
    s = sim.Simulation(sim.MakeConfig(), ic)
    s.FindNearestNeighboursPeriodic((0, 1), (0, 1))
    rho = {}
    for k in (sim.TopHat2D, sim.Monahan2D, sim.Wendtland2D):
        s.Density2D(k)
        rho[k.name] = s.Particles(("rho",))["rho"]
    h = s.Particles(("h",))["h"]
    assert np.allclose(rho["TopHat2D"], 32 / (np.pi * h * h), rtol=1e-13)  # DensityTopHat2D, sph.go:231-234
    assert 0.5 < np.median(rho["Monahan2D"] / rho["TopHat2D"]) < 2
    with pytest.raises(sim.SimPanic):
        s.FindNearestNeighboursPeriodic((-sim.MaxFloat64, 1.0), (0, 1))
    s.Close()


SOURCES = """[[Start]]
[UniformRect]
NParticles 10
UpperLeft 0 0
LowerRight 1 1
[[Sources]]
[Point]
Pos 0.2 0.2
Rate 100
[Point]
Rate 10
Pos 0.5 0.5
"""


def test_point_sources_parse_and_spawn_from_the_running_go_stream():
    """[[Sources]] [Point] (config-parser.go:384-423) and PointSource.Spawn (config-parser.go:82-102)"""
    from sphugo_b200 import gorand
    c = sim.MakeConfigFromText(SOURCES)
    assert [(s.origin, s.rate) for s in c.Sources] == [((0.2, 0.2), 100.0), ((0.5, 0.5), 10.0)]
    c.Start[0].Spawn(0)  # re-seeds the package-level stream and draws 20 uniforms + 10 Z
    src = c.Sources[0]
    assert len(src.Spawn(0.0)["pos"]) == 0  # t = 0: nothing yet
    new = src.Spawn(0.035)  # int(0.035 * 100) = 3 particles, LastSpwned advances by 3 cooldowns
    assert new["pos"].shape == (3, 2) and abs(src.LastSpwned - 0.03) < 1e-15
    assert (new["rho"] == 100).all() and (new["e"] == 0.002).all()
    r = gorand.Rand(12345678)
    r.Float64s(20); r.Ints(10)
    dy = 0.01 * (-1 + 2 * r.Float64()); dx = 0.01 * (-1 + 2 * r.Float64())  # dy is drawn first (config-parser.go:91-92)
    assert tuple(new["pos"][0]) == (0.2 + dx, 0.2 + dy) and new["z"][0] == r.Int()
    with pytest.raises(ValueError):
        sim.MakeConfigFromText("[[Sources]]\n[Point]\nPos 0.1 0.1\n")


def test_config_text_spawns_the_golden_scene():
    """the reader + the spawners on the Go stream give exactly the particles tests/golden/c2_example_config_go.npz was made from"""
    import os
    c = sim.MakeConfigFromText(EXAMPLE)
    pos = np.concatenate([s.Spawn(0)["pos"] for s in c.Start])
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "c2_example_config_go.npz"))
    assert np.array_equal(pos, g["pos0"])
