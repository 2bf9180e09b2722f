#!/usr/bin/env python
"""bench.py — particle-updates/s of the full SPH step (sort, kNN k=32, density, force, leapfrog) on B200.

Contract (driver):  python bench.py --gpus N --steps K --warmup W [--impl reference]
prints ONE JSON line on rank 0.  A "step" is one (*Simulation).Step() (reference sim/sph.go:64-198) over
all particles.  Workload (see DESIGN.md "Measurement"): weak scaling, 2^25 jittered-lattice particles per
GPU in a periodic box (BASELINE.json configs[4]: 256 M on 8 GPUs = 32 M per GPU), fp64.  --workload c3
selects the 2^20-particle single-GPU box (configs[2]).

value   : device-resident state, CUDA events around K asynchronous sphb_step calls (max over ranks)
e2e     : the same steps through the C ABI with HOST buffers: per step upload of the caller-set fields (Pos, Vel, E)
          from pinned memory, sphb_step, download of the frame data the per-step consumer draws (sphb_frame) + sum E
roofline: whole step as the dominant "kernel" chain, algorithmic bytes 652 B/particle, plus
          per-phase fractions from the library's CUDA-event phase timers
cpu_baseline / --impl reference: the CPU restatement of the reference Go path (oracle/, 1 core: package
          sim is serial) on a bounded sample of the same lattice.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# algorithmic bytes / particle-update, SURVEY §8d: 300 + 44 s with s = sizeof(real)
B_ALG_BY_PREC = {64: {"keys": 40, "sort": 16, "reorder": 148, "knn": 176, "force": 272},
                 32: {"keys": 24, "sort": 16, "reorder": 84, "knn": 152, "force": 200}}
B_ALG = B_ALG_BY_PREC[64]
B_ALG_TOTAL = 652
_JSON_OUT = None


def emit(line: dict):
    """the one JSON line, on the process's ORIGINAL stdout (main() points fd 1 at stderr for everything else; the saved
    descriptor travels through the environment because slab.py imports this file as a second module object)"""
    fd = os.environ.get("SPHB_BENCH_JSON_FD")
    data = (json.dumps(line) + "\n").encode()
    if fd:
        os.write(int(fd), data)
    else:
        sys.stdout.write(data.decode())
        sys.stdout.flush()


METRIC = "particle-updates/s per SPH step (k=32)"
UNIT = "particle-updates/s"


def ncu_traffic(kernel_key):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel at this workload, from the
    committed `ncu --set full` capture (profiles/traffic.json, written by tools/ncu_traffic.py); None if absent"""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    try:
        return json.load(open(p)).get(kernel_key)
    except Exception:
        return None


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def workload(name: str, n_gpus: int):
    """(nx, ny, box, physics kwargs, description): an nx x ny jittered lattice filling the periodic box
    [0, box[0]] x [0, box[1]].  Lattice spacing fixes h, so dt scales with it."""
    phys = dict(accel=(0.0, 0.2), gamma=1.66666, particle_mass=1.0, kernel=1)
    if name == "c3":
        nx = ny = 1024
        box = (1.0, 1.0)
        phys["dt_half"] = 0.001
        desc = "C3-P: 2^20 jittered-lattice particles, periodic [0,1]^2, Monaghan, g=(0,0.2), fp64"
    elif name == "c5":
        # weak scaling at C5's lattice spacing 2^-14: 2^25 sites per GPU; 8 GPUs = the 16384^2 box of configs[4]
        nx, ny = {1: (8192, 4096), 2: (8192, 8192), 4: (16384, 8192), 8: (16384, 16384)}[n_gpus]
        box = (nx / 16384.0, ny / 16384.0)
        phys["dt_half"] = 6e-5
        desc = (f"C5 share (BASELINE configs[4]): 2^25 jittered-lattice particles per GPU x {n_gpus} GPU(s) = {nx}x{ny} "
                f"sites at spacing 2^-14, periodic box {box[0]}x{box[1]}, Monaghan, g=(0,0.2), fp64")
    elif name == "c4":
        # BASELINE configs[3]: 2^24 particles, shock tube (number-density ratio 4:1 across x = 0.5), strong scaling
        nx = ny = 4096
        box = (1.0, 1.0)
        phys["dt_half"] = 2.5e-4
        phys["accel"] = (0.0, 0.0)
        desc = ("C4 (BASELINE configs[3]): 2^24-particle shock tube, jittered lattices with number-density ratio 4:1 left/right "
                f"of x = 0.5, periodic [0,1]^2, Monaghan, g = 0, fp64; {n_gpus} GPU(s), equal-count x-slabs")
    else:
        raise SystemExit(f"unknown workload {name}")
    return nx, ny, box, phys, desc


class ClockSampler:
    """nvidia-smi clocks during the timed region (B200_PROFILING.md recipe)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "50", "-i", str(self.index)], stdout=subprocess.PIPE, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append(ln)

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=3)
        except Exception:
            self.proc.kill()
        if not any(len(ln.split(",")) >= 8 for ln in self.lines):
            # the timed region was shorter than nvidia-smi's first report: one synchronous sample right after it
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i", str(self.index)],
                                     capture_output=True, text=True, timeout=10).stdout
                self.lines += out.splitlines()
            except Exception:
                pass
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def cpu_baseline(steps=2, nx=512):
    """Reference CPU path (oracle port of the Go code, faithful tree-walk mode) on a bounded sample: a
    (nx x nx)-site periodic jittered lattice of the same kind, 1 thread (package sim is serial)."""
    from oracle import oracle as orc
    from sphugo_b200 import gen
    pos = gen.jittered_lattice(nx, nx)
    n = len(pos)
    dt = 0.001 * (1024.0 / nx)
    po = orc.make_params(hor=(0.0, 1.0), ver=(0.0, 1.0), accel=(0.0, 0.2), dt_half=dt)
    o = orc.Oracle(po, pos, None, np.full(n, 0.01))
    o.step(1)  # step 0 evaluates the forces twice (sph.go:89-103): excluded like in the GPU timing
    t0 = time.perf_counter()
    o.step(steps)
    dt_s = time.perf_counter() - t0
    o.close()
    return n * steps / dt_s, n, dt_s / steps


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    nx, ny, box, phys, desc = workload(args.workload, args.gpus)
    t_all = time.perf_counter()
    # The reference arm runs BASELINE configs[2] in full - the 2^20-particle C3-P box, the largest BASELINE configuration
    # a serial CPU path steps in seconds - with its own dt: the repo arm's `legs` hold the same configuration (leg c3p,
    # f64), so that one published pair is like for like.  Against the headline workload (2^25 particles per GPU) the
    # figure is an extrapolation that flatters the CPU (its cost per particle grows ~ log N).
    sample_nx = 1024
    from oracle import oracle as orc
    from sphugo_b200 import gen
    pos = gen.jittered_lattice(sample_nx, sample_nx)
    n_s = len(pos)
    po = orc.make_params(hor=(0.0, 1.0), ver=(0.0, 1.0), accel=(0.0, 0.2), dt_half=0.001)
    o = orc.Oracle(po, pos, None, np.full(n_s, 0.01))
    Wr = min(args.warmup, 1)
    o.step(1 + Wr)  # step 0 evaluates the forces twice (sph.go:89-103): excluded, like in the repo arm
    K = max(1, min(args.steps, 3))
    t0 = time.perf_counter()
    o.step(K)
    el = time.perf_counter() - t0
    o.close()
    v = n_s * K / el
    line = {
        "metric": METRIC, "value": v, "unit": UNIT, "impl": "reference", "n_gpus": args.gpus, "steps": K,
        "warmup": 1 + Wr, "ms_per_step": el / K * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": desc,
                   "sample": f"BASELINE configs[2] in full: {sample_nx}x{sample_nx} periodic jittered lattice ({n_s} particles), dt_half 0.001, "
                             "g=(0,0.2), Monaghan - identical to leg c3p / f64 of the repo arm (same_config there); an extrapolation "
                             "against the headline workload"},
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": 1, "kind": "port",
                         "sample": f"{K} Step() calls on {n_s} particles; C restatement of the serial Go path "
                                   "(Go toolchain absent here and on the GPU box, profiles/r02_go_probe.txt; package sim starts no "
                                   "goroutines so GOMAXPROCS is irrelevant)"},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0, "wall_s": time.perf_counter() - t_all,
    }
    emit(line)


def make_ic_c4(rank, world):
    """this rank's equal-count x-slab of the 2^24-particle shock tube: (pos, ids, bounds)"""
    from sphugo_b200 import gen, slab
    pos = gen.shock_tube(1 << 24)
    bounds = slab.equal_count_bounds(pos[:, 0], world, 0.0, 1.0)
    m = (pos[:, 0] >= bounds[rank]) & (pos[:, 0] < bounds[rank + 1])
    return np.ascontiguousarray(pos[m]), np.nonzero(m)[0].astype(np.int64), bounds


def make_ic(nx, ny, box, rank, world):
    """this rank's x-slab (equal site count) of the global nx x ny jittered lattice on [0,box]"""
    from sphugo_b200 import gen
    cols = nx // world
    sx = box[0] / nx
    x0 = rank * cols * sx
    pos = gen.jittered_lattice(cols, ny, (x0, 0.0), (x0 + cols * sx, box[1]), 0.25, seed=gen.DEFAULT_SEED + rank)
    return pos, (x0, x0 + cols * sx)


# ---- additional legs of the default run: the other BASELINE.json configurations on this GPU -----------------------
def leg_ic(name):
    """(pos, params kwargs, description, e0) of a single-GPU leg"""
    from sphugo_b200 import gen, gorand
    mon = dict(gamma=1.66666, particle_mass=1.0, kernel=1)
    if name == "c5":
        return (gen.jittered_lattice(8192, 4096, (0.0, 0.0), (0.5, 0.25)),
                dict(mon, hor=(0.0, 0.5), ver=(0.0, 0.25), accel=(0.0, 0.2), dt_half=6e-5),
                "C5 share (the headline workload): 2^25 jittered-lattice particles, periodic box 0.5x0.25, Monaghan, g=(0,0.2)")
    if name == "c3p":
        return (gen.jittered_lattice(1024, 1024), dict(mon, hor=(0.0, 1.0), ver=(0.0, 1.0), accel=(0.0, 0.2), dt_half=0.001),
                "C3-P (BASELINE configs[2]): 2^20 jittered-lattice particles, periodic [0,1]^2, Monaghan, g=(0,0.2)")
    if name == "c3u":
        return (gen.uniform_rect(1 << 20), dict(mon, hor=(0.0, 1.0), ver=(0.0, 1.0), accel=(0.0, 0.2), dt_half=0.001),
                "C3-U (BASELINE configs[2]): 2^20 i.i.d. uniform particles, periodic [0,1]^2, Monaghan, g=(0,0.2)")
    if name == "c4":
        return (gen.shock_tube(1 << 24), dict(mon, hor=(0.0, 1.0), ver=(0.0, 1.0), accel=(0.0, 0.0), dt_half=2.5e-4),
                "C4 shock tube (BASELINE configs[3]): 2^24 particles, number-density ratio 4:1 across x = 0.5, periodic [0,1]^2, Monaghan")
    if name == "c4dam":
        from sphugo_b200 import _lib as L
        return (gen.dam_break(1 << 24), dict(gamma=1.66666, particle_mass=1.0, kernel=2, accel=(0.0, 0.55), dt_half=1.25e-4,
                                             refl=(0.0, 1.0, L.OPEN_LO, 1.0)),
                "C4 dam break (BASELINE configs[3]): 2^24 particles filling [0,0.25]x[0.5,1], open box, reflections L 0 / R 1 / D 1, "
                "g=(0,0.55), Wendland (the physics of the reference's tube config, config-parser.go:926-973)")
    if name == "speed":
        return (gorand.uniform_rect_spawn(100000)["pos"], dict(mon, accel=(0.0, 0.2), dt_half=0.02),
                "examples/speed-test (speed-test.go:22-45): 100000 particles of the Go math/rand stream, open box, dt_half 0.02, g=(0,0.2)")
    raise SystemExit(f"unknown leg {name}")


def run_leg(name, precision, K, W, device, fresh=False, flags=0):
    """device-resident throughput of one leg; fresh = time the first K steps of a new simulation (speed-test's protocol)"""
    import torch
    from sphugo_b200 import _lib as L
    pos, kw, desc = leg_ic(name)
    n = len(pos)
    g = L.Handle(L.make_params(precision=precision, device=device, flags=flags, **kw), pos, None, np.full(n, 0.01))
    del pos
    ext = torch.cuda.ExternalStream(g.stream, device=torch.device("cuda", device))
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if not fresh:
        g.step(1 + W)
    g.sync()
    c0 = g.counters()
    t0 = time.perf_counter()
    ev0.record(ext)
    g.step(K)
    ev1.record(ext)
    g.sync()
    wall = time.perf_counter() - t0
    ms = ev0.elapsed_time(ev1) / K
    c1 = g.counters()
    sum_e = g.reduce(L.SUM_E)
    g.close()
    peak, _ = measured_peak()
    b = sum(B_ALG_BY_PREC[precision].values())
    return {"leg": name + ("+reuse" if flags & 2 else ""), "workload": desc, "particles": n, "dtype": "f64" if precision == 64 else "f32", "steps": K,
            "ms_per_step": ms, "value": n / (ms * 1e-3), "unit": UNIT, "wall_ms_per_step": wall / K * 1e3,
            "step_roofline_frac": n * b / (ms * 1e-3) / 1e9 / peak, "alg_bytes_per_particle": b,
            "fallback_fraction": (c1["knn_fallback"] - c0["knn_fallback"]) / (n * K),
            "reuse_steps": c1["reuse_steps"] - c0["reuse_steps"], "sum_E": sum_e,
            "protocol": "first K steps of a fresh simulation (step 0 evaluates the forces twice)" if fresh else f"K steps after step 0 + {W} warm-up steps"}


def speed_test_cpu(steps):
    """the reference arm of the speed-test leg: the oracle on the IDENTICAL input, full size, from a fresh simulation"""
    from oracle import oracle as orc
    from sphugo_b200 import gorand
    ic = gorand.uniform_rect_spawn(100000)
    o = orc.Oracle(orc.make_params(dt_half=0.02, accel=(0.0, 0.2)), ic["pos"], ic["vel"], ic["e"])
    t0 = time.perf_counter()
    o.step(steps)
    el = time.perf_counter() - t0
    e = o.total_energy()
    o.close()
    return {"value": 100000 * steps / el, "unit": UNIT, "ms_per_step": el / steps * 1e3, "steps": steps, "cores": 1, "kind": "port", "sum_E": e}


def run_legs(args, device):
    legs = []
    for name, precs in (("c3p", (64, 32)), ("c3u", (64, 32)), ("c4", (64, 32)), ("c4dam", (64,))):
        for prec in precs:
            try:
                legs.append(run_leg(name, prec, 20 if name.startswith("c3") else 10, 6, device))
            except Exception as ex:  # a leg must not take the headline line down with it
                legs.append({"leg": name, "dtype": f"f{prec}", "error": str(ex)[:300]})
    # the headline workload with the certified list reuse switched on (SPHB_FLAG_REUSE_LISTS; off by default): long
    # enough for the schedule to settle (rebuild + reuse cycles), same device-resident protocol
    for prec in (64, 32):
        try:
            legs.append(run_leg("c5", prec, 42, 6, device, flags=2))
        except Exception as ex:
            legs.append({"leg": "c5+reuse", "dtype": f"f{prec}", "error": str(ex)[:300]})
    try:
        run_leg("speed", 64, 20, 0, device, fresh=True)  # (the first run warms the device up)
        sp = run_leg("speed", 64, 20, 0, device, fresh=True)
        if not args.no_cpu:
            cpu = speed_test_cpu(20)
            sp["cpu_same_input"] = cpu
            sp["same_config"] = True
            sp["ratio_vs_cpu_same_input"] = sp["value"] / cpu["value"]
        legs.append(sp)
    except Exception as ex:
        legs.append({"leg": "speed", "error": str(ex)[:300]})
    return legs


def run_ours(args):
    import torch
    from sphugo_b200 import _lib as L
    from sphugo_b200 import build as B
    B.build()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; libsphb has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    nx, ny, box, phys, desc = workload("c4" if args.workload == "c4dam" else args.workload, world)
    global B_ALG, B_ALG_TOTAL
    B_ALG = B_ALG_BY_PREC[args.precision]
    B_ALG_TOTAL = sum(B_ALG.values())
    phys["precision"] = args.precision
    if args.reuse:
        phys["flags"] = 2
    if args.precision == 32:
        desc = desc.replace("fp64", "fp32 build (fp32 pair arithmetic, fp64 state)")
    if world > 1:
        from sphugo_b200 import slab
        return slab.bench(args, nx, ny, box, phys, desc, rank, world, local)

    if args.workload in ("c4", "c4dam"):
        raise SystemExit("single-GPU C4 numbers are legs of the default run (bench.py without --workload); --workload c4 / c4dam is for --gpus N")
    else:
        pos, _ = make_ic(nx, ny, box, 0, 1)
    n = len(pos)
    prm = L.make_params(hor=(0.0, box[0]), ver=(0.0, box[1]), device=local, **phys)
    e0 = np.full(n, 0.01)
    g = L.Handle(prm, pos, None, e0)
    del pos
    K, W = args.steps, max(args.warmup, 3)
    g.step(1 + W)  # step 0 (double force evaluation) + W warm-up steps, untimed
    g.sync()
    c0 = g.counters()
    sampler = ClockSampler(local)
    sampler.start()
    ext = torch.cuda.ExternalStream(g.stream, device=torch.device("cuda", local))
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    ev0.record(ext)          # CUDA events on the stream the kernels are launched on
    g.step(K)                # asynchronous: enqueues K full steps
    ev1.record(ext)
    g.sync()
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    dev_ms = ev0.elapsed_time(ev1)
    clocks = sampler.stop()
    c1 = g.counters()
    ms_per_step = dev_ms / K
    # per-phase split from the library's CUDA-event phase timers (separate, untimed steps), by kind of evaluation:
    # a rebuild (sort, reorder, tile search [+ annulus pass when it starts a reuse cycle]) or a reuse evaluation
    # (predict in the "reorder" slot, exact kNN from the stored candidates in the "knn" slot)
    KP = max(min(K, 14), 2)
    kinds = {"rebuild": [], "reuse": []}
    for _ in range(KP):
        r0 = g.counters()["reuse_steps"]
        g.step(1)
        pt = g.phase_times()
        kinds["reuse" if g.counters()["reuse_steps"] > r0 else "rebuild"].append(pt)
    value = n * K / (dev_ms * 1e-3)
    peak, peak_src = measured_peak()
    step_achieved = n * B_ALG_TOTAL / (ms_per_step * 1e-3) / 1e9
    n_reuse_timed = c1["reuse_steps"] - c0["reuse_steps"]
    share = {"reuse": n_reuse_timed / K, "rebuild": 1.0 - n_reuse_timed / K}  # of the timed region
    if not kinds["reuse"]:
        share = {"reuse": 0.0, "rebuild": 1.0}
    mean_ms = {kind: {k: float(np.mean([p[k] for p in v])) for k in L.PHASES} for kind, v in kinds.items() if v}
    phases = {}
    for k in ("keys", "sort", "reorder", "knn", "force"):
        # average launch time of the phase over the timed region's mix of evaluations
        ms = sum(share[kind] * mean_ms[kind][k] for kind in mean_ms)
        # a phase fused into another kernel (keys: emitted by the force epilogue on periodic steps) has no launch of
        # its own: the timer brackets nothing and a bandwidth figure would be meaningless
        fused = ms < 0.02 and k in ("keys", "sort")
        gbs = n * B_ALG[k] / (ms * 1e-3) / 1e9 if ms > 0 and not fused else None
        phases[k] = {"ms": ms, "alg_GBps": gbs, "frac": gbs / peak if gbs else None}
        if fused:
            phases[k]["fused_into"] = "force epilogue of the previous step (keys) / not run by reuse evaluations"
    phases["by_kind_ms"] = mean_ms
    phases["mix_of_timed_region"] = share
    dom = max(("keys", "sort", "reorder", "knn", "force"), key=lambda k: phases[k]["ms"])
    dom_kernel = {"knn": "k_knn_tile + k_knn_annulus (rebuilds) / k_knn_reuse (reuse evaluations), + k_knn_fallback for refused particles",
                  "force": "k_force_st", "reorder": "k_reorder / k_predict", "keys": "k_keys", "sort": "counting-sort kernels"}[dom]
    traffic = ncu_traffic(f"{dom}_{args.workload}_f{args.precision}")

    h2d = d2h = 0
    e2e_val, e2e_serial, Ke = 0.0, 0.0, 0
    if not args.no_e2e:
        # ---- e2e: host buffers through the C ABI, every step: upload state, step, download results
        import ctypes as C
        # inputs: the particle fields a caller sets (spawners / examples write Pos, Vel, E: config-parser.go:68-77,
        # density.go:12-15); outputs: what the per-step consumer needs - the frame data of (*Animator).CurrentFrame
        # (pixel coordinates, colour-ramp index, Z: animator.go:60-101), extracted on the device (sphb_frame) - and
        # the energy sum simviewer plots (simviewer.go:302).  VDot / EDot of the previous step stay on the device.
        # The caller keeps its particles in its own fixed order (element k = particle id k): sphb_upload_by_id and
        # sphb_frame(id_out = NULL) translate to and from the device's cell order.
        up_fields = ["pos", "vel", "e"]
        st = g.download(up_fields + ["id"])
        host = {}
        for f in up_fields:
            shp, dt = L.FIELD_SHAPE[f]
            host[f] = torch.empty((n,) + shp, dtype=torch.float64).pin_memory().numpy()
            host[f][st["id"]] = st[f]  # the current state, in id order
        del st
        fr = {"xy": torch.empty((n, 2), dtype=torch.float32).pin_memory().numpy(),
              "colour": torch.empty((n,), dtype=torch.uint8).pin_memory().numpy()}
        h2d = sum(host[f].nbytes for f in up_fields)
        d2h = sum(a.nbytes for a in fr.values()) + 8
        Ke = max(3, min(K, 10))
        # (a) serial protocol: every call returns before the next one starts
        for _ in range(2):
            g.upload_by_id(**host); g.step(1); g.frame(1280, 720, ids=False, out=fr); g.reduce(L.SUM_E)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(Ke):
            g.upload_by_id(**host)
            g.step(1)
            g.frame(1280, 720, ids=False, out=fr)
            g.reduce(L.SUM_E)
        torch.cuda.synchronize()
        e2e_serial = n * Ke / (time.perf_counter() - t0)
        # (b) the same bytes per step with the upload split in two (sphb_upload_by_id_begin / _end): the inputs of step
        # k + 1 cross the bus while step k runs.  Every iteration uploads one step's inputs, runs one step and downloads
        # one step's results; a step consumes the inputs uploaded during the previous iteration.
        g.upload_by_id(**host)
        for _ in range(2):
            g.step(1); g.upload_by_id_begin(**host); g.frame(1280, 720, ids=False, out=fr); g.reduce(L.SUM_E); g.upload_by_id_end()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(Ke):
            g.step(1)
            g.upload_by_id_begin(**host)
            g.frame(1280, 720, ids=False, out=fr)
            g.reduce(L.SUM_E)
            g.upload_by_id_end()
        torch.cuda.synchronize()
        e2e_s = time.perf_counter() - t0
        e2e_val = n * Ke / e2e_s

    g.close()

    other = None
    if not args.no_other_build:
        # the other build of the library on the same workload (device-resident steps only), for the record
        op = 32 if args.precision == 64 else 64
        prm2 = L.make_params(hor=(0.0, box[0]), ver=(0.0, box[1]), device=local, **dict(phys, precision=op))
        pos2 = make_ic(nx, ny, box, 0, 1)[0]
        g2 = L.Handle(prm2, pos2, None, e0)
        del pos2
        g2.step(1 + W)
        g2.sync()
        ext2 = torch.cuda.ExternalStream(g2.stream, device=torch.device("cuda", local))
        a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a0.record(ext2)
        g2.step(K)
        a1.record(ext2)
        g2.sync()
        torch.cuda.synchronize()
        ms2 = a0.elapsed_time(a1) / K
        b2 = sum(B_ALG_BY_PREC[op].values())
        other = {"dtype": "f32" if op == 32 else "f64", "value": n / (ms2 * 1e-3), "unit": UNIT, "ms_per_step": ms2,
                 "step_roofline_frac": n * b2 / (ms2 * 1e-3) / 1e9 / peak, "alg_bytes_per_particle": b2,
                 "tolerance_vs_reference": 1e-5 if op == 32 else 1e-12}
        g2.close()

    cb_v, cb_n, cb_s = cpu_baseline(steps=2, nx=512) if not args.no_cpu else (None, 0, 0)
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": 1, "steps": K, "warmup": W,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong" if args.workload == "c4" else "weak", "vs_baseline": None,
        "dtype": "f64" if args.precision == 64 else "f32", "data": "synthetic",
        "config": {"workload": desc, "particles": n, "l2": "state (>= 280 B/particle) exceeds the 126 MB L2; no flush needed",
                   "timing": "CUDA events on the library stream around K asynchronous steps", "wall_ms_per_step": wall / K * 1e3},
        # dominant kernel: algorithmic bytes of its phase x particles per launch / its launch time (CUDA
        # events of the library's phase timers); "step" is the same for the whole 5-phase chain (north_star's figure)
        "roofline": {"bound": "hbm", "achieved": phases[dom]["alg_GBps"], "peak": peak, "unit": "GB/s",
                     "frac": phases[dom]["frac"], "traffic": traffic, "peak_source": peak_src, "kernel": dom_kernel,
                     "alg_bytes_per_particle": B_ALG[dom], "launch_ms": phases[dom]["ms"],
                     "step": {"achieved": step_achieved, "frac": step_achieved / peak, "alg_bytes_per_particle": B_ALG_TOTAL},
                     "phases": phases},
        "cpu_baseline": None if cb_v is None else {
            "value": cb_v, "unit": UNIT, "cores": 1, "kind": "port",
            "sample": f"2 Step() calls on a {cb_n}-particle periodic jittered lattice ({cb_s:.2f} s/step); C restatement of the serial Go path"},
        "e2e": {"value": e2e_val, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "steps": Ke,
                "what": "per step, through the C ABI with pinned host buffers, wall clock: sphb_step(1) on the inputs uploaded during the previous "
                        "iteration | sphb_upload_by_id_begin(Pos,Vel,E of the next step: copies overlap the running step) | sphb_frame(xy f32, "
                        "colour u8, caller's order) | sphb_reduce(sum E) | sphb_upload_by_id_end",
                "serial": {"value": e2e_serial, "unit": UNIT,
                           "what": "the same bytes with every call finished before the next starts: sphb_upload_by_id + sphb_step(1) + sphb_frame + sphb_reduce (round 1's protocol)"}},
        "gpu_launches": c1["kernel_launches"] - c0["kernel_launches"],
        "knn_fallback_particles": c1["knn_fallback"] - c0["knn_fallback"],
        # certified reuse of the neighbour lists inside the timed region: evaluations that took the exact kNN from the
        # stored candidates (no sort / reorder / tile search), rebuild evaluations, and the particles the certificate
        # refused (they took the ring-expansion search; counted in knn_fallback_particles)
        "reuse": {"reuse_steps": c1["reuse_steps"] - c0["reuse_steps"], "rebuild_steps": K - (c1["reuse_steps"] - c0["reuse_steps"]),
                  "refused_fraction": (c1["knn_fallback"] - c0["knn_fallback"]) / (n * K)},
        "clocks": clocks,
        "other_build": other,
        # the other BASELINE.json configurations that fit one GPU, device-resident, for the record (not the headline)
        "legs": None if args.no_legs else run_legs(args, local),
    }
    emit(line)


def main():
    # rank 0 prints exactly one JSON line on stdout.  Libraries may write to fd 1 on their own (NCCL_DEBUG=VERSION / INFO
    # print with a bare printf): the process's fd 1 is pointed at stderr for the whole run and the JSON line goes to
    # the saved original stdout.  The caller's NCCL_* environment is left as it is.
    sys.stdout.flush()
    os.environ["SPHB_BENCH_JSON_FD"] = str(os.dup(1))
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c5", choices=["c3", "c4", "c4dam", "c5"])
    ap.add_argument("--precision", type=int, default=64, choices=[64, 32], help="64: reference arithmetic; 32: the fp32 build")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg (profiling runs)")
    ap.add_argument("--no-e2e", action="store_true", help="skip the host-buffer leg (tuning sweeps only)")
    ap.add_argument("--no-other-build", action="store_true", help="skip the extra leg that times the other precision build")
    ap.add_argument("--no-legs", action="store_true", help="skip the legs on the other BASELINE configurations (C3-U/P, C4, speed-test)")
    ap.add_argument("--reuse", action="store_true", help="switch the certified neighbour-list reuse on (SPHB_FLAG_REUSE_LISTS)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
