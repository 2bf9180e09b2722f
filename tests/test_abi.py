"""CPU: the C-ABI shared library loads and exports every symbol include/sphb.h declares; without a GPU the
product path fails loudly instead of computing anything on the CPU."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "sphb.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(sphb_[a-z_0-9]+)\s*\(", src)))


def test_library_builds_and_exports_every_declared_symbol():
    from sphugo_b200 import build
    so = build.build()
    out = subprocess.run(["nm", "-D", "--defined-only", so], capture_output=True, text=True, check=True).stdout
    exported = set(re.findall(r" T (sphb_[a-z_0-9]+)", out))
    declared = _declared()
    assert len(declared) >= 28
    missing = [f for f in declared if f not in exported]
    assert not missing, missing
    from sphugo_b200 import _lib
    assert sorted(_lib.EXPORTS) == declared  # the ctypes binding covers the whole ABI
    lib = _lib.lib()
    for f in declared:
        getattr(lib, f)


def test_no_torch_or_cxx_types_in_the_abi():
    src = open(os.path.join(ROOT, "include", "sphb.h")).read()
    assert "torch" not in src and "std::" not in src and "#include <stdint.h>" in src
    assert 'extern "C"' in src


def test_no_cpu_fallback():
    """without a CUDA device sphb_create must fail with SPHB_E_CUDA; the package never imports oracle/"""
    import torch
    from sphugo_b200 import _lib
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(_lib.SphbError) as ei:
        _lib.Handle(_lib.make_params(), np.random.rand(100, 2))
    assert ei.value.code == _lib.E_CUDA
    assert "no CPU fallback" in str(ei.value)
    for dp, _, files in os.walk(os.path.join(ROOT, "sphugo_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".inc")):
                txt = open(os.path.join(dp, f)).read()
                assert "import oracle" not in txt and "from oracle" not in txt and "liborc" not in txt, f


def test_params_struct_layout_matches_header():
    from sphugo_b200 import _lib
    from oracle import oracle as orc
    assert C.sizeof(_lib.Params) == 13 * 8 + 4 * 4 == C.sizeof(orc.OrcParams)
    assert C.sizeof(_lib.Slab) == 4 * 8 + 2 * 4


def test_sm100a_sass_only():
    from sphugo_b200 import build
    out = subprocess.run(["cuobjdump", "-lelf", build.build()], capture_output=True, text=True).stdout
    assert "sm_100a" in out and "sm_90" not in out


def test_fp64_force_kernel_stages_with_the_bulk_copy_engine():
    """every fp64 instantiation of k_force_st stages its neighbour records with cp.async.bulk (SASS UBLKCP) behind
    an mbarrier (SYNCS), keeps its 96-register budget and does not spill; the fp32 twin converts on the way in and has none."""
    from sphugo_b200 import build
    lib = build.build()
    res = subprocess.run(["cuobjdump", "-res-usage", lib], capture_output=True, text=True).stdout
    usage = dict(re.findall(r"Function (\S+?):\s*\n?\s*(REG:\d+ STACK:\d+)", res))
    f64 = [k for k in usage if "k_force_stIL" in k]
    assert len(f64) == 8, sorted(usage)[:5]
    for k in f64:
        assert usage[k] == "REG:96 STACK:0", (k, usage[k])
    sass = subprocess.run(["cuobjdump", "-sass", "-fun", f64[0], lib], capture_output=True, text=True).stdout
    assert sass.count("UBLKCP") == 3 and "SYNCS.ARRIVE.TRANS64" in sass and "SYNCS.PHASECHK.TRANS64.TRYWAIT" in sass
    f32 = [k for k in usage if "k_force_st32IL" in k]
    sass32 = subprocess.run(["cuobjdump", "-sass", "-fun", f32[0], lib], capture_output=True, text=True).stdout
    assert len(sass32) > 1000 and "UBLKCP" not in sass32


def test_every_entry_point_is_mapped_in_integration_md():
    """INTEGRATION.md names, for every declared entry point, the reference interface it replaces"""
    text = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    missing = [f for f in _declared() if f not in text]
    assert not missing, missing


def test_header_cites_the_reference_for_the_step_path():
    src = open(os.path.join(ROOT, "include", "sphb.h")).read()
    for cite in ("sph.go", "nearest-neighbour.go", "core.go", "config-parser.go", "animator.go"):
        assert cite in src, cite


def test_header_is_plain_c_and_a_c_client_links():
    """include/sphb.h compiles as C99 and a C program (tests/c/abi_smoke.c, what cgo would build) links against
    libsphb.so; on a box without a GPU it must see SPHB_E_CUDA from sphb_create, with a GPU one step must run"""
    from sphugo_b200 import build
    so = build.build()
    inc = os.path.join(ROOT, "include")
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-fsyntax-only", "-x", "c", os.path.join(inc, "sphb.h")], check=True)
    exe = os.path.join(ROOT, "tests", "c", "abi_smoke")
    libdir = os.path.dirname(so)
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-I", inc, os.path.join(ROOT, "tests", "c", "abi_smoke.c"),
                    "-L", libdir, "-lsphb", "-Wl,-rpath," + libdir, "-o", exe], check=True)
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr


def test_cxx_mirror_of_the_sim_api_compiles_and_behaves():
    """include/sphb_sim.hpp (C++ mirror of the Go `sim` step API) builds with g++ against libsphb.so; without a GPU
    its constructor throws Panic(SPHB_E_CUDA), with one it steps the speed-test shape (tests/c/sim_smoke.cpp)"""
    from sphugo_b200 import build
    so = build.build()
    inc, libdir = os.path.join(ROOT, "include"), os.path.dirname(so)
    exe = os.path.join(ROOT, "tests", "c", "sim_smoke")
    subprocess.run(["g++", "-std=c++17", "-Wall", "-Werror", "-O1", "-I", inc, os.path.join(ROOT, "tests", "c", "sim_smoke.cpp"),
                    "-L", libdir, "-lsphb", "-Wl,-rpath," + libdir, "-o", exe], check=True)
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr


def test_cxx_gorand_known_answers_and_python_twin():
    """include/sphb_gorand.hpp: Go's math/rand in C++ for the compiled host side; same KATs as tests/test_gorand.py, and the
    reference's seed 12345678 gives the same draws as sphugo_b200/gorand.py"""
    from sphugo_b200 import gorand
    exe = os.path.join(ROOT, "tests", "c", "gorand_kat")
    subprocess.run(["g++", "-std=c++17", "-Wall", "-Werror", "-O2", "-I", os.path.join(ROOT, "include"),
                    os.path.join(ROOT, "tests", "c", "gorand_kat.cpp"), "-o", exe], check=True)
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 0, r.returncode
    f0, f1, z = r.stdout.split()
    g = gorand.Rand(12345678)
    assert (float(f0), float(f1), int(z)) == (g.Float64(), g.Float64(), g.Int())


def test_cxx_examples_build_and_panic_without_a_device():
    """examples/*.cpp: the reference's three BASELINE example mains over the C++ host side build warning-free; without a GPU
    they end like a Go panic (message on stderr, exit status 2) - no CPU fallback"""
    import torch
    from sphugo_b200 import build
    build.build()
    subprocess.run(["make", "-C", os.path.join(ROOT, "examples"), "-s"], check=True)
    if torch.cuda.is_available():
        return
    for exe in ("speed_test", "density", "sph_simulation"):
        r = subprocess.run([os.path.join(ROOT, "examples", exe)], capture_output=True, text=True)
        assert r.returncode == 2 and r.stderr.startswith("panic: no CUDA device"), (exe, r.returncode, r.stderr)
