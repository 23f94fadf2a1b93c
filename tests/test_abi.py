"""The C-ABI library loads and exports every symbol include/abm.h declares; argument checking that needs no GPU."""
import ctypes as C
import os
import re

import pytest

from conftest import ROOT, build, capi


def _declared():
    src = open(os.path.join(ROOT, "include", "abm.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(abm_[a-z_0-9]+)\s*\(", src)))


def test_header_declares_the_boundary():
    names = _declared()
    for must in ("abm_create", "abm_destroy", "abm_set_grid", "abm_set_agents", "abm_step", "abm_get_agents", "abm_get_grid",
                 "abm_reduce", "abm_plan_routes", "abm_get_paths", "abm_last_error", "abm_set_weight_lut"):
        assert must in names


def test_libabm_exports_every_declared_symbol():
    path = build.build_libabm()
    dll = C.CDLL(path)
    for name in _declared():
        assert hasattr(dll, name), f"{name} declared in include/abm.h but not exported by libabm.so"


def test_binding_covers_every_declared_symbol():
    assert sorted("abm_" + n for n in capi.SIGNATURES) == _declared()


def test_config_struct_layout_matches_header():
    lib = capi.CLib(build.build_libabm(), "abm_")
    cfg = lib.default_config()
    assert cfg.struct_size == C.sizeof(capi.AbmConfig) == 144
    assert (cfg.nx, cfg.ny, cfg.regrowth_time, cfg.seed) == (80, 80, 30, 321)  # scripts/v0.1.jl:36-46


def test_bad_config_is_rejected_without_a_gpu():
    lib = capi.CLib(build.build_libabm(), "abm_")
    for kw in (dict(nx=0), dict(walk_type=9), dict(regrowth_time=500), dict(walk_type=2, periodic=0), dict(scheduler=2)):
        with pytest.raises(capi.AbmError) as ei:
            lib.create(lib.default_config(**kw))
        assert ei.value.code == capi.ABM_E_BADARG


def test_no_cpu_fallback():
    """Without a CUDA device abm_create fails loudly with ABM_E_CUDA (the product has no CPU path)."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    lib = capi.CLib(build.build_libabm(), "abm_")
    with pytest.raises(capi.AbmError) as ei:
        lib.create(lib.default_config())
    assert ei.value.code == capi.ABM_E_CUDA


def test_package_never_loads_test_libraries():
    """Only tests/bench may touch oracle/ or tests/emu: the package sources must not name them as load targets."""
    pkg = os.path.join(ROOT, "animal-movement-abm_b200")
    for fn in os.listdir(pkg):
        if fn.endswith(".py") and fn != "build.py":
            s = open(os.path.join(pkg, fn)).read()
            for token in ("LIBORACLE", "LIBEMU", "build_oracle", "build_emu", "oracle_", "emu_"):
                code = "\n".join(l for l in s.split("\n") if "``" not in l)  # docstring bullet lines name the prefixes
                assert token not in code, (fn, token)


def _build_c_example(tmp_path):
    import subprocess
    exe = str(tmp_path / "c_host")
    libdir = os.path.dirname(build.build_libabm())
    cmd = ["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I" + os.path.join(ROOT, "include"),
           os.path.join(ROOT, "examples", "c_host.c"), "-L" + libdir, "-labm", "-Wl,-rpath," + libdir, "-o", exe]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return exe


def test_header_is_plain_c_and_the_c_example_links(tmp_path):
    """include/abm.h compiles as strict C99 and examples/c_host.c links against libabm.so; without a GPU the program
    stops at abm_create with ABM_E_CUDA (no CPU path)."""
    import subprocess
    import torch
    exe = _build_c_example(tmp_path)
    if not torch.cuda.is_available():
        r = subprocess.run([exe], capture_output=True, text=True)
        assert r.returncode == 1 and "-> -2" in r.stderr and "no CPU fallback" in r.stderr


@pytest.mark.gpu
def test_c_example_runs_on_the_gpu(tmp_path):
    import subprocess
    exe = _build_c_example(tmp_path)
    r = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    lines = r.stdout.strip().splitlines()
    assert len(lines) == 11 and lines[0].split()[:4] == ["tick", "0", "sheep", "100"] and lines[-1].startswith("tick 100")


def test_binding_signatures_match_the_prototypes():
    """capi.SIGNATURES against the prototypes of include/abm.h: argument count, pointers where the header has pointers, scalar
    widths and float-ness elsewhere, and the return type (a mismatch here corrupts the stack silently at call time)."""
    from test_julia_shim import C_SIZES, _c_prototypes
    protos = _c_prototypes()
    for name, (res, args) in capi.SIGNATURES.items():
        cret, cargs = protos["abm_" + name]
        assert len(args) == len(cargs), f"abm_{name}: binding has {len(args)} arguments, header {len(cargs)}"
        if cret == "void":
            assert res is None, name
        elif "*" in cret:
            assert res in (C.c_char_p, C.c_void_p), name
        else:
            assert C.sizeof(res) == C_SIZES[cret.split()[-1]], name
        for k, (a, ca) in enumerate(zip(args, cargs)):
            is_ptr = a in (C.c_void_p, C.c_char_p) or hasattr(a, "contents") or issubclass(a, C._Pointer)
            if "*" in ca:
                assert is_ptr, f"abm_{name} argument {k}: {ca}"
            else:
                cty = [t for t in ca.replace("const", "").split() if t in C_SIZES][0]
                assert not is_ptr and C.sizeof(a) == C_SIZES[cty], f"abm_{name} argument {k}: {ca}"
                assert (cty in ("double", "float")) == (a in (C.c_double, C.c_float)), f"abm_{name} argument {k}"
