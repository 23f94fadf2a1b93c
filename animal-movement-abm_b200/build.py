"""In-tree builds: libabm.so (nvcc, sm_100a — the product), plus the two test-only CPU libraries.

nvcc cross-compiles without a GPU; the built .so files are git-ignored but travel to the GPU box with the
repository snapshot.
"""
from __future__ import annotations

import os
import shutil
import subprocess

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
CSRC = os.path.join(PKG, "csrc")
LIBABM = os.path.join(PKG, "libabm.so")
LIBORACLE = os.path.join(ROOT, "oracle", "liboracle.so")
LIBEMU = os.path.join(ROOT, "tests", "emu", "libabm_emu.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-fmad=false",  # the gregarious weights and energies must round exactly like the reference's IEEE ops
    "-Xcompiler", "-fPIC,-Wall", "-shared", "-ldl",
]


def _sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC))


def _stale(target: str, deps) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def _run(cmd):
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("build failed: " + " ".join(cmd) + "\n" + r.stdout + r.stderr)
    return r.stdout + r.stderr


def nvcc_path() -> str:
    return shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"


def build_libabm(force: bool = False, verbose: bool = False) -> str:
    deps = _sources() + [os.path.join(ROOT, "include", "abm.h")]
    if force or _stale(LIBABM, deps):
        cmd = [nvcc_path()] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", LIBABM, os.path.join(CSRC, "abm_engine.cu")]
        out = _run(cmd)
        if verbose:
            print(out)
    return LIBABM


def build_oracle(force: bool = False) -> str:
    src = os.path.join(ROOT, "oracle", "abm_oracle.cpp")
    if force or _stale(LIBORACLE, [src, os.path.join(ROOT, "include", "abm.h")]):
        _run(["make", "-C", os.path.join(ROOT, "oracle"), "-B", "liboracle.so"])
    return LIBORACLE


def build_emu(force: bool = False) -> str:
    """Engine sources compiled for the host with -DABM_EMU (tests only; see csrc/abm_platform.h)."""
    deps = _sources() + [os.path.join(ROOT, "include", "abm.h")]
    if force or _stale(LIBEMU, deps):
        os.makedirs(os.path.dirname(LIBEMU), exist_ok=True)
        _run(["g++", "-O2", "-g", "-std=c++17", "-fPIC", "-Wall", "-Wextra", "-Wno-unknown-pragmas", "-ffp-contract=off",
              "-DABM_EMU", "-x", "c++", "-shared", "-o", LIBEMU, os.path.join(CSRC, "abm_engine.cu")])
    return LIBEMU


if __name__ == "__main__":
    print(build_libabm(force=True, verbose=True))
    print(build_oracle(force=True))
    print(build_emu(force=True))
