"""Where the cycles of a CTA-structured kernel go: runs the C2 world on a library built with -DABM_PHASE_CLOCKS and prints
the share of every marked phase (thread 0's clock64 between the marks, summed over all CTAs).

    nvcc <abm_b200.build.NVCC_FLAGS> -DABM_PHASE_CLOCKS -o animal-movement-abm_b200/libabm_phase.so animal-movement-abm_b200/csrc/abm_engine.cu
    python tools/phase_probe.py animal-movement-abm_b200/libabm_phase.so [--workload C2] [--size 4096]
"""
import argparse
import ctypes
import importlib.util
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import abm_b200  # noqa: E402,F401
from abm_b200 import capi  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("lib")
    ap.add_argument("--workload", default="C2")
    ap.add_argument("--size", type=int, default=4096)
    ap.add_argument("--ticks", type=int, default=10)
    args = ap.parse_args()
    spec = importlib.util.spec_from_file_location("bench", os.path.join(ROOT, "bench.py"))
    b = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(b)
    cfg, w = b.make_world(args.workload, args.size, args.size * args.size // 4)
    lib = capi.CLib(args.lib, "abm_")
    raw = ctypes.CDLL(args.lib)
    e = b.load(lib, cfg, w, device=0)
    e.step(3)
    buf = (ctypes.c_ulonglong * 32)()
    raw.abm_debug_phase_clocks(buf)
    e.step(args.ticks)
    raw.abm_debug_phase_clocks(buf)
    tot = sum(buf)
    for k, v in enumerate(buf):
        if v:
            print(f"phase {k:2d}: {100.0 * v / tot:5.1f} %   {v / args.ticks:.3e} cycles per tick")
    e.close()


if __name__ == "__main__":
    main()
