"""A/B probe of engine builds on one GPU: the same C2 world stepped by each library given on the command line
(3 warm-up + 10 timed ticks, per-kernel CUDA-event times, no L2 flush).  Seconds of GPU time instead of a bench run.

    nvcc ... -DABM_TWO_MAYBE -o animal-movement-abm_b200/libabm_two.so animal-movement-abm_b200/csrc/abm_engine.cu
    python tools/variant_probe.py animal-movement-abm_b200/libabm.so animal-movement-abm_b200/libabm_two.so [--workload C3] [--size 4096]

(Build variants in-tree, `*.so` is git-ignored but travels with gpurun; `abm_b200.build.NVCC_FLAGS` has the flags.)
"""
import argparse
import importlib.util
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import abm_b200  # noqa: E402,F401
from abm_b200 import capi  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("libs", nargs="+")
    ap.add_argument("--workload", default="C2")
    ap.add_argument("--size", type=int, default=4096)
    ap.add_argument("--ticks", type=int, default=10)
    args = ap.parse_args()
    spec = importlib.util.spec_from_file_location("bench", os.path.join(ROOT, "bench.py"))
    b = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(b)
    cfg, w = b.make_world(args.workload, args.size, args.size * args.size // 4)
    for path in args.libs:
        e = b.load(capi.CLib(path, "abm_"), cfg, w, device=0)
        e.set_profiling(True)
        e.step(3)
        tot, ms, rounds, launches = {}, 0.0, 0.0, 0.0
        for _ in range(args.ticks):
            e.step(1)
            tm = e.timing()
            ms += tm["ms_total"] / args.ticks
            rounds += tm["rounds"] / args.ticks
            launches += tm["launches"] / args.ticks
            for k, v in e.kernel_stats().items():
                tot[k] = tot.get(k, 0) + v["ms"] / args.ticks
        print(os.path.basename(path), f"tick {ms:.4f} ms", {k: round(v, 4) for k, v in tot.items()}, f"per-agent rounds {rounds:.1f} launches {launches:.1f}", flush=True)
        e.close()


if __name__ == "__main__":
    main()
