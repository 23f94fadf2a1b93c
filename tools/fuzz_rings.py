"""Random strip configurations on rings of 2-3 processes (gloo, emulated engine) against the oracle.

    python tools/fuzz_rings.py START COUNT      # 24 seeds were clean when this was written (~4 s per seed)
"""
import sys, os, time, tempfile, pathlib
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np
import test_strips as ts
from conftest import load_case, assert_same_state, build, capi
from abm_b200 import strips
def main():
    oracle = capi.CLib(build.build_oracle(), "oracle_")
    build.build_emu()
    start, count = int(sys.argv[1]), int(sys.argv[2])
    bad = 0
    for seed in range(start, start + count):
        name = f"fuzz{seed}"
        c = ts._fuzz_case(seed)
        rows = c["ny"] // 32
        world = int(np.random.default_rng(seed + 7).choice([w for w in (2, 3) if rows % w == 0] or [1]))
        if world == 1: continue
        ticks = 6
        with tempfile.TemporaryDirectory() as d:
            per_rank = ts._run_ring(world, name, ticks, pathlib.Path(d))
        cfgkw, w = ts._case(name)
        a = load_case(oracle, cfgkw, w)
        try:
            for t in range(ticks):
                a.step(1)
                assert_same_state(strips.merge_snapshots([per_rank[r][t] for r in range(world)]), a.snapshot(), f"{name} world {world} tick {t}")
            print(seed, world, c["wt"], (c["nx"], c["ny"]), c["n"], "ok", flush=True)
        except AssertionError as exc:
            bad += 1
            print(seed, "DIFF", exc, c, flush=True)
    print("bad", bad)

if __name__ == '__main__':
    main()
