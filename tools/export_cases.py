"""Write the initial worlds of tests/cases.py as plain CSV, one directory per case, for tools/dump_reference.jl.

    python tools/export_cases.py OUT_DIR [case ...]

Per case: config.txt (key=value), agents.csv (id,x,y,energy; 1-based positions, id order), fully_grown.csv /
countdown.csv [/ walkmap.csv / elevation.csv]: ny rows of nx comma-separated integers (row = y, column = x).
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

DEFAULTS = dict(eat=1, reproduce=1, walkmap_regrowth=0, regrowth_time=30, visual_distance=5, counter=50, prob_random_walk=0.9,
                delta_energy=4.0, movement_cost=1.0, reproduction_prob=0.001, astar_metric=0, astar_penalty=0)


def export_case(name, out_dir):
    from cases import CASES, build_case
    cfgkw, w, state, ticks = build_case(name)
    d = os.path.join(out_dir, name)
    os.makedirs(d, exist_ok=True)
    cfg = dict(DEFAULTS)
    cfg.update(cfgkw)
    cfg["behav"], cfg["behav_counter"] = (state[1], state[2]) if state else (0, 0)
    cfg["ticks"] = max(ticks)
    with open(os.path.join(d, "config.txt"), "w") as f:
        for k in sorted(cfg):
            f.write(f"{k}={cfg[k]!r}\n" if isinstance(cfg[k], float) else f"{k}={int(cfg[k])}\n")
    with open(os.path.join(d, "agents.csv"), "w") as f:
        f.write("id,x,y,energy\n")
        for i, x, y, e in zip(w["ids"], w["x"], w["y"], w["energy"]):
            f.write(f"{int(i)},{int(x)},{int(y)},{float(e)!r}\n")
    for key, fn in (("fully_grown", "fully_grown.csv"), ("countdown", "countdown.csv"), ("walkmap", "walkmap.csv"), ("elevation", "elevation.csv")):
        if w.get(key) is not None:
            np.savetxt(os.path.join(d, fn), np.asarray(w[key]).astype(np.int64), fmt="%d", delimiter=",")
    return d


if __name__ == "__main__":
    from cases import CASES
    out = sys.argv[1]
    names = sys.argv[2:] or sorted(CASES)
    for n in names:
        print(export_case(n, out))
