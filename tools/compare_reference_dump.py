"""Compare the per-tick dumps of tools/dump_reference.jl (ref_tick_NNNN_*.csv) with the CPU oracle, bit-exact.

    python tools/compare_reference_dump.py CASES_DIR            # after the Julia run
    python tools/compare_reference_dump.py CASES_DIR --selftest # write the dumps FROM the oracle first (checks the file formats only)
    python tools/compare_reference_dump.py CASES_DIR --variants # try every alternative reading of the Appendix-A items the
                                                                # survey left at confidence M / L (oracle_set_variants: offset order,
                                                                # random_position draw order, randomwalk ifempty default) and print
                                                                # which combination reproduces the Julia dumps

The day this passes on a box with Julia + Agents.jl 5.14, the oracle stops being "parity unpinned".
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def read_config(path):
    cfg = {}
    for line in open(path):
        if "=" in line:
            k, v = line.split("=")
            cfg[k.strip()] = float(v)
    return cfg


def oracle_engine(d, ifempty=None):
    import abm_b200  # noqa: F401  (registers the package shim)
    from abm_b200 import build, capi
    lib = capi.CLib(build.build_oracle(), "oracle_")
    cfg = read_config(os.path.join(d, "config.txt"))
    ints = ("nx", "ny", "periodic", "walk_type", "eat", "reproduce", "walkmap_regrowth", "regrowth_time", "visual_distance", "counter",
            "astar_metric", "astar_penalty", "seed")
    kw = {k: (int(cfg[k]) if k in ints else cfg[k]) for k in cfg if k not in ("behav", "behav_counter", "ticks")}
    if ifempty is not None:
        kw["randomwalk_ifempty"] = int(ifempty)
    e = lib.create(lib.default_config(**kw))
    grid = lambda fn: np.loadtxt(os.path.join(d, fn), delimiter=",", dtype=np.int64, ndmin=2) if os.path.exists(os.path.join(d, fn)) else None
    wm = grid("walkmap.csv")
    e.set_grid(grid("fully_grown.csv").astype(np.uint8), grid("countdown.csv"), None if wm is None else wm.astype(np.uint8), grid("elevation.csv"))
    a = np.loadtxt(os.path.join(d, "agents.csv"), delimiter=",", skiprows=1, ndmin=2)
    e.set_agents(a[:, 0].astype(np.int64), a[:, 1].astype(np.int64), a[:, 2].astype(np.int64), a[:, 3].astype(np.float64))
    if os.path.exists(os.path.join(d, "lut.csv")):  # Julia's exp(-sqrt(k)/3) table, written by dump_reference.jl for the gregarious rule
        e.set_weight_lut(np.loadtxt(os.path.join(d, "lut.csv"), dtype=np.float64, ndmin=1))
    if int(cfg["walk_type"]) == 3:
        e.set_global_state(0, int(cfg["behav"]), int(cfg["behav_counter"]))
    return e, int(cfg["ticks"])


def write_dump(d, t, snap):
    p = os.path.join(d, f"ref_tick_{t:04d}_")
    with open(p + "agents.csv", "w") as f:
        f.write("id,x,y,energy\n")
        for i, x, y, en in zip(snap["ids"], snap["x"], snap["y"], snap["energy"]):
            f.write(f"{int(i)},{int(x)},{int(y)},{float(en)!r}\n")
    np.savetxt(p + "fully_grown.csv", snap["fully_grown"].astype(np.int64), fmt="%d", delimiter=",")
    np.savetxt(p + "countdown.csv", snap["countdown"].astype(np.int64), fmt="%d", delimiter=",")
    with open(p + "state.txt", "w") as f:
        f.write(f"behav={snap['behav']}\nbehav_counter={snap['behav_counter']}\n")


def compare_case(d, selftest=False, variants=0, ifempty=None):
    e, ticks = oracle_engine(d, ifempty)
    if variants:
        import ctypes as C
        fn = e.lib.dll.oracle_set_variants
        fn.argtypes, fn.restype = [C.c_void_p, C.c_uint32], C.c_int
        assert fn(e.h, variants) == 0
    if selftest:
        w, _ = oracle_engine(d)
        for t in range(1, ticks + 1):
            w.step(1)
            write_dump(d, t, w.snapshot())
    for t in range(1, ticks + 1):
        e.step(1)
        s = e.snapshot()
        p = os.path.join(d, f"ref_tick_{t:04d}_")
        if not os.path.exists(p + "agents.csv"):
            raise SystemExit(f"{d}: no reference dump for tick {t} (run tools/dump_reference.jl first)")
        a = np.loadtxt(p + "agents.csv", delimiter=",", skiprows=1, ndmin=2)
        a = a.reshape(-1, 4)
        ok = (a.shape[0] == s["ids"].size and np.array_equal(a[:, 0].astype(np.int64), s["ids"]) and np.array_equal(a[:, 1].astype(np.int64), s["x"])
              and np.array_equal(a[:, 2].astype(np.int64), s["y"]) and np.array_equal(a[:, 3], s["energy"]))
        fg = np.loadtxt(p + "fully_grown.csv", delimiter=",", dtype=np.int64, ndmin=2)
        cd = np.loadtxt(p + "countdown.csv", delimiter=",", dtype=np.int64, ndmin=2)
        ok = ok and np.array_equal(fg, s["fully_grown"].astype(np.int64)) and np.array_equal(cd, s["countdown"].astype(np.int64))
        st = read_config(p + "state.txt")
        ok = ok and int(st["behav"]) == s["behav"] and int(st["behav_counter"]) == s["behav_counter"]
        if not ok:
            return f"MISMATCH at tick {t}"
    return f"identical for {ticks} ticks"


if __name__ == "__main__":
    base = sys.argv[1]
    selftest = "--selftest" in sys.argv
    if "--variants" in sys.argv:
        names = {1: "offsets second-coordinate-fastest", 2: "random_position y first"}
        cases = [os.path.join(base, n) for n in sorted(os.listdir(base)) if os.path.exists(os.path.join(base, n, "config.txt"))]
        winners = []
        for ifempty in (1, 0):
            for mask in (0, 1, 2, 3):
                res = [compare_case(d, False, mask, ifempty) for d in cases]
                ok = sum(not r.startswith("MISMATCH") for r in res)
                label = ", ".join([f"ifempty={ifempty}"] + [names[b] for b in (1, 2) if mask & b]) or "the engine's reading"
                print(f"variants={mask} {label}: {ok}/{len(cases)} cases identical")
                if ok == len(cases):
                    winners.append((mask, ifempty))
        print("readings that reproduce every dump:", winners or "none — something else differs")
        sys.exit(0 if winners else 1)
    bad = 0
    for name in sorted(os.listdir(base)):
        d = os.path.join(base, name)
        if os.path.exists(os.path.join(d, "config.txt")):
            res = compare_case(d, selftest)
            bad += res.startswith("MISMATCH")
            print(f"{name}: {res}")
    sys.exit(1 if bad else 0)
