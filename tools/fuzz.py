"""Differential fuzzing on random configurations (sizes from 1x1 to 128x64, every walk type, sparse to crowded, rules on/off).

    python tools/fuzz.py engine START COUNT   # the engine's kernels (emulated launcher) against the CPU oracle
    python tools/fuzz.py oracle START COUNT   # the CPU oracle against the independent Python restatement (tests/pyref.py)
    python tools/fuzz.py astar START COUNT    # abm_plan_routes / abm_get_paths (random terrain, goals anywhere, replans,
                                              # 1-64 search slots, route pools that overflow): engine vs oracle

`FUZZ_LIB=abm` (GPU box) fuzzes libabm.so itself instead of the emulated build.  `FUZZ_SCHED=1` steps every case under Schedulers.randomly (default: half of them).  `FORCE_GENERAL=1` sends a third of the tile-sized engine cases down the general path; `ABM_TILE_ROUNDS=1` leaves most
conflicts to the per-agent windows.  Every seed is reproducible: `engine_vs_oracle(seed)` / `oracle_vs_pyref(seed)` return
"ok", "skip ...", "refused ..." (abm_create turned the configuration down), "both raised", or a line starting with DIFF /
ERROR.  tests/test_fuzz.py runs a bounded number of seeds; 2,300 engine, 4,400 oracle and 5,600 A* seeds were clean when this
was written, and at the end of round 2 another 11,000 engine seeds (3,000 of them under Schedulers.randomly with a third on the
forced general path), 4,000 oracle seeds and 4,900 A* seeds (thin grids included), plus the sanitizer and reversed-lane-order
passes of tools/sanitize.sh.
"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402
import pyref  # noqa: E402
from conftest import assert_same_state, build, capi, load_case, make_case  # noqa: E402

_libs = {}


def _lib(which):
    if which not in _libs:
        if which == "oracle":
            _libs[which] = capi.CLib(build.build_oracle(), "oracle_")
        elif os.environ.get("FUZZ_LIB") == "abm":  # on a GPU box: the product library instead of the emulated build
            _libs[which] = capi.CLib(build.LIBABM, "abm_")
        else:
            _libs[which] = capi.CLib(build.build_emu(), "emu_")
    return _libs[which]


def _sched(rng):
    """Schedulers.by_id or Schedulers.randomly (FUZZ_SCHED=0/1 forces one)"""
    v = int(rng.integers(0, 2))
    return int(os.environ["FUZZ_SCHED"]) if "FUZZ_SCHED" in os.environ else v


SIZES = [(64, 64), (96, 64), (64, 96), (128, 64), (5, 7), (33, 31), (40, 30), (1, 9), (2, 2), (17, 64), (64, 33)]


def engine_vs_oracle(seed):
    rng = np.random.default_rng(seed)
    nx, ny = SIZES[rng.integers(len(SIZES))]
    wt = int(rng.integers(0, 5))
    per = int(rng.integers(0, 2))
    if wt == 2: per = 1   # gregarious on non-periodic throws in the reference
    cells = nx * ny
    dens = [0.02, 0.1, 0.3, 0.8, 1.5, 3.0][rng.integers(6)]
    n = max(1, int(cells * dens))
    terr = wt in (3, 4) or (rng.random() < 0.2 and wt != 2)
    kw = dict(reproduction_prob=float(rng.choice([0.0, 0.01, 0.1, 0.5])), regrowth_time=int(rng.integers(1, 8)), delta_energy=float(rng.integers(1, 5)),
              prob_random_walk=float(rng.choice([0.0, 0.3, 0.9, 1.0])), visual_distance=int(rng.integers(1, 8)), eat=int(rng.random() < 0.8), reproduce=int(rng.random() < 0.8),
              movement_cost=float(rng.choice([0.5, 1.0, 2.0])), scheduler=_sched(rng))
    if wt == 4: kw["walkmap_regrowth"] = int(rng.integers(0, 2))
    if wt == 3:
        kw.update(counter=int(rng.integers(2, 9)), astar_metric=int(rng.integers(0, 2)), astar_penalty=int(rng.integers(0, 2)))
        if min(nx, ny) < 5: return "skip"
    try:
        cfgkw, w = make_case(nx, ny, n, wt, per, terrain=terr, seed=int(seed % 100000), **kw)
    except Exception as exc:
        return f"skip ({exc})"
    if os.environ.get("FORCE_GENERAL") and rng.random() < 0.3: cfgkw["reserved0"] = 1
    try:
        a, b = load_case(_lib("oracle"), cfgkw, w), load_case(_lib("emu"), cfgkw, w)
    except Exception as exc:
        return 'refused ' + str(exc)[:50]
    if wt == 3:
        for e in (a, b): e.set_global_state(0, 0, kw["counter"])
    for t in range(int(rng.integers(3, 9))):
        try:
            ra = None; a.step(1)
        except Exception as exc: ra = str(exc)[:60]
        try:
            rb = None; b.step(1)
        except Exception as exc: rb = str(exc)[:60]
        if ra or rb:
            if bool(ra) != bool(rb): return f"ERROR MISMATCH oracle={ra} emu={rb} cfg={cfgkw} n={n}"
            return "both raised"
        try:
            assert_same_state(b.snapshot(), a.snapshot(), f"tick {t}")
        except AssertionError as exc:
            return f"DIFF {exc} cfg={cfgkw} n={n} terr={terr}"
    return "ok"


def oracle_vs_pyref(seed):
    rng = np.random.default_rng(seed)
    nx, ny = int(rng.integers(1, 22)), int(rng.integers(1, 22))
    wt = int(rng.integers(0, 5))
    per = int(rng.integers(0, 2))
    r = int(rng.integers(1, 5))
    if wt == 2:
        per = 1
        nx, ny = max(nx, 2 * r + 3), max(ny, 2 * r + 3)
    if wt == 3: nx, ny = max(nx, 5), max(ny, 5)
    n = max(1, int(nx * ny * [0.05, 0.3, 1.0, 2.5][rng.integers(4)]))
    terr = wt in (3, 4) or (rng.random() < 0.2 and wt != 2)
    kw = dict(reproduction_prob=float(rng.choice([0.0, 0.02, 0.2])), regrowth_time=int(rng.integers(1, 8)), delta_energy=float(rng.integers(1, 5)),
              prob_random_walk=float(rng.choice([0.0, 0.3, 0.9, 1.0])), visual_distance=r, eat=int(rng.random() < 0.8), reproduce=int(rng.random() < 0.8),
              movement_cost=float(rng.choice([0.5, 1.0, 2.0])), scheduler=_sched(rng))
    if wt == 4: kw["walkmap_regrowth"] = int(rng.integers(0, 2))
    if wt == 3: kw.update(counter=int(rng.integers(2, 9)), astar_metric=int(rng.integers(0, 2)), astar_penalty=int(rng.integers(0, 2)))
    try:
        cfgkw, w = make_case(nx, ny, n, wt, per, terrain=terr, seed=int(seed % 100000), **kw)
    except Exception as exc:
        return f"skip ({exc})"
    e = load_case(_lib("oracle"), cfgkw, w)
    m = pyref.Model(nx, ny, per, wt, seed=cfgkw["seed"], eat=bool(kw["eat"]), reproduce=bool(kw["reproduce"]), walkmap_regrowth=bool(kw.get("walkmap_regrowth", 0)),
                    regrowth_time=kw["regrowth_time"], visual_distance=r, counter=kw.get("counter", 50), prob_random_walk=kw["prob_random_walk"], delta_energy=kw["delta_energy"],
                    movement_cost=kw["movement_cost"], reproduction_prob=kw["reproduction_prob"], astar_metric=kw.get("astar_metric", 0), astar_penalty=kw.get("astar_penalty", 0),
                    scheduler=kw["scheduler"])
    m.set_grid(w["fully_grown"], w["countdown"], w["walkmap"], w["elevation"])
    m.set_agents(w["ids"], w["x"], w["y"], w["energy"])
    if wt == 3:
        e.set_global_state(0, 0, kw["counter"]); m.behav, m.behav_counter = 0, kw["counter"]
    for t in range(int(rng.integers(3, 10))):
        ra = rb = None
        try: e.step(1)
        except Exception as exc: ra = str(exc)[:60]
        try: m.step(1)
        except Exception as exc: rb = repr(exc)[:60]
        if ra or rb:
            if bool(ra) != bool(rb): return f"ERROR MISMATCH oracle={ra} pyref={rb} cfg={cfgkw} n={n}"
            return "both raised"
        try:
            assert_same_state(e.snapshot(), m.snapshot(), f"tick {t}")
        except AssertionError as exc:
            return f"DIFF {exc} cfg={cfgkw} n={n} terr={terr}"
    return "ok"


def astar_vs_oracle(seed):
    rng = np.random.default_rng(seed)
    nx, ny = int(rng.integers(3, 70)), int(rng.integers(3, 70))
    per = int(rng.integers(0, 2)); metric = int(rng.integers(0, 2)); pen = int(rng.integers(0, 2))
    if seed % 8 == 0: nx = int(rng.integers(1, 4))  # thin grids: fine when not periodic, refused (dims >= 3) when periodic
    if seed % 16 == 0: ny = int(rng.integers(1, 4))
    n = int(rng.integers(1, 80))
    os.environ["ABM_ASTAR_SLOTS"] = str(int(rng.choice([1, 3, 64])))
    os.environ["ABM_ROUTE_POOL_GUESS"] = str(int(rng.choice([5, 200, 100000])))
    try:
        cfgkw, w = make_case(nx, ny, n, 3, per, terrain=True, seed=int(seed % 99991), astar_metric=metric, astar_penalty=pen, counter=9)
    except Exception as exc:
        return "skip"
    ok = np.argwhere(np.ones_like(w["walkmap"]) == 1)   # goals anywhere, walkable or not
    goals = ok[rng.integers(0, len(ok), n)]
    res = []
    for lib in (_lib("oracle"), _lib("emu")):
        try:
            e = load_case(lib, cfgkw, w)
        except Exception as exc:
            if lib is _lib("emu") and per and min(nx, ny) < 3 and "dims >= 3" in str(exc): return "refused thin periodic grid"
            raise
        e.plan_routes(w["ids"], goals[:, 1] + 1, goals[:, 0] + 1)
        res.append(e.get_paths(w["ids"]))
        # replan half of them to new goals: routes are overwritten / kept when unreachable
        g2 = ok[np.random.default_rng(seed + 1).integers(0, len(ok), n)]
        half = w["ids"][: max(1, n // 2)]
        e.plan_routes(half, g2[: half.size, 1] + 1, g2[: half.size, 0] + 1)
        res.append(e.get_paths(w["ids"]))
    for k in range(2):
        for u, v in zip(res[k], res[k + 2]):
            if not np.array_equal(u, v): return f"DIFF seed {seed} stage {k} cfg={cfgkw} n={n}"
    return "ok"


if __name__ == "__main__":
    fn = {"engine": engine_vs_oracle, "oracle": oracle_vs_pyref, "astar": astar_vs_oracle}[sys.argv[1]]
    t0, stats = time.time(), {}
    start, count = int(sys.argv[2]), int(sys.argv[3])
    for s in range(start, start + count):
        r = fn(s)
        k = r.split()[0]
        stats[k] = stats.get(k, 0) + 1
        if k in ("DIFF", "ERROR"):
            print(s, r)
    print(stats, round(time.time() - t0, 1), "s")
    sys.exit(1 if stats.get("DIFF") or stats.get("ERROR") else 0)
