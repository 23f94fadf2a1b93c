"""Trajectory parity on the named cases of tests/cases.py.

  oracle  vs committed golden fixtures     (not gpu)  — the oracle has not silently changed
  emu     vs oracle, tick by tick           (not gpu)  — the engine's kernel bodies + host orchestration, on the host
  libabm  vs golden and vs oracle           (gpu)      — the product, through the C ABI

Bar: bit-exact ids / cells / grass / countdown / global state; energy rtol 1e-6 (observed exact).
"""
import json
import os

import numpy as np
import pytest

from cases import CASES, build_case
from conftest import ROOT, assert_same_state, capi, load_case, make_case

GOLDEN = os.path.join(ROOT, "tests", "golden")


def _golden(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    meta = json.loads(bytes(z["meta"]).decode())
    snaps = {}
    for t in meta["ticks"]:
        s = {k: z[f"t{t}_{k}"] for k in ("ids", "x", "y", "energy", "fully_grown", "countdown")}
        s["x"], s["y"], s["countdown"] = s["x"].astype(np.int64), s["y"].astype(np.int64), s["countdown"].astype(np.int64)
        for k in ("tick", "behav", "behav_counter", "next_id"):
            s[k] = int(z[f"t{t}_{k}"])
        snaps[t] = s
    return snaps


# walk types the CUDA engine does not implement yet (ALTERNATED_WALK + A*): the engine refuses them loudly
ENGINE_PENDING = set()


def _run_against_golden(lib, name):
    if lib.prefix != "oracle_" and name in ENGINE_PENDING:
        pytest.xfail("ALTERNATED_WALK not implemented in the engine yet")
    cfgkw, w, state, ticks = build_case(name)
    e = load_case(lib, cfgkw, w)
    if state:
        e.set_global_state(*state)
    gold = _golden(name)
    done = 0
    for t in ticks:
        e.step(t - done)
        done = t
        assert_same_state(e.snapshot(), gold[t], f"{name} @tick {t}")
    e.close()


@pytest.mark.parametrize("name", sorted(CASES))
def test_oracle_matches_golden(oracle, name):
    _run_against_golden(oracle, name)


EMU_CASES = [n for n in sorted(CASES) if n != "C1_v0.1_100x100"]  # C1's 1000 ticks are covered on the GPU


@pytest.mark.parametrize("name", EMU_CASES)
def test_emulated_engine_matches_golden(emu, name):
    _run_against_golden(emu, name)


@pytest.mark.parametrize("name", ["dense_rw", "dense_gregarious_r3", "dense_attractor", "dense_walkmap", "dense_alternated"])
@pytest.mark.parametrize("shuffle", [1, 2, 3])
def test_emulated_engine_is_schedule_independent(emu, oracle, name, shuffle):
    """The emulation visits thread indices in a shuffled order; any order must give the sequential result."""
    if name in ENGINE_PENDING:
        pytest.xfail("ALTERNATED_WALK not implemented in the engine yet")
    emu.dll.emu_set_shuffle_seed(1000 * shuffle + 7)
    cfgkw, w, state, ticks = build_case(name)
    a, b = load_case(oracle, cfgkw, w), load_case(emu, cfgkw, w)
    if state:
        a.set_global_state(*state)
        b.set_global_state(*state)
    for t in range(12):
        a.step(1)
        b.step(1)
        assert_same_state(b.snapshot(), a.snapshot(), f"{name} shuffle {shuffle} tick {t}")


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(CASES))
def test_gpu_matches_golden(libabm, name):
    _run_against_golden(libabm, name)


@pytest.mark.gpu
@pytest.mark.parametrize("wt,per,terr,n,nx,ny,extra", [
    (0, 1, 0, 20000, 256, 192, {}),
    (0, 0, 0, 20000, 250, 190, {}),
    (1, 0, 0, 20000, 256, 192, dict(prob_random_walk=0.6)),
    (2, 1, 0, 16000, 256, 256, dict(visual_distance=3)),
    (2, 1, 0, 9000, 200, 130, dict(visual_distance=5)),
    (4, 1, 1, 20000, 256, 256, dict(walkmap_regrowth=1)),
    (3, 1, 1, 6000, 192, 160, dict(counter=50, astar_metric=1, astar_penalty=1)),
    # Schedulers.randomly on the general, tile and stream paths
    (0, 0, 0, 20000, 250, 190, dict(scheduler=1)),
    (2, 1, 0, 16000, 256, 256, dict(visual_distance=3, scheduler=1)),
    (4, 1, 1, 20000, 256, 256, dict(walkmap_regrowth=1, scheduler=1)),
    (3, 1, 1, 6000, 192, 160, dict(counter=50, astar_metric=1, astar_penalty=1, scheduler=1)),
])
def test_gpu_matches_oracle_tick_by_tick(libabm, oracle, wt, per, terr, n, nx, ny, extra):
    """Mid-size seeded worlds (the oracle needs seconds): every tick compared."""
    from conftest import make_case
    kw = dict(reproduction_prob=0.02, regrowth_time=6, delta_energy=3.0)
    kw.update(extra)
    cfgkw, w = make_case(nx, ny, n, wt, per, terrain=bool(terr), seed=99, **kw)
    a, b = load_case(oracle, cfgkw, w), load_case(libabm, cfgkw, w)
    if wt == 3:
        a.set_global_state(0, 0, kw["counter"])
        b.set_global_state(0, 0, kw["counter"])
    for t in range(20):
        a.step(1)
        b.step(1)
        assert_same_state(b.snapshot(), a.snapshot(), f"wt {wt} tick {t}")
    a.step(30)
    b.step(30)
    assert_same_state(b.snapshot(), a.snapshot(), f"wt {wt} tick 50")


@pytest.mark.gpu
def test_gpu_weight_lut_is_honoured(libabm, oracle):
    """abm_set_weight_lut: the host's exp table decides the wsample index; a distorted table must change both sides alike."""
    from conftest import make_case
    from abm_b200 import worlds
    cfgkw, w = make_case(96, 96, 4000, 2, 1, seed=4, visual_distance=3, prob_random_walk=0.2)
    lut = worlds.weight_lut(3)
    lut[1:] *= 0.5
    a, b = load_case(oracle, cfgkw, w, lut), load_case(libabm, cfgkw, w, lut)
    a.step(10)
    b.step(10)
    assert_same_state(b.snapshot(), a.snapshot(), "distorted LUT")


def _plan_case(lib, metric, penalty, periodic, nx=40, ny=31, n=60, seed=12):
    from conftest import make_case
    cfgkw, w = make_case(nx, ny, n, 3, periodic, terrain=True, seed=seed, astar_metric=metric, astar_penalty=penalty, counter=9)
    e = load_case(lib, cfgkw, w)
    ok = np.argwhere(w["walkmap"] == 1)
    rng = np.random.default_rng(seed)
    goals = ok[rng.integers(0, len(ok), n)]
    goals[0] = (w["y"][0] - 1, w["x"][0] - 1)  # start == goal: an empty route
    e.plan_routes(w["ids"], goals[:, 1] + 1, goals[:, 0] + 1)
    return e, w


def _same_paths(a, b, ids):
    oa, xa, ya, ca = a.get_paths(ids)
    ob, xb, yb, cb = b.get_paths(ids)
    assert np.array_equal(oa, ob) and np.array_equal(xa, xb) and np.array_equal(ya, yb) and np.array_equal(ca, cb)
    return oa


@pytest.mark.parametrize("metric,penalty,periodic", [(1, 1, 1), (0, 0, 1), (1, 1, 0), (0, 1, 0)])
def test_emulated_astar_paths_match_oracle(emu, oracle, metric, penalty, periodic):
    """abm_plan_routes / abm_get_paths: every path cell and the path cost, bit-exact; then the routes are followed."""
    a, w = _plan_case(oracle, metric, penalty, periodic)
    b, _ = _plan_case(emu, metric, penalty, periodic)
    off = _same_paths(a, b, w["ids"])
    assert off[-1] > 50
    for lib_e in (a, b):
        lib_e.set_global_state(0, 2, 5)  # behaviour 2 (follow the route) for the next 5 agent-steps, then a switch
    for t in range(6):
        a.step(1)
        b.step(1)
        assert_same_state(b.snapshot(), a.snapshot(), f"route following tick {t}")
        ids = a.get_agents()[0]
        _same_paths(a, b, ids)


def test_emulated_astar_slot_reuse_and_pool_overflow(emu, oracle, monkeypatch):
    """More than 255 searches through one slot (the emulated launcher lets the first CTA drain the queue: the 8-bit
    generation stamp wraps and the slot is wiped) and a route pool that starts far too small (requests re-run)."""
    monkeypatch.setenv("ABM_ROUTE_POOL_GUESS", "40")
    monkeypatch.setenv("ABM_ASTAR_SLOTS", "3")  # fewer slots than requests: the far-goals-first ordering runs too
    a, w = _plan_case(oracle, 1, 1, 1, nx=48, ny=40, n=300, seed=5)
    b, _ = _plan_case(emu, 1, 1, 1, nx=48, ny=40, n=300, seed=5)
    off = _same_paths(a, b, w["ids"])
    assert off[-1] > 2000
    # a second batch over the same agents rebinds every route
    rng = np.random.default_rng(9)
    ok = np.argwhere(w["walkmap"] == 1)
    goals = ok[rng.integers(0, len(ok), 300)]
    for e in (a, b):
        e.plan_routes(w["ids"], goals[:, 1] + 1, goals[:, 0] + 1)
    _same_paths(a, b, w["ids"])


def test_emulated_astar_open_list_overflow_is_rerun(emu, oracle, monkeypatch):
    """Open lists start small (2^18 entries per slot; here 4) and a search that overflows runs again after the heaps grew
    fourfold (4 -> 16 -> 64 -> ...): same paths and costs as the oracle's unbounded open list."""
    monkeypatch.setenv("ABM_ASTAR_HEAP0", "4")
    monkeypatch.setenv("ABM_ASTAR_SLOTS", "5")
    a, w = _plan_case(oracle, 1, 1, 1, nx=48, ny=40, n=120, seed=6)
    b, _ = _plan_case(emu, 1, 1, 1, nx=48, ny=40, n=120, seed=6)
    off = _same_paths(a, b, w["ids"])
    assert off[-1] > 800
    monkeypatch.delenv("ABM_ASTAR_HEAP0")  # the grown capacity stays with the handle
    rng = np.random.default_rng(10)
    ok = np.argwhere(w["walkmap"] == 1)
    goals = ok[rng.integers(0, len(ok), 120)]
    for e in (a, b):
        e.plan_routes(w["ids"], goals[:, 1] + 1, goals[:, 0] + 1)
    _same_paths(a, b, w["ids"])


@pytest.mark.gpu
def test_gpu_astar_open_list_overflow_is_rerun(libabm, oracle, monkeypatch):
    monkeypatch.setenv("ABM_ASTAR_HEAP0", "16")
    a, w = _plan_case(oracle, 1, 1, 1, nx=160, ny=120, n=500, seed=7)
    b, _ = _plan_case(libabm, 1, 1, 1, nx=160, ny=120, n=500, seed=7)
    _same_paths(a, b, w["ids"])


def test_emulated_astar_refuses_periodic_grids_thinner_than_three(emu):
    """Two of the 8 neighbour offsets wrap onto the same cell there; the batched search relaxes its 8 neighbours side by side
    and says so instead of guessing (the refusal is host code, the same in libabm.so)."""
    with pytest.raises(Exception, match="dims >= 3"):
        emu.create(emu.default_config(nx=7, ny=2, periodic=1, walk_type=3, counter=9))
    e = emu.create(emu.default_config(nx=7, ny=2, periodic=1, walk_type=0))
    e.set_grid(np.ones((2, 7), np.uint8), np.full((2, 7), 30, np.int64), np.ones((2, 7), np.uint8), np.ones((2, 7), np.int64))
    e.set_agents(np.array([1]), np.array([1]), np.array([1]), np.array([5.0]))
    with pytest.raises(Exception, match="dims >= 3"):
        e.plan_routes(np.array([1]), np.array([5]), np.array([2]))
    # not periodic: nothing wraps, thin grids are fine
    e = emu.create(emu.default_config(nx=7, ny=2, periodic=0, walk_type=0))
    e.set_grid(np.ones((2, 7), np.uint8), np.full((2, 7), 30, np.int64), np.ones((2, 7), np.uint8), np.ones((2, 7), np.int64))
    e.set_agents(np.array([1]), np.array([1]), np.array([1]), np.array([5.0]))
    e.plan_routes(np.array([1]), np.array([5]), np.array([2]))
    off, px, py, cost = e.get_paths(np.array([1]))
    assert off[-1] == 4 and px[-1] == 5 and py[-1] == 2


def _island_case(lib, periodic):
    """A wall of water splits the map; agents on both sides, on the wall (not walkable) and goals everywhere."""
    nx, ny = 40, 32
    wm = np.ones((ny, nx), np.uint8)
    wm[:, 18:21] = 0       # the wall
    wm[10, 18:21] = 1 if periodic else 0  # periodic case: a bridge through the wall
    wm[:, 0] = 0           # second wall on the seam: without the bridge the halves are disconnected even when periodic
    wm[20:24, 30:34] = 0
    wm[21:23, 31:33] = 1   # a lake with an island in it (diagonal moves cannot cross a 1-cell ring either)
    el = (np.arange(nx)[None, :] * 3 + np.arange(ny)[:, None] * 5) % 17 + 1
    fg = np.ones((ny, nx), np.uint8)
    cd = np.full((ny, nx), 30, np.int64)
    starts = [(3, 3), (25, 5), (19, 4), (32, 22), (5, 28), (38, 30), (2, 16), (22, 11)]   # (x, y) 1-based; (19,4) stands on the wall
    goals = [(15, 20), (4, 25), (6, 6), (8, 8), (32, 22), (33, 23), (19, 20), (35, 11)]     # (19,20) is water; (32,22)/(33,23) the island
    n = len(starts)
    e = lib.create(lib.default_config(nx=nx, ny=ny, periodic=periodic, walk_type=3, astar_metric=1, astar_penalty=1, counter=9))
    e.set_grid(fg, cd, wm, el.astype(np.int64))
    ids = np.arange(1, n + 1)
    e.set_agents(ids, np.array([s[0] for s in starts]), np.array([s[1] for s in starts]), np.full(n, 5.0))
    e.plan_routes(ids, np.array([g[0] for g in goals]), np.array([g[1] for g in goals]))
    return e, ids


@pytest.mark.parametrize("periodic", [0, 1])
def test_emulated_astar_unreachable_goals(emu, oracle, periodic):
    """Goals in another component of the walkmap, on water, and a start that stands on water: the engine answers the
    first two from its component labels without searching; results must equal the oracle's exhaustive search."""
    a, ids = _island_case(oracle, periodic)
    b, _ = _island_case(emu, periodic)
    off = _same_paths(a, b, ids)
    _, _, _, cost = b.get_paths(ids)
    assert (cost < 0).sum() >= 3 and (cost >= 0).sum() >= 2


@pytest.mark.gpu
@pytest.mark.parametrize("periodic", [0, 1])
def test_gpu_astar_unreachable_goals(libabm, oracle, periodic):
    a, ids = _island_case(oracle, periodic)
    b, _ = _island_case(libabm, periodic)
    _same_paths(a, b, ids)


@pytest.mark.gpu
def test_gpu_astar_pool_overflow_and_many_requests(libabm, oracle, monkeypatch):
    """The device-side request queue with more requests than search slots, and the pool-overflow retry, on the GPU."""
    monkeypatch.setenv("ABM_ROUTE_POOL_GUESS", "1000")
    monkeypatch.setenv("ABM_ASTAR_GB", "1")
    a, w = _plan_case(oracle, 1, 1, 1, nx=256, ny=192, n=6000, seed=8)
    b, _ = _plan_case(libabm, 1, 1, 1, nx=256, ny=192, n=6000, seed=8)
    _same_paths(a, b, w["ids"])


@pytest.mark.gpu
@pytest.mark.parametrize("metric,penalty,periodic", [(1, 1, 1), (0, 0, 1), (1, 1, 0)])
def test_gpu_astar_paths_match_oracle(libabm, oracle, metric, penalty, periodic):
    a, w = _plan_case(oracle, metric, penalty, periodic, nx=160, ny=120, n=400)
    b, _ = _plan_case(libabm, metric, penalty, periodic, nx=160, ny=120, n=400)
    _same_paths(a, b, w["ids"])
    for lib_e in (a, b):
        lib_e.set_global_state(0, 2, 7)
    for t in range(8):
        a.step(1)
        b.step(1)
        assert_same_state(b.snapshot(), a.snapshot(), f"route following tick {t}")
    _same_paths(a, b, a.get_agents()[0])


@pytest.mark.parametrize("name", ["tile_rw_96x64", "tile_gregarious_r3_96x96", "tile_gregarious_r7_p0_64x64", "tile_attractor_128x64", "tile_crowded_64x64"])
def test_emulated_tile_path_with_per_agent_windows(emu, oracle, name, monkeypatch):
    """One in-window round only: most conflicts are then left to the per-agent windows (resolve_agents)."""
    monkeypatch.setenv("ABM_TILE_ROUNDS", "1")
    cfgkw, w, state, ticks = build_case(name)
    a, b = load_case(oracle, cfgkw, w), load_case(emu, cfgkw, w)
    used = 0
    for t in range(8):
        a.step(1)
        b.step(1)
        used += b.kernel_stats().get("resolve_agents", dict(launches=0))["launches"]
        assert_same_state(b.snapshot(), a.snapshot(), f"{name} tick {t}")
    assert used > 0


@pytest.mark.parametrize("name", ["tile_rw_96x64", "tile_gregarious_r3_96x96", "tile_walkmap_128x96"])
def test_emulated_general_path_on_tile_sized_grids(emu, oracle, name):
    """abm_config.reserved0 = 1 forces the general path: both engine paths must agree with the oracle."""
    cfgkw, w, state, ticks = build_case(name)
    a, b = load_case(oracle, cfgkw, w), load_case(emu, dict(cfgkw, reserved0=1), w)
    for t in range(6):
        a.step(1)
        b.step(1)
        assert "resolve_tiles" not in {k for k, v in b.kernel_stats().items() if v["launches"]}
        assert_same_state(b.snapshot(), a.snapshot(), f"{name} tick {t}")


@pytest.mark.gpu
@pytest.mark.parametrize("wt,per,terr,n,nx,ny,extra", [
    (0, 1, 0, 60000, 512, 256, {}),
    (0, 0, 0, 60000, 256, 512, {}),
    (1, 0, 0, 60000, 512, 256, dict(prob_random_walk=0.6)),
    (2, 1, 0, 40000, 384, 384, dict(visual_distance=3)),
    (2, 1, 0, 20000, 256, 256, dict(visual_distance=7, prob_random_walk=0.5)),
    (4, 1, 1, 60000, 512, 512, dict(walkmap_regrowth=1)),
])
def test_gpu_tile_path_matches_oracle(libabm, oracle, wt, per, terr, n, nx, ny, extra):
    from conftest import make_case
    kw = dict(reproduction_prob=0.02, regrowth_time=6, delta_energy=3.0)
    kw.update(extra)
    cfgkw, w = make_case(nx, ny, n, wt, per, terrain=bool(terr), seed=77, **kw)
    a, b = load_case(oracle, cfgkw, w), load_case(libabm, cfgkw, w)
    for t in range(12):
        a.step(1)
        b.step(1)
        assert ("stream_tick" if wt == 4 else "commit_tiles") in {k for k, v in b.kernel_stats().items() if v["launches"]}
        assert_same_state(b.snapshot(), a.snapshot(), f"wt {wt} tick {t}")
    a.step(20)
    b.step(20)
    assert_same_state(b.snapshot(), a.snapshot(), f"wt {wt} tick 32")


def _full_size_world(size, wt, seed=5):
    from abm_b200 import worlds
    wm = el = None
    if wt == 4:
        el = worlds.make_terrain(size, size, seed=seed + 1)
        wm = worlds.walkmap_from_elevation(el)
    w = worlds.make_world(size, size, size * size // 4, seed=seed, walkmap=wm, periodic=True)
    w["walkmap"], w["elevation"] = wm, el
    return w


@pytest.mark.gpu
@pytest.mark.parametrize("wt,size,extra", [
    (2, 4096, dict(visual_distance=3, reproduction_prob=0.001)),          # BASELINE.json configs[1] (C2) at full size
    (4, 4096, dict(walkmap_regrowth=1, reproduction_prob=0.004)),         # C3 rules at a quarter of its area
    (0, 2048, dict(reproduction_prob=0.01)),
])
def test_gpu_full_size_two_engine_paths_and_invariants(libabm, wt, size, extra):
    """At sizes the CPU oracle cannot reach in seconds: the two independently written GPU paths (tile path, general
    path forced with reserved0 = 1) must agree bit for bit, and the trajectory must satisfy the size-independent
    properties of the rules: ids strictly increasing and unique, every survivor moved at most one cell (Chebyshev,
    periodic), population = survivors + offspring, offspring ids = the next free ids, at most one meal per grown cell,
    energy bookkeeping (every meal adds exactly delta_energy), grass countdown within range."""
    w = _full_size_world(size, wt)
    cfg = dict(nx=size, ny=size, periodic=1, walk_type=wt, seed=77, **extra)
    a, b = load_case(libabm, cfg, w), load_case(libabm, dict(cfg, reserved0=1), w)
    prev = a.snapshot()
    for t in range(3):
        a.step(1)
        b.step(1)
        assert ("stream_tick" if wt == 4 else "commit_tiles") in {k for k, v in a.kernel_stats().items() if v["launches"]}
        assert not {"commit_tiles", "stream_tick"} & {k for k, v in b.kernel_stats().items() if v["launches"]}
        sa, sb = a.snapshot(), b.snapshot()
        assert_same_state(sa, sb, f"tile vs general path, walk type {wt}, tick {t}")
        ids, pid = sa["ids"], prev["ids"]
        assert np.all(np.diff(ids) > 0)
        old = np.isin(ids, pid, assume_unique=True)
        born = ids[~old]
        assert born.size == sa["next_id"] - prev["next_id"] and (born.size == 0 or (born.min() == prev["next_id"] and born.max() == sa["next_id"] - 1))
        was = np.searchsorted(pid, ids[old])
        dx = np.abs(sa["x"][old] - prev["x"][was]); dy = np.abs(sa["y"][old] - prev["y"][was])
        dx = np.minimum(dx, size - dx); dy = np.minimum(dy, size - dy)
        assert max(dx.max(), dy.max()) <= 1
        eaten = int(((prev["fully_grown"] == 1) & (sa["fully_grown"] == 0)).sum())
        assert 0 < eaten <= pid.size
        assert sa["countdown"].min() >= 0 and sa["countdown"].max() <= 30
        assert np.all(sa["energy"][old] >= 0)  # agents below zero were removed (reduce_energy!, strict <); offspring of a killed parent may be below
        prev = sa


@pytest.mark.gpu
@pytest.mark.parametrize("label,wt,nx,ny,ticks,extra", [
    ("C2 full size", 2, 4096, 4096, 2, dict(visual_distance=3, prob_random_walk=0.9, reproduction_prob=0.001)),  # BASELINE.json configs[1]
    ("C3 slab", 4, 8192, 1024, 2, dict(walkmap_regrowth=1, reproduction_prob=0.004)),                              # configs[2] rules, an 8192 x 1024 slab
])
def test_gpu_baseline_configs_match_the_oracle(libabm, oracle, label, wt, nx, ny, ticks, extra):
    """The BASELINE.json configurations against the ORACLE (not against a second GPU path): the oracle runs ~1.4e6
    agent-updates/s, so two ticks of the full C2 world (2^22 sheep) are affordable."""
    from abm_b200 import worlds
    wm = el = None
    if wt == 4:
        el = worlds.make_terrain(nx, ny, seed=6)
        wm = worlds.walkmap_from_elevation(el)
    w = worlds.make_world(nx, ny, nx * ny // 4, seed=5, walkmap=wm, periodic=True)
    w["walkmap"], w["elevation"] = wm, el
    cfg = dict(nx=nx, ny=ny, periodic=1, walk_type=wt, seed=77, **extra)
    a, b = load_case(oracle, cfg, w), load_case(libabm, cfg, w)
    for t in range(ticks):
        a.step(1)
        b.step(1)
        assert_same_state(b.snapshot(), a.snapshot(), f"{label} tick {t}")
    a.close(); b.close()


@pytest.mark.parametrize("name", ["tile_rw_96x64", "tile_gregarious_r3_96x96", "tile_walkmap_128x96", "tile_crowded_64x64"])
def test_emulated_tile_path_with_migrant_lists(emu, oracle, name, monkeypatch):
    """ABM_MIGRANTS=1: the commit reads its own core plus the per-core list of border crossers (off by default)."""
    monkeypatch.setenv("ABM_MIGRANTS", "1")
    monkeypatch.setenv("ABM_STREAM", "0")  # the walkmap case would otherwise take the stream path (no commit kernel)
    cfgkw, w, state, ticks = build_case(name)
    a, b = load_case(oracle, cfgkw, w), load_case(emu, cfgkw, w)
    for t in range(6):
        a.step(1)
        b.step(1)
        assert "migrant_scan" in {k for k, v in b.kernel_stats().items() if v["launches"]}
        assert_same_state(b.snapshot(), a.snapshot(), f"{name} tick {t}")


def _edge_world(nx, ny, n, energy):
    rng = np.random.default_rng(3)
    fg = rng.integers(0, 2, (ny, nx)).astype(np.uint8)
    cd = np.where(fg == 1, 5, rng.integers(0, 5, (ny, nx))).astype(np.int64)
    ids = np.arange(1, n + 1, dtype=np.int64)
    return dict(fully_grown=fg, countdown=cd, ids=ids, x=rng.integers(1, nx + 1, n).astype(np.int64),
                y=rng.integers(1, ny + 1, n).astype(np.int64), energy=np.full(n, energy, np.float64))


def _edge_run(lib_a, lib_b, nx, ny, n, energy, wt, ticks=6, **kw):
    cfgkw = dict(nx=nx, ny=ny, periodic=1, walk_type=wt, eat=0, reproduce=1, regrowth_time=5, reproduction_prob=0.3, seed=77, **kw)
    w = _edge_world(nx, ny, n, energy)
    a, b = load_case(lib_a, cfgkw, w), load_case(lib_b, cfgkw, w)
    for t in range(ticks):
        a.step(1)
        b.step(1)
        assert_same_state(b.snapshot(), a.snapshot(), f"{nx}x{ny} n={n} tick {t}")
    return b


@pytest.mark.parametrize("nx,ny,wt", [(64, 64, 0), (64, 96, 2), (40, 30, 0), (1, 1, 0), (64, 64, 4)])
def test_emulated_empty_world_and_extinction(emu, oracle, nx, ny, wt):
    """No sheep at all (the grass still regrows), and a herd that starves on its first step (energy 0, no eating:
    reduce_energy! kills at energy < 0 after the move, reference src/agent_actions.jl:85-92) and leaves an empty world."""
    b = _edge_run(oracle, emu, nx, ny, 0, 0.0, wt)
    assert b.count_agents() == 0
    b = _edge_run(oracle, emu, nx, ny, min(50, 3 * nx * ny), 0.5, wt)
    assert b.count_agents() == 0


@pytest.mark.gpu
@pytest.mark.parametrize("nx,ny,wt", [(64, 64, 0), (64, 96, 2), (40, 30, 0), (1, 1, 0)])
def test_gpu_empty_world_and_extinction(libabm, oracle, nx, ny, wt):
    b = _edge_run(oracle, libabm, nx, ny, 0, 0.0, wt)
    assert b.count_agents() == 0
    b = _edge_run(oracle, libabm, nx, ny, min(50, 3 * nx * ny), 0.5, wt)
    assert b.count_agents() == 0


def _init_pair(lib_a, lib_b, nx, ny, n, wt, terrain, seed=11, **kw):
    from conftest import make_case
    cfgkw, w = make_case(nx, ny, 1, wt, 1, terrain=terrain, seed=seed, regrowth_time=6, delta_energy=3.0, reproduction_prob=0.05, **kw)
    out = []
    for lib in (lib_a, lib_b):
        e = lib.create(lib.default_config(**cfgkw))
        if terrain:  # only the terrain of this call survives abm_init_random
            e.set_grid(np.zeros((ny, nx), np.uint8), np.zeros((ny, nx), np.int64), w["walkmap"], w["elevation"])
        e.init_random(n, 4242)
        out.append(e)
    return out[0], out[1], w


@pytest.mark.parametrize("nx,ny,n,wt,terrain", [(40, 30, 300, 0, False), (64, 64, 1500, 2, False), (96, 64, 2000, 4, True), (33, 7, 0, 0, False), (64, 96, 900, 3, True)])
def test_emulated_device_side_initialisation(emu, oracle, nx, ny, n, wt, terrain):
    """abm_init_random (the random part of initialize_model): engine and oracle build the same world from
    the per-sheep / per-cell streams, the world has the reference's shape, and stepping it stays bit-exact."""
    a, b, w = _init_pair(oracle, emu, nx, ny, n, wt, terrain, counter=5, astar_metric=1, astar_penalty=1)
    sa = a.snapshot()
    assert_same_state(b.snapshot(), sa, "after init")
    assert sa["ids"].tolist() == list(range(1, n + 1)) and sa["next_id"] == n + 1
    assert ((sa["energy"] >= 0) & (sa["energy"] <= 14) & (sa["energy"] == np.floor(sa["energy"]))).all()  # rand(1:3*5) - 1
    walk = np.ones((ny, nx), bool) if not terrain else w["walkmap"].astype(bool)
    assert walk[sa["y"] - 1, sa["x"] - 1].all()
    fg, cd = sa["fully_grown"].astype(bool), sa["countdown"]
    assert not fg[~walk].any() and (cd[~walk] == 0).all()          # scripts/v0.5.jl:115-122
    assert (cd[fg] == 6).all() and ((cd[walk & ~fg] >= 0) & (cd[walk & ~fg] <= 5)).all()
    if walk.sum() > 500:
        assert 0.4 < fg[walk].mean() < 0.6
    for e in (a, b):
        if wt == 3:
            e.set_global_state(0, 0, 5)
    for t in range(5):
        a.step(1)
        b.step(1)
        assert_same_state(b.snapshot(), a.snapshot(), f"tick {t} after init")


def test_device_side_initialisation_on_a_strip_needs_the_whole_walkmap(emu, oracle):
    """abm_init_random on a strip handle: random_walkable draws positions over the whole GridSpace, so a strip with terrain
    refuses until the global walkmap was given; then (a ring of one here) it builds the single handle's world."""
    from abm_b200 import capi as _capi, strips
    cfgkw, w = make_case(64, 96, 500, 4, 1, terrain=True, seed=9, walkmap_regrowth=1)
    e = strips.load_strip(emu, cfgkw, w, 1, 0)
    with pytest.raises(_capi.AbmError) as ei:
        e.init_random(500, 3)
    assert ei.value.code == _capi.ABM_E_STATE
    e.set_global_walkmap(w["walkmap"])
    e.init_random(500, 3)
    a = load_case(oracle, cfgkw, w)
    a.init_random(500, 3)
    for t in range(4):
        a.step(1); e.step(1)
        assert_same_state(strips.merge_snapshots([strips.strip_snapshot(e)]), a.snapshot(), f"strip init tick {t}")


@pytest.mark.gpu
@pytest.mark.parametrize("nx,ny,n,wt,terrain", [(100, 70, 2000, 0, False), (128, 96, 6000, 4, True)])
def test_gpu_device_side_initialisation(libabm, oracle, nx, ny, n, wt, terrain):
    a, b, _ = _init_pair(oracle, libabm, nx, ny, n, wt, terrain)
    assert_same_state(b.snapshot(), a.snapshot(), "after init")
    for t in range(5):
        a.step(1)
        b.step(1)
        assert_same_state(b.snapshot(), a.snapshot(), f"tick {t} after init")


# ------------------------------------------------------------------ narrow transfer forms (abm_*_packed)

def _check_packed_round_trip(lib, oracle):
    """abm_set_grass_packed / abm_set_agents_packed feed the same world as the wide forms, the *_packed getters return the
    same state, and a host that round-trips the model through them every tick stays on the oracle's trajectory."""
    for walk_type, terrain, nx, ny in ((2, False, 96, 64), (4, True, 64, 96), (0, False, 33, 20)):
        cfgkw, w = make_case(nx, ny, nx * ny // 4, walk_type, 1 if walk_type else 0, seed=77, terrain=terrain, visual_distance=3, reproduction_prob=0.02,
                             walkmap_regrowth=1 if terrain else 0)
        ref = load_case(oracle, cfgkw, w)
        e = load_case(lib, cfgkw, w)  # terrain through the wide abm_set_grid
        st = (w["fully_grown"].astype(np.uint8) << 7) | w["countdown"].astype(np.uint8)
        for tick in range(4):
            e.set_grass_packed(st)
            if tick == 0:
                e.set_agents_packed(w["ids"], w["x"].astype(np.uint32) | (w["y"].astype(np.uint32) << 16), w["energy"])
            else:
                e.set_agents_packed(ids, xy, en, nid)
            e.set_global_state(tick)
            e.step(1)
            ref.step(1)
            ids, xy, en = e.get_agents_packed()
            st = e.get_grass_packed()
            nid = e.get_global_state()["next_id"]
            o = ref.snapshot()
            assert np.array_equal(ids, o["ids"]) and np.array_equal(xy & 0xFFFF, o["x"]) and np.array_equal(xy >> 16, o["y"]), (walk_type, tick)
            assert np.array_equal(en, o["energy"]) and nid == o["next_id"]
            assert np.array_equal(st >> 7, o["fully_grown"]) and np.array_equal(st & 0x7F, o["countdown"])
        # the oracle's own packed forms agree with its wide ones
        oi, oxy, oe = ref.get_agents_packed()
        assert np.array_equal(oi, ids) and np.array_equal(oxy, xy) and np.array_equal(ref.get_grass_packed(), st)
        with pytest.raises(capi.AbmError):
            e.set_grass_packed(np.full((ny, nx), 0x7F, np.uint8))  # countdown 127 is the padding value
        e.close(); ref.close()


def test_packed_transfer_forms_emulated(emu, oracle):
    _check_packed_round_trip(emu, oracle)


@pytest.mark.gpu
def test_gpu_packed_transfer_forms(libabm, oracle):
    _check_packed_round_trip(libabm, oracle)


# ------------------------------------------------------------------ stream path (abm_stream.h): rules without conflicts

def _check_stream_path(lib, oracle, monkeypatch, cap):
    """RANDOM_WALKMAP and RANDOM_WALK without `ifempty` on tile grids take the one-kernel-per-tick stream path: arrival lists
    per core, the next step drawn in the same kernel.  Checked against the oracle tick by tick, with the getters and
    setters of the ABI called between the ticks (they convert between the lists and the canonical arrays), with few list
    slots per core (overflow list, re-filing) and across several ticks per abm_step."""
    if cap:
        monkeypatch.setenv("ABM_STREAM_CAP", str(cap))
    cases = [
        dict(nx=128, ny=96, n=9000, wt=4, per=1, terrain=True, kw=dict(walkmap_regrowth=1, reproduction_prob=0.05, regrowth_time=3, delta_energy=3.0)),
        dict(nx=64, ny=96, n=7000, wt=0, per=0, terrain=False, kw=dict(randomwalk_ifempty=0, reproduction_prob=0.1, regrowth_time=2, delta_energy=2.0)),
        dict(nx=96, ny=64, n=300, wt=0, per=1, terrain=False, kw=dict(randomwalk_ifempty=0, reproduction_prob=0.3, regrowth_time=1, delta_energy=5.0)),
        dict(nx=64, ny=64, n=20000, wt=4, per=1, terrain=True, kw=dict(walkmap_regrowth=0, reproduction_prob=0.02, regrowth_time=4, eat=0)),
        # everybody starts on one core of sixteen: the lists are filed again with more slots per core
        dict(nx=128, ny=128, n=12000, wt=0, per=1, terrain=False, crowd=True, kw=dict(randomwalk_ifempty=0, reproduction_prob=0.05, regrowth_time=2, delta_energy=3.0)),
    ]
    for ci, c in enumerate(cases):
        cfgkw, w = make_case(c["nx"], c["ny"], c["n"], c["wt"], c["per"], terrain=c["terrain"], seed=41 + ci, **c["kw"])
        if c.get("crowd"):
            w["x"], w["y"] = 33 + (w["x"] - 1) % 32, 65 + (w["y"] - 1) % 32
        a, b = load_case(oracle, cfgkw, w), load_case(lib, cfgkw, w)
        assert "stream_tick" not in b.kernel_stats()
        for t in range(5):
            a.step(1); b.step(1)
            assert_same_state(b.snapshot(), a.snapshot(), f"stream case {ci} tick {t}")
        assert b.kernel_stats()["stream_tick"]["launches"] == 1
        a.step(4); b.step(4)  # several ticks per call: no getter in between
        assert_same_state(b.snapshot(), a.snapshot(), f"stream case {ci} after step(4)")
        assert b.reduce(capi.REDUCE_SUM_ENERGY) == a.reduce(capi.REDUCE_SUM_ENERGY) and b.count_agents() == a.count_agents()
        # the host rewinds the clock: the filed steps were drawn for another tick and must be drawn again
        a.set_global_state(2); b.set_global_state(2)
        a.step(2); b.step(2)
        assert_same_state(b.snapshot(), a.snapshot(), f"stream case {ci} after a tick change")
        # new grass through the narrow form keeps the filed steps; a new population replaces them
        st = b.get_grass_packed()
        st[::2, ::3] = 0x80 | 3
        a.set_grass_packed(st); b.set_grass_packed(st)
        a.step(1); b.step(1)
        assert_same_state(b.snapshot(), a.snapshot(), f"stream case {ci} after new grass")
        s = a.snapshot()
        keep = s["ids"] % 3 != 0
        for e in (a, b):
            e.set_agents(s["ids"][keep], s["x"][keep], s["y"][keep], s["energy"][keep] + 1.0, s["next_id"])
            e.step(3)
        assert_same_state(b.snapshot(), a.snapshot(), f"stream case {ci} after a new population")
        a.close(); b.close()


@pytest.mark.parametrize("cap", [0, 8])
def test_stream_path_emulated(emu, oracle, monkeypatch, cap):
    _check_stream_path(emu, oracle, monkeypatch, cap)


@pytest.mark.gpu
@pytest.mark.parametrize("cap", [0, 8])
def test_gpu_stream_path(libabm, oracle, monkeypatch, cap):
    _check_stream_path(libabm, oracle, monkeypatch, cap)
