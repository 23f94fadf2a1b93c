#!/usr/bin/env python
"""bench.py — headline measurement of the per-tick stepping path (BASELINE.json metric: agent-updates/s).

Workload (BASELINE.json configs[1], "C2"): v0.3 rules — RANDOM_WALK_GREGARIOUS, visual_distance = 3,
prob_random_walk = 0.9, eat + reproduce + grass regrowth — on a periodic 4096 x 4096 GridSpace with 2^22 sheep,
synthetic seeded world (abm_b200.worlds).  A "step" is one tick: every agent's agent_step! in id order, then
model_step!.  At N > 1 the GridSpace is 4096 x (4096 N) with 2^22 N sheep, cut into N row strips (one per GPU, weak
scaling); abm_step exchanges the boundary rows with the ring neighbours every tick — over NVLink peer memory
(kernels store into the neighbours' mailboxes; `--transport auto`, the default, falls back to NCCL when a rank cannot map
its peers).  Other workloads: `--workload C3` (v0.5 walkmap rules, 8192^2), `C5` (the same rules on 32768 x 4096 rows per
GPU: 32768^2 and 2^28 sheep at 8 GPUs), `C1`, and `C4` (batched A* replans; prints its own line).

  value      agent-updates/s with the world resident in HBM; per-tick CUDA-event time on the library's stream,
             L2 flushed (256 MiB write) between ticks, max over ranks.
  e2e        the same metric through the C ABI with HOST buffers: per tick abm_set_grass_packed + abm_set_agents_packed
             (pinned host -> device), abm_step(1), abm_get_agents_packed + abm_get_grass_packed (device -> pinned host).
  roofline   dominant kernel: algorithmic bytes (A = 36 B per agent-update, C = 2 B per cell-update,
             +1 B per cell for the gregarious occupancy read) / its CUDA-event time, against MEASURED_PEAKS.json.
  configs    (N = 1) short driver-run records of the other single-GPU configurations: C3 (configs[2]).
  c5         every N: the configuration the weak-scaling target is quoted on (configs[4]): 32768 x 4096 rows and 2^25 sheep
             per GPU, v0.5 rules, world built on the device (abm_init_random).
  verify     N > 1: after the timed ticks every strip's state is hashed and compared with a single-handle run of the same
             global world on rank 0's GPU.
  cpu_baseline / --impl reference: the CPU restatement of the reference (oracle/, sequential like Agents.jl's
             step!) on the same configuration (a bounded number of ticks).  The oracle is only the timed baseline here.
"""
import argparse
import hashlib
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

A_BYTES, C_BYTES, C_BYTES_GREG = 36.0, 2.0, 3.0
WORKLOADS = {
    # name: (walk_type, terrain, extra config)
    "C2": dict(walk_type=2, terrain=False, cfg=dict(visual_distance=3, prob_random_walk=0.9, reproduction_prob=0.001),
               desc="v0.3 RANDOM_WALK_GREGARIOUS r=3, periodic {n}x{m} grid, {a} sheep at tick 0 (BASELINE.json configs[1])"),
    "C3": dict(walk_type=4, terrain=True, cfg=dict(walkmap_regrowth=1, reproduction_prob=0.004),
               desc="v0.5 RANDOM_WALKMAP + walkmap regrowth, periodic {n}x{m} terrain, {a} sheep at tick 0 (BASELINE.json configs[2])"),
    "C5": dict(walk_type=4, terrain=True, cfg=dict(walkmap_regrowth=1, reproduction_prob=0.004),
               desc="v0.5 RANDOM_WALKMAP + walkmap regrowth, periodic {n}x{m} terrain, {a} sheep at tick 0 (BASELINE.json configs[4] per-GPU share: 32768x4096 rows of 32768x(4096 N))"),
    "C1": dict(walk_type=0, terrain=False, cfg=dict(reproduction_prob=0.001),
               desc="v0.1 RANDOM_WALK, {n}x{m} grid, {a} sheep at tick 0"),
    # the v0.4 / v0.6 rules (src/agent_actions.jl:211-242): model-global behaviour counter, a third of the switches plans an A*
    # route (PenaltyMap(elevation, MaxDistance), scripts/v0.6.jl:87-93).  General path, one GPU; the batched A* dominates the tick.
    "V6": dict(walk_type=3, terrain=True, cfg=dict(counter=50, reproduction_prob=0.004, astar_metric=1, astar_penalty=1, walkmap_regrowth=1),
               desc="v0.6 ALTERNATED_WALK (counter 50) + A* routes on PenaltyMap(elevation, MaxDistance), periodic {n}x{m} terrain, {a} sheep at tick 0"),
}
DEFAULT_SIZE = {"C2": (4096, 4096), "C3": (8192, 8192), "C5": (32768, 4096), "C1": (4096, 4096), "C4": (2048, 2048), "V6": (1024, 1024)}  # (nx, rows per GPU)


def make_world(workload, size, agents, seed=321, nx=None, rank=0, world=1):
    """One rank's share, generated with numpy: `size` rows of an nx-wide world of size * world rows (nx = size,
    world = 1: the square single-GPU worlds).  With world > 1 the terrain arrays carry the 32 ghost rows above and below
    (true neighbours' terrain, the y wrap included); grass and agents cover the own rows only."""
    from abm_b200 import worlds
    wl = WORKLOADS[workload]
    nx = nx or size
    wm = el = place = None
    if wl["terrain"] and (world > 1 or nx != size):
        from abm_b200 import strips
        gh = strips.GHOST if world > 1 else 1
        el = worlds.make_terrain_rows(nx, size * world, size * rank - gh, size + 2 * gh, seed=322)
        wm = worlds.walkmap_from_elevation(el)
        place = (wm.astype(bool) & worlds.has_walkable_neighbour(wm, True))[gh:-gh]
        if world == 1:
            el, wm = el[gh:-gh], wm[gh:-gh]
        w = worlds.make_world(nx, size, agents, seed=seed, walkmap=wm[gh:-gh] if world > 1 else wm, periodic=True, placeable=place)
        w["walkmap"], w["elevation"] = wm, el
        cfg = dict(nx=nx, ny=size, periodic=1, walk_type=wl["walk_type"], seed=seed)
        cfg.update(wl["cfg"])
        return cfg, w
    if wl["terrain"]:
        el = worlds.make_terrain(nx, size, seed=seed + 1)
        wm = worlds.walkmap_from_elevation(el)
    w = worlds.make_world(nx, size, agents, seed=seed, walkmap=wm, periodic=True)
    w["walkmap"], w["elevation"] = wm, el
    cfg = dict(nx=nx, ny=size, periodic=1, walk_type=wl["walk_type"], seed=seed)
    cfg.update(wl["cfg"])
    return cfg, w


def load(lib, cfg, w, device=0):
    e = lib.create(lib.default_config(device=device, **cfg))
    e.set_grid(w["fully_grown"], w["countdown"], w.get("walkmap"), w.get("elevation"))
    e.set_agents(w["ids"], w["x"], w["y"], w["energy"])
    if cfg.get("walk_type") == 3:  # model.behav_counter[1] = counter, scripts/v0.4.jl:97
        e.set_global_state(0, 0, cfg["counter"])
    return e


def device_walkmap(nx, ny, seed=322, octaves=4, base=64):
    """The walkmap of the synthetic terrain (abm_b200.worlds.make_terrain_rows + walkmap_from_elevation: periodic value noise
    quantised to 1..40, walkable iff 1 < e < 18, scripts/v0.5.jl:56-68) computed with torch on the GPU — the numpy generator
    needs half a minute for one C5 strip.  Cells whose eight neighbours are all blocked are closed (random_walkmap throws
    on them, rand(1:0)).  Returns a (ny, nx) uint8 CUDA tensor; every rank computes the same map from the seed."""
    import torch
    from abm_b200 import worlds
    rng = worlds._rng(seed)
    lattices = []
    for o in range(octaves):
        cell = max(2, base >> o)
        gx, gy = max(1, -(-nx // cell)), max(1, -(-ny // cell))
        lattices.append((cell, rng.random((gy, gx))))
    sy, sx = max(1, ny // 1024), max(1, nx // 1024)
    coarse = worlds._noise_rows(nx, ny, lattices, np.arange(0, ny, sy), np.arange(0, nx, sx))
    lo, hi = np.quantile(coarse, 0.02), np.quantile(coarse, 0.98)
    dev = torch.device("cuda")
    out = torch.empty((ny, nx), dtype=torch.uint8, device=dev)
    cols = torch.arange(nx, device=dev, dtype=torch.float64)
    lats = [(cell, torch.from_numpy(lat).to(dev)) for cell, lat in lattices]
    chunk = max(1, (1 << 24) // nx)
    for r0 in range(0, ny, chunk):
        r1 = min(ny, r0 + chunk)
        rows = torch.arange(r0, r1, device=dev, dtype=torch.float64)
        h = torch.zeros((r1 - r0, nx), dtype=torch.float64, device=dev)
        amp, total = 1.0, 0.0
        for cell, lat in lats:
            gy, gx = lat.shape
            ys, xs = (rows / cell) % gy, (cols / cell) % gx
            y0, x0 = ys.floor().long(), xs.floor().long()
            fy, fx = ys - y0, xs - x0
            fy, fx = fy * fy * (3 - 2 * fy), fx * fx * (3 - 2 * fx)
            y1, x1 = (y0 + 1) % gy, (x0 + 1) % gx
            top = lat[y0][:, x0]
            top = top + (lat[y0][:, x1] - top) * fx[None, :]
            bot = lat[y1][:, x0]
            bot = bot + (lat[y1][:, x1] - bot) * fx[None, :]
            h += amp * (top + (bot - top) * fy[:, None])
            total += amp
            amp *= 0.5
        h = ((h / total - lo) / max(hi - lo, 1e-12)).clamp_(0.0, 1.0)
        e = torch.floor(h * (39 * 0.62)) + 1
        out[r0:r1] = ((e > 1) & (e < 18)).to(torch.uint8)
    # close the cells without a walkable neighbour (one pass: a pair of adjacent cells keeps both)
    nb = torch.zeros_like(out)
    for dy in (-1, 0, 1):
        for dx in (-1, 0, 1):
            if dx or dy:
                nb |= torch.roll(out, (dy, dx), (0, 1))
    return out & nb


def load_device_world(lib, workload, nx, size, agents_total, rank, world, device, seed=321, stationary=False):
    """C3 / C5 worlds without host arrays: the terrain's walkmap from the GPU generator, grass and sheep drawn by the engine
    itself (abm_init_random; on strips every rank keeps the sheep that land on its rows)."""
    import torch
    from abm_b200 import strips
    wl = WORKLOADS[workload]
    ny = size * world
    wm = device_walkmap(nx, ny) if wl["terrain"] else None
    cfg = dict(nx=nx, ny=ny, periodic=1, walk_type=wl["walk_type"], seed=seed, device=device)
    cfg.update(wl["cfg"])
    if stationary:
        cfg.update(reproduce=0, movement_cost=0.0)
    if world > 1:
        cfg.update(strip_y0=size * rank, strip_ny=size, n_ranks=world, rank=rank)
    e = lib.create(lib.default_config(**cfg))
    frac = None
    if world > 1:
        e.ny = size + 2 * strips.GHOST
    if wm is not None:
        if world > 1:
            rows = (torch.arange(size * rank - strips.GHOST, size * (rank + 1) + strips.GHOST, device=wm.device)) % ny
            e.set_grid(None, None, wm[rows].cpu().numpy(), None)
            w32 = wm.view(-1, 32).to(torch.int64)
            bits = (w32 << torch.arange(32, device=wm.device, dtype=torch.int64)[None, :]).sum(1)
            e.set_global_walkmap_bits((bits & 0xFFFFFFFF).cpu().numpy().astype(np.uint32))
            del w32, bits
        else:
            e.set_grid(None, None, wm.cpu().numpy(), None)
        frac = float(wm.float().mean().item())
        del wm
        torch.cuda.empty_cache()
    e.init_random(agents_total, seed)
    return cfg, e, frac


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self.stop_flag, self.proc = index, [], False, None

    def run(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100", "-i", str(self.index)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                self.rows.append([t.strip() for t in line.split(",")] + [time.perf_counter()])
                if self.stop_flag:
                    break
        except Exception:
            pass

    def finish(self, t0=None, t1=None):
        self.stop_flag = True
        if self.proc is not None:
            self.proc.terminate()
        rows = [r for r in self.rows if len(r) >= 10]
        inside = [r for r in rows if t0 is not None and t0 - 0.15 <= r[-1] <= t1 + 0.15]
        window = "timed region"
        if not inside:  # the timed region is shorter than nvidia-smi's sampling period: use the whole loaded span
            inside, window = rows, "warm-up + timed region + e2e (timed region shorter than the sampling period)"
        self.rows = inside
        sm = [float(r[1]) for r in self.rows if r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if r[2].replace(".", "").isdigit()]
        reasons = set()
        self.window = window
        for r in self.rows:
            if len(r) >= 9:
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        return dict(sm_mhz=float(np.median(sm)) if sm else None, sm_max_mhz=max(mx) if mx else None, reasons=sorted(reasons), samples=len(sm), window=window)


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def traffic_for(workload, kernel):
    """DRAM bytes per launch of `kernel` on `workload` from the committed ncu capture (profiles/traffic.json), or None."""
    tp = os.path.join(ROOT, "profiles", "traffic.json")
    if not os.path.exists(tp):
        return None
    t = json.load(open(tp))
    v = t.get(workload)
    return v.get(kernel) if isinstance(v, dict) else None


def oracle_lib():
    from abm_b200 import build, capi
    return capi.CLib(build.build_oracle(), "oracle_")


def time_oracle(workload, nx, size, agents, ticks, warm=0, stationary=False):
    """agent-updates/s of the CPU restatement (sequential, 1 thread — Agents.jl's step! is serial)."""
    lib = oracle_lib()
    cfg, w = make_world(workload, size, agents, nx=nx)
    if stationary:
        cfg.update(reproduce=0, movement_cost=0.0)
    e = load(lib, cfg, w)
    e.step(warm)
    upd, t0 = 0, time.perf_counter()
    for _ in range(ticks):
        upd += e.count_agents()
        e.step(1)
    dt = time.perf_counter() - t0
    e.close()
    return upd / dt, dt, upd


def run_astar(args):
    """BASELINE.json configs[3] (C4): batched A* replans (PenaltyMap(elevation, MaxDistance), scripts/v0.6.jl:87-93) on
    synthetic terrain.  Not the headline metric: prints its own line (replans/s, expansions/s) for profiles/."""
    from abm_b200 import build, capi, worlds
    size, n = args.size, args.replans
    el = worlds.make_terrain(size, size, seed=322)
    wm = worlds.walkmap_from_elevation(el)
    w = worlds.make_world(size, size, n, seed=321, walkmap=wm, periodic=True)
    ok = np.flatnonzero(wm.reshape(-1))
    rng = np.random.Generator(np.random.Philox(key=99))
    goal = ok[rng.integers(0, ok.size, n)]  # random_walkable: uniform over walkable cells
    if args.max_dist:  # bounded replans (local goals): |goal - start| <= max_dist per axis, still walkable
        gx = np.clip(w["x"] - 1 + rng.integers(-args.max_dist, args.max_dist + 1, n), 0, size - 1)
        gy = np.clip(w["y"] - 1 + rng.integers(-args.max_dist, args.max_dist + 1, n), 0, size - 1)
        lin = gy * size + gx
        keep = wm.reshape(-1)[lin] == 1
        goal = np.where(keep, lin, (w["y"] - 1) * size + (w["x"] - 1))
    gx, gy = goal % size + 1, goal // size + 1
    cfg = dict(nx=size, ny=size, periodic=1, walk_type=3, astar_metric=1, astar_penalty=1, counter=50, seed=321)
    res = {}
    for name, lib, m in (("b200", capi.CLib(build.LIBABM, "abm_"), n), ("oracle", oracle_lib(), min(n, args.ref_replans))):
        e = lib.create(lib.default_config(**cfg))
        e.set_grid(w["fully_grown"], w["countdown"], wm, el)
        e.set_agents(w["ids"], w["x"], w["y"], w["energy"])
        e.set_profiling(True) if name == "b200" else None
        # warm-up: every agent "replans" to the cell it stands on (found at once, no expansion) so that the search slots
        # and the route pool are allocated outside the timed call
        e.plan_routes(w["ids"][:m], w["x"][:m], w["y"][:m])
        t0 = time.perf_counter()
        e.plan_routes(w["ids"][:m], gx[:m], gy[:m])
        dt = time.perf_counter() - t0
        tm = e.timing()
        off, px, py, cost = e.get_paths(w["ids"][: min(m, 2000)])
        res[name] = dict(replans=m, seconds=dt, device_ms=tm["ms_total"], expansions=tm["astar_expansions"], replans_per_s=m / dt,
                         expansions_per_s=tm["astar_expansions"] / dt, mean_path_cells=float(off[-1]) / max(1, min(m, 2000)), found=int((cost >= 0).sum()))
        e.close()
    print(json.dumps({"metric": "astar_replans_per_s", "value": res["b200"]["replans_per_s"], "unit": "replans/s", "n_gpus": 1,
                      "config": {"workload": f"v0.6 batched A* replans, PenaltyMap(elevation, MaxDistance), {size}x{size} synthetic terrain, {n} replans"
                                 + (f", goals within {args.max_dist} cells" if args.max_dist else ", goals = random_walkable")},
                      "b200": res["b200"], "cpu_baseline": dict(kind="port", cores=1, **res["oracle"]), "data": "synthetic", "dtype": "int32 costs"}))


def run_reference(args, rank, world):
    """--impl reference: the reference's CPU path (oracle port; Julia is not installed) on the b200 arm's configuration,
    all the host threads it can use (1: Agents.jl's step! is serial).  Rank 0 alone runs it."""
    if rank != 0:
        return
    nx, size = args.ref_nx, args.ref_size
    agents = nx * size // 4
    wl = WORKLOADS[args.workload]
    lib = oracle_lib()
    cfg, w = make_world(args.workload, size, agents, nx=nx)
    if args.stationary:
        cfg.update(reproduce=0, movement_cost=0.0)
    e = load(lib, cfg, w)
    # one "step" = one tick of the b200 arm's own configuration; the sequential CPU path needs seconds per tick at that size,
    # so the warm-up is one tick and the timed ticks stop after --ref-budget seconds (at least two are timed)
    warm = min(args.warmup, 1)
    e.step(warm)
    upd, timed, t0 = 0, 0, time.perf_counter()
    for _ in range(args.steps):
        upd += e.count_agents()
        e.step(1)
        timed += 1
        if timed >= 2 and time.perf_counter() - t0 > args.ref_budget:
            break
    dt = time.perf_counter() - t0
    v = upd / dt
    same = (nx, size) == (args.size_x, args.size)
    sample = f"{nx}x{size} grid, {agents} sheep at tick 0, {warm} warm-up + {timed} timed ticks ({dt:.0f} s)" + (" (the b200 arm's per-GPU configuration)" if same else f", same rules/density as the {args.size_x}x{args.size} workload")
    print(json.dumps({
        "impl": "reference", "metric": "agent_updates_per_s", "value": v, "unit": "agent-updates/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "steps_timed": timed, "ms_per_step": 1e3 * dt / timed, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "int32 cells / f64 energy", "data": "synthetic",
        "config": {"workload": wl["desc"].format(n=args.size_x, m=args.size, a=args.size_x * args.size // 4) + (" [stationary: reproduce=false, movement_cost=0]" if args.stationary else ""),
                   "sample": sample, "same_config": same, "agents_live_mean": upd / timed},
        "cpu_baseline": {"value": v, "unit": "agent-updates/s", "cores": 1, "kind": "port", "sample": sample,
                         "note": "C++ restatement of the reference rules + Agents.jl semantics (oracle/); Julia is not installed; Agents.jl step! is serial, hence 1 thread"},
        "e2e": {"value": v, "unit": "agent-updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def bind_to_gpu_numa_node(local):
    """Pinned staging buffers should live on the NUMA node the GPU hangs off: with 8 ranks sharing one host the copies
    otherwise cross the socket link.  Best effort (sysfs); returns the node or None."""
    try:
        bus = subprocess.run(["nvidia-smi", "--query-gpu=pci.bus_id", "--format=csv,noheader", "-i", str(local)], capture_output=True, text=True).stdout.strip().lower()
        if bus.startswith("00000000:"):
            bus = "0000:" + bus[9:]
        node = int(open(f"/sys/bus/pci/devices/{bus}/numa_node").read())
        if node < 0:
            return None
        ids = set()
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            a, _, b = part.partition("-")
            ids.update(range(int(a), int(b or a) + 1))
        os.sched_setaffinity(0, (ids & os.sched_getaffinity(0)) or ids)
        return node
    except Exception:
        return None


def state_digest(eng, own_rows, ghost):
    """sha256 over a strip's (or a whole handle's) observable state: agents in id order, grass of the own rows."""
    ids, xy, en = eng.get_agents_packed()
    g = eng.get_grass_packed()
    g = g[ghost:ghost + own_rows] if ghost else g
    h = hashlib.sha256()
    for a in (ids, xy, en.view(np.uint64), np.ascontiguousarray(g)):
        h.update(np.ascontiguousarray(a).tobytes())
    return h.hexdigest(), int(ids.size)


def measure(lib, args, workload, nx, size, steps, warmup, e2e_steps, rank, world, local, dist, sampler=None, device_world=False, verify=False, kernel_events=True):
    """Builds the world of `workload` (nx wide, `size` rows per GPU), runs warm-up + timed ticks (+ the end-to-end loop) and
    returns the record (rank 0) — the body of the JSON line and of its sub-records."""
    import torch
    from abm_b200 import strips
    wl = WORKLOADS[workload]
    agents = nx * size // 4
    t_build0 = time.perf_counter()
    walk_frac = None
    if device_world:
        cfg, eng, walk_frac = load_device_world(lib, workload, nx, size, agents * world, rank, world, local, stationary=args.stationary)
    else:
        cfg, w = make_world(workload, size, agents, seed=321 + rank, nx=nx, rank=rank, world=world)
        if args.stationary:
            cfg.update(reproduce=0, movement_cost=0.0)
        if world > 1:
            # strip `rank` of an nx x (size * world) GridSpace: this rank's rows and sheep (global ids and y), ghost rows for the grids
            cfg.update(ny=size * world, strip_y0=size * rank, strip_ny=size, n_ranks=world, rank=rank, seed=321, device=local)
            eng = lib.create(lib.default_config(**cfg))
            pad = lambda a: np.concatenate([np.zeros((strips.GHOST, nx), a.dtype), a, np.zeros((strips.GHOST, nx), a.dtype)], 0)
            eng.ny = size + 2 * strips.GHOST
            eng.set_grid(pad(w["fully_grown"]), pad(w["countdown"]), w["walkmap"], w["elevation"])  # terrain comes with its ghost rows
            eng.set_agents(w["ids"] + rank * agents, w["x"], w["y"] + size * rank, w["energy"], agents * world + 1)
        else:
            eng = load(lib, cfg, w, device=local)
        del w
    transport = args.transport
    if world > 1:
        if transport == "auto":
            # peer-memory mailboxes need CUDA IPC + peer access between all ranks of the node; every rank must agree on the choice
            try:
                strips.init_p2p(eng)
                ok = 1
            except Exception as exc:  # noqa: BLE001 -- any failure to map a peer means NCCL for everybody
                print(f"[bench rank {rank}] peer-memory transport unavailable ({exc}); trying NCCL", file=sys.stderr)
                ok = 0
            flag = torch.tensor([ok], device="cuda")
            dist.all_reduce(flag, op=dist.ReduceOp.MIN)
            transport = "p2p" if int(flag.item()) else "nccl"
            if transport == "nccl":
                strips.init_nccl(eng)
        else:
            (strips.init_p2p if transport == "p2p" else strips.init_nccl)(eng)
    t_build = time.perf_counter() - t_build0
    eng.set_profiling(kernel_events)
    G = nx * size
    flush = None if args.no_flush else torch.empty(256 << 20, dtype=torch.uint8, device="cuda")

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def tick():
        if flush is not None:
            flush.fill_(1)
            torch.cuda.synchronize()
        eng.step(1)
        return eng.timing(), eng.kernel_stats()

    for _ in range(warmup):
        tick()
    barrier()
    t_wall0 = time.perf_counter()
    ms = upd = launches = rounds = 0.0
    per_tick = []
    kstats = {}
    for _ in range(steps):
        t, ks = tick()
        ms += t["ms_total"]; upd += t["agent_updates"]; launches += t["launches"]; rounds += t["rounds"]
        per_tick.append(t["ms_total"])
        for k, v in ks.items():
            a = kstats.setdefault(k, dict(ms=0.0, launches=0, units=0))
            a["ms"] += v["ms"]; a["launches"] += v["launches"]; a["units"] += v["units"]
    barrier()
    t_wall1 = time.perf_counter()

    tt = torch.tensor([ms, upd, launches, float(np.median(per_tick))], dtype=torch.float64, device="cuda")
    if world > 1:
        mx = tt.clone(); dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        sm = tt.clone(); dist.all_reduce(sm, op=dist.ReduceOp.SUM)
        ms_max, upd_all, launches_all, med_max = float(mx[0]), float(sm[1]), float(sm[2]), float(mx[3])
    else:
        ms_max, upd_all, launches_all, med_max = ms, upd, launches, float(np.median(per_tick))
    value = upd_all / (ms_max * 1e-3)

    # ---- N > 1: the strips' state against a single-handle run of the same global world (rank 0's GPU)
    verify_rec = None
    if verify and world > 1:
        ticks_done = warmup + steps
        digest, n_mine = state_digest(eng, size, strips.GHOST)
        got = [None] * world
        dist.all_gather_object(got, (digest, n_mine))
        if rank == 0:
            try:
                t0 = time.perf_counter()
                if device_world:
                    _, ref, _ = load_device_world(lib, workload, nx, size * world, agents * world, 0, 1, local, stationary=args.stationary)
                else:
                    parts = [make_world(workload, size, agents, seed=321 + r, nx=nx, rank=r, world=world) for r in range(world)]
                    cfg1 = dict(parts[0][0]); cfg1.update(ny=size * world, seed=321, device=local)
                    if args.stationary:
                        cfg1.update(reproduce=0, movement_cost=0.0)
                    ref = lib.create(lib.default_config(**cfg1))
                    cat = lambda k: np.concatenate([p[1][k] for p in parts], 0)
                    gh = strips.GHOST
                    wm = el = None
                    if parts[0][1].get("walkmap") is not None:
                        wm = np.concatenate([p[1]["walkmap"][gh:-gh] for p in parts], 0); el = np.concatenate([p[1]["elevation"][gh:-gh] for p in parts], 0)
                    ref.set_grid(cat("fully_grown"), cat("countdown"), wm, el)
                    ref.set_agents(np.concatenate([p[1]["ids"] + r * agents for r, p in enumerate(parts)]), cat("x"),
                                   np.concatenate([p[1]["y"] + size * r for r, p in enumerate(parts)]), cat("energy"), agents * world + 1)
                    del parts
                ref.step(ticks_done)
                ids, xy, en = ref.get_agents_packed()
                g = ref.get_grass_packed()
                ref.close()
                y = (xy >> 16).astype(np.int64)
                ok, detail = True, []
                for r in range(world):
                    m = (y > size * r) & (y <= size * (r + 1))
                    h = hashlib.sha256()
                    for a in (ids[m], xy[m], en[m].view(np.uint64), np.ascontiguousarray(g[size * r: size * (r + 1)])):
                        h.update(np.ascontiguousarray(a).tobytes())
                    same = h.hexdigest() == got[r][0]
                    ok = ok and same
                    detail.append(dict(rank=r, agents=got[r][1], match=same))
                verify_rec = {"status": "ok" if ok else "MISMATCH", "ticks": ticks_done,
                              "against": "single-handle run of the same global world on rank 0's GPU (sha256 of ids, positions, energy bits, grass per strip)",
                              "strips": detail, "seconds": time.perf_counter() - t0}
            except Exception as exc:  # noqa: BLE001
                verify_rec = {"status": "error", "error": str(exc)[:300]}
        barrier()

    # ---- end to end through the C ABI with pinned HOST buffers: what a host that keeps the model on its side pays per tick.
    # The narrow transfer forms of include/abm.h (1 grass byte per cell, 16 bytes per agent); the terrain (walkmap,
    # elevation) is a static model property and was uploaded once by abm_set_grid.
    def pinned_empty(n, dtype):
        t = torch.empty(int(n), dtype=dtype).pin_memory()
        return t.numpy(), t
    e2e_upd = h2d = d2h = 0
    e2e_dt = 0.0
    if e2e_steps > 0:
        keep = []
        Gl = G if world == 1 else eng.ny * nx  # grid arrays of a strip carry 2 x 32 ghost rows
        n_now = eng.count_agents()
        cap = int(n_now * 1.25) + 4096  # room for the offspring of the timed ticks
        host = {}
        for k, dt in (("ids", torch.int32), ("xy", torch.int32), ("energy", torch.float64)):
            host[k], t = pinned_empty(cap, dt); keep.append(t)
        host["ids"] = host["ids"].view(np.uint32); host["xy"] = host["xy"].view(np.uint32)
        host["grass"], t = pinned_empty(Gl, torch.uint8); keep.append(t)
        host["grass"] = host["grass"].reshape(-1, nx)
        out_agents = (host["ids"], host["xy"], host["energy"])
        # the model as the host sees it after the timed ticks (pinned buffers, as a Julia host would hold its arrays)
        ids, xy, en = eng.get_agents_packed(out_agents)
        eng.get_grass_packed(host["grass"])
        st = eng.get_global_state()
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            n = ids.size
            eng.set_grass_packed(host["grass"])
            eng.set_agents_packed(host["ids"][:n], host["xy"][:n], host["energy"][:n], st["next_id"])
            h2d += Gl + n * 16
            eng.step(1)
            e2e_upd += n
            ids, xy, en = eng.get_agents_packed(out_agents)
            eng.get_grass_packed(host["grass"])
            st = eng.get_global_state()
            d2h += ids.size * 16 + Gl
        barrier()
        e2e_dt = time.perf_counter() - t0
    clocks = sampler.finish(t_wall0, t_wall1) if sampler is not None else None
    e2 = torch.tensor([e2e_dt, e2e_upd], dtype=torch.float64, device="cuda")
    if world > 1:
        mx = e2.clone(); dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        sm = e2.clone(); dist.all_reduce(sm, op=dist.ReduceOp.SUM)
        e2e_value = float(sm[1]) / max(float(mx[0]), 1e-12)
    else:
        e2e_value = e2e_upd / max(e2e_dt, 1e-12)
    eng.close()
    del flush
    torch.cuda.empty_cache()
    if rank != 0:
        return None

    peak, peak_src = measured_peak()
    cell_b = C_BYTES_GREG if wl["walk_type"] == 2 else C_BYTES
    dom = max(kstats, key=lambda k: kstats[k]["ms"]) if kstats else None
    roof = None
    if dom:
        d = kstats[dom]
        per_cell = dom in ("regrow", "cell_scan", "memset_cell_counts")
        fused = dom == "stream_tick"  # one kernel does the agents AND the cells of the tick
        # the agent kernels jointly move A bytes per agent-update, the cell sweep C bytes per cell; the dominant kernel
        # is charged the whole per-unit figure of its kind (the other kernels of the tick are overhead on top)
        alg = (cell_b * G * steps) if per_cell else (A_BYTES * upd + (cell_b * G * steps if fused else 0.0))
        ach = alg / (d["ms"] * 1e-3) / 1e9
        roof = {"bound": "hbm", "kernel": dom, "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak, "traffic": traffic_for(workload, dom),
                "peak_source": peak_src, "launches_per_tick": d["launches"] / steps, "avg_launch_us": 1e3 * d["ms"] / max(d["launches"], 1),
                "algorithmic_bytes_per_launch": alg / max(d["launches"], 1),
                "algorithmic_bytes": "A*N (36 B per agent-update)" + (" + C*G (2 B per cell): this kernel is the whole tick" if fused else "")}
    tick_bytes = A_BYTES * upd / steps + cell_b * G
    tick_ach = tick_bytes / (ms / steps * 1e-3) / 1e9
    out = {
        "metric": "agent_updates_per_s", "value": value, "unit": "agent-updates/s", "n_gpus": world, "steps": steps, "warmup": warmup,
        "ms_per_step": ms_max / steps, "ms_per_step_median": med_max, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "int32 cells / f64 energy", "data": "synthetic",
        "config": {"workload": wl["desc"].format(n=nx, m=size * world, a=agents * world) + (" [stationary: reproduce=false, movement_cost=0]" if args.stationary else ""),
                   "grid": [nx, size * world], "rows_per_gpu": size, "agents_at_tick0": agents * world, "agents_live_mean": upd_all / steps, "ticks": steps,
                   "world_built_by": "abm_init_random on the device (terrain walkmap from a GPU value-noise generator)" if device_world else "numpy (abm_b200.worlds), uploaded through abm_set_grid / abm_set_agents",
                   "world_build_s": t_build,
                   "l2": "flushed between ticks (256 MiB device write)" if not args.no_flush else "not flushed",
                   "parallelism": "single GPU" if world == 1 else f"{world} row strips of a {nx}x{size * world} grid, one per GPU, exchange inside abm_step over " + ("NVLink peer memory (kernel stores into the neighbours' mailboxes)" if transport == "p2p" else "NCCL"),
                   "cells_per_s": G * world * steps / (ms_max * 1e-3), "resolution_rounds_per_tick": rounds / steps},
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": "agent-updates/s", "h2d_bytes_per_step": h2d // max(e2e_steps, 1), "d2h_bytes_per_step": d2h // max(e2e_steps, 1),
                "steps": e2e_steps, "ms_per_step": 1e3 * e2e_dt / max(e2e_steps, 1),
                "api": "abm_set_grass_packed + abm_set_agents_packed, abm_step(1), abm_get_agents_packed + abm_get_grass_packed (pinned host buffers; terrain uploaded once)"},
        "gpu_launches": int(launches_all),
        "roofline": roof,
        "roofline_tick": {"bound": "hbm", "achieved": tick_ach, "peak": peak, "unit": "GB/s", "frac": tick_ach / peak,
                          "algorithmic_bytes_per_tick": tick_bytes, "note": "whole tick (all kernels): A*N + C*G over the tick's device time"},
        "kernels": {k: dict(ms_per_tick=v["ms"] / steps, launches_per_tick=v["launches"] / steps) for k, v in sorted(kstats.items(), key=lambda kv: -kv[1]["ms"])},
        "wall_s_timed_region": t_wall1 - t_wall0,
    }
    if walk_frac is not None:
        out["config"]["walkable_fraction"] = walk_frac
    if verify_rec is not None:
        out["verify"] = verify_rec
    if e2e_steps <= 0:
        del out["e2e"]
    return out


def slim(rec):
    """A sub-record: the figures, without the parts that only make sense on the top-level line."""
    if rec is None:
        return None
    keep = ("value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "ms_per_step_median", "config", "roofline", "roofline_tick", "kernels", "gpu_launches", "verify", "e2e")
    return {k: rec[k] for k in keep if k in rec}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="C2", choices=sorted(WORKLOADS) + ["C4"])
    ap.add_argument("--transport", default=os.environ.get("ABM_TRANSPORT", "auto"), choices=["auto", "nccl", "p2p"],
                    help="multi-GPU halo transport: NVLink peer-memory mailboxes, NCCL, or auto (p2p if every rank can map its peers, else NCCL)")
    ap.add_argument("--replans", type=int, default=1 << 14)
    ap.add_argument("--ref-replans", type=int, default=64)
    ap.add_argument("--max-dist", type=int, default=0)
    ap.add_argument("--size", type=int, default=0, help="rows per GPU; default = the BASELINE.json size of the workload (C2 4096, C3 8192, C4 2048, C5 4096 rows)")
    ap.add_argument("--size-x", type=int, default=0, help="grid width when it differs from --size (C5: 32768)")
    ap.add_argument("--ref-size", type=int, default=0, help="rows of the CPU arm's world (default: the workload's own size = same configuration)")
    ap.add_argument("--ref-nx", type=int, default=0)
    ap.add_argument("--ref-budget", type=float, default=120.0, help="--impl reference: stop timing ticks after this many seconds (at least two ticks are timed)")
    ap.add_argument("--cpu-ticks", type=int, default=2, help="timed ticks of the cpu_baseline leg of the b200 arm (same configuration, after 1 warm-up tick)")
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--scheduler", default="by_id", choices=["by_id", "randomly"], help="Agents.Schedulers.by_id (the north star's fixed order; default) or Schedulers.randomly (what the reference scripts select: a Philox-keyed order per tick, include/abm.h)")
    ap.add_argument("--stationary", action="store_true", help="reproduce=false and movement_cost=0: nobody is born, nobody dies — the population of tick 0 stays (BASELINE.md §3)")
    ap.add_argument("--device-world", action="store_true", help="build the world on the device (abm_init_random) instead of numpy (always for C5 and for the sub-records)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-flush", action="store_true", help="do not flush L2 between timed ticks")
    ap.add_argument("--no-kernel-events", action="store_true", help="do not bracket every kernel with CUDA events (no per-kernel table, no roofline): shows what the events cost")
    ap.add_argument("--no-sub", action="store_true", help="skip the configs / c5 sub-records")
    ap.add_argument("--no-verify", action="store_true", help="N > 1: skip the comparison with a single-handle run")
    args = ap.parse_args()
    dnx, dsize = DEFAULT_SIZE[args.workload]
    if args.size <= 0:
        args.size = dsize
    if args.size_x <= 0:
        args.size_x = dnx if args.size == dsize else args.size
    if args.ref_size <= 0:
        args.ref_size, args.ref_nx = args.size, args.size_x
    elif args.ref_nx <= 0:
        args.ref_nx = args.ref_size

    if args.scheduler == "randomly":
        for wl in WORKLOADS.values():
            wl["cfg"]["scheduler"] = 1
            wl["desc"] += " [scheduler = Schedulers.randomly]"
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))

    if args.workload == "C4":
        if rank == 0:
            run_astar(args)
        return
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    from abm_b200 import build, capi

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: libabm has no CPU path (use --impl reference for the CPU baseline)")
    torch.cuda.set_device(local)
    numa = bind_to_gpu_numa_node(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    lib = capi.CLib(build.LIBABM if os.path.exists(build.LIBABM) else build.build_libabm(), "abm_")
    sampler = ClockSampler(local)
    sampler.start()
    nx, size = args.size_x, args.size
    total_agents = nx * size // 4 * world
    out = measure(lib, args, args.workload, nx, size, args.steps, args.warmup, args.e2e_steps, rank, world, local, dist, sampler=sampler,
                  device_world=args.device_world or args.workload == "C5", verify=not args.no_verify and total_agents <= (1 << 26),
                  kernel_events=not args.no_kernel_events)
    if rank == 0:
        out["config"]["pinned_host_numa_node"] = numa
    if not args.no_sub:
        # driver-run records of the other configurations (short: W = 3, K = 5 ticks; worlds built on the device)
        if world == 1 and args.workload != "C3":
            try:
                c3 = slim(measure(lib, args, "C3", 8192, 8192, 5, 3, 1, rank, world, local, dist, device_world=True))
            except Exception as exc:  # noqa: BLE001 -- a sub-record must never cost the headline line
                c3 = {"error": str(exc)[:300]}
            if rank == 0:
                out["configs"] = {"C3": c3}
        if args.workload != "C5":
            try:
                c5 = slim(measure(lib, args, "C5", 32768, 4096, 5, 3, 0, rank, world, local, dist, device_world=True, verify=False))
            except Exception as exc:  # noqa: BLE001
                c5 = {"error": str(exc)[:300]}
            if rank == 0:
                out["c5"] = c5
    if rank == 0:
        if world == 1 and not args.no_cpu_baseline:
            v, dt, u = time_oracle(args.workload, nx, size, nx * size // 4, ticks=max(1, args.cpu_ticks), stationary=args.stationary)
            out["cpu_baseline"] = {"value": v, "unit": "agent-updates/s", "cores": 1, "kind": "port",
                                   "sample": f"the same configuration ({nx}x{size} grid, {nx * size // 4} sheep at tick 0): the first {max(1, args.cpu_ticks)} ticks, {u} agent-updates in {dt:.1f} s",
                                   "note": "oracle/ C++ restatement, sequential like Agents.jl step!; Julia not installed on the box"}
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
