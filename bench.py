#!/usr/bin/env python
"""bench.py — headline measurement of the per-tick stepping path (BASELINE.json metric: agent-updates/s).

Workload (BASELINE.json configs[1], "C2"): v0.3 rules — RANDOM_WALK_GREGARIOUS, visual_distance = 3,
prob_random_walk = 0.9, eat + reproduce + grass regrowth — on a periodic 4096 x 4096 GridSpace with 2^22 sheep,
synthetic seeded world (abm_b200.worlds).  A "step" is one tick: every agent's agent_step! in id order, then
model_step!.  At N > 1 the GridSpace is 4096 x (4096 N) with 2^22 N sheep, cut into N row strips (one per GPU, weak
scaling); abm_step exchanges the boundary sub-tile rows with the ring neighbours every tick — over NVLink peer memory
(kernels store into the neighbours' mailboxes; `--transport auto`, the default, falls back to NCCL when a rank cannot map
its peers).  Other workloads: `--workload C3` (v0.5 walkmap rules, 8192^2), `C5` (the same rules on 32768 x 4096 rows per
GPU: 32768^2 and 2^28 sheep at 8 GPUs), `C1`, and `C4` (batched A* replans; prints its own line).

  value      agent-updates/s with the world resident in HBM; per-tick CUDA-event time on the library's stream,
             L2 flushed (256 MiB write) between ticks, max over ranks.
  e2e        the same metric through the C ABI with HOST buffers: per tick abm_set_grid + abm_set_agents (pinned
             host -> device), abm_step(1), abm_get_agents + abm_get_grid (device -> pinned host).
  roofline   dominant kernel: algorithmic bytes (A = 36 B per agent-update, C = 2 B per cell-update,
             +1 B per cell for the gregarious occupancy read) / its CUDA-event time, against MEASURED_PEAKS.json.
  cpu_baseline / --impl reference: the CPU restatement of the reference (oracle/, sequential like Agents.jl's
             step!) on a bounded sample of the same workload.  The oracle is only the timed baseline here.
"""
import argparse
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
               desc="v0.3 RANDOM_WALK_GREGARIOUS r=3, periodic {n}x{n} grid, {a} sheep (BASELINE.json configs[1])"),
    "C3": dict(walk_type=4, terrain=True, cfg=dict(walkmap_regrowth=1, reproduction_prob=0.004),
               desc="v0.5 RANDOM_WALKMAP + walkmap regrowth, periodic {n}x{n} terrain, {a} sheep (BASELINE.json configs[2])"),
    "C5": dict(walk_type=4, terrain=True, cfg=dict(walkmap_regrowth=1, reproduction_prob=0.004),
               desc="v0.5 RANDOM_WALKMAP + walkmap regrowth, periodic {n}x{m} terrain, {a} sheep (BASELINE.json configs[4] per-GPU share: 32768x4096 rows of 32768x(4096 N))"),
    "C1": dict(walk_type=0, terrain=False, cfg=dict(reproduction_prob=0.001),
               desc="v0.1 RANDOM_WALK, {n}x{n} grid, {a} sheep"),
}


def make_world(workload, size, agents, seed=321, nx=None, rank=0, world=1):
    """One rank's share: `size` rows of an nx-wide world of size * world rows (nx = size, world = 1: the square
    single-GPU worlds).  With world > 1 the terrain arrays carry the 32 ghost rows above and below (true neighbours'
    terrain, the y wrap included); grass and agents cover the own rows only."""
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
        el = worlds.make_terrain(size, size, seed=seed + 1)
        wm = worlds.walkmap_from_elevation(el)
    w = worlds.make_world(size, size, agents, seed=seed, walkmap=wm, periodic=True)
    w["walkmap"], w["elevation"] = wm, el
    cfg = dict(nx=size, ny=size, periodic=1, walk_type=wl["walk_type"], seed=seed)
    cfg.update(wl["cfg"])
    return cfg, w


def load(lib, cfg, w, device=0):
    e = lib.create(lib.default_config(device=device, **cfg))
    e.set_grid(w["fully_grown"], w["countdown"], w.get("walkmap"), w.get("elevation"))
    e.set_agents(w["ids"], w["x"], w["y"], w["energy"])
    return e


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


def oracle_lib():
    from abm_b200 import build, capi
    return capi.CLib(build.build_oracle(), "oracle_")


def time_oracle(workload, size, agents, ticks, warm=1):
    """agent-updates/s of the CPU restatement (sequential, 1 thread — Agents.jl's step! is serial)."""
    lib = oracle_lib()
    cfg, w = make_world(workload, size, agents)
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
    import torch
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
    if rank != 0:
        return
    size, agents = args.ref_size, args.ref_size * args.ref_size // 4
    wl = WORKLOADS[args.workload]
    # one "step" = one tick of the bounded sample; W warm-up ticks, K timed ticks
    lib = oracle_lib()
    cfg, w = make_world(args.workload, size, agents)
    e = load(lib, cfg, w)
    e.step(args.warmup)
    upd, t0 = 0, time.perf_counter()
    for _ in range(args.steps):
        upd += e.count_agents()
        e.step(1)
    dt = time.perf_counter() - t0
    v = upd / dt
    sample = f"{size}x{size} grid, {agents} sheep, same rules/density as the {args.size}x{args.size} workload, {args.steps} ticks"
    print(json.dumps({
        "impl": "reference", "metric": "agent_updates_per_s", "value": v, "unit": "agent-updates/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "int32 cells / f64 energy", "data": "synthetic",
        "config": {"workload": wl["desc"].format(n=args.size_x or args.size, m=args.size, a=(args.size_x or args.size) * args.size // 4), "sample": sample},
        "cpu_baseline": {"value": v, "unit": "agent-updates/s", "cores": 1, "kind": "port", "sample": sample,
                         "note": "C++ restatement of the reference rules + Agents.jl semantics (oracle/); Julia is not installed; Agents.jl step! is serial, hence 1 thread"},
        "e2e": {"value": v, "unit": "agent-updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


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
    ap.add_argument("--size", type=int, default=0, help="grid side (rows per GPU); default = the BASELINE.json size of the workload (C2 4096, C3 8192, C4 2048, C5 4096 rows)")
    ap.add_argument("--size-x", type=int, default=0, help="grid width when it differs from --size (C5: 32768)")
    ap.add_argument("--ref-size", type=int, default=1024)
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-flush", action="store_true", help="do not flush L2 between timed ticks")
    ap.add_argument("--no-kernel-events", action="store_true", help="do not bracket every kernel with CUDA events (no per-kernel table, no roofline): shows what the events cost")
    args = ap.parse_args()
    if args.size <= 0:
        args.size = {"C3": 8192, "C4": 2048}.get(args.workload, 4096)
    if args.size_x <= 0 and args.workload == "C5":
        args.size_x = 32768

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
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    lib = capi.CLib(build.LIBABM if os.path.exists(build.LIBABM) else build.build_libabm(), "abm_")
    size = args.size                      # rows per GPU
    nx = args.size_x or size              # width of the world
    agents = nx * size // 4
    wl = WORKLOADS[args.workload]
    cfg, w = make_world(args.workload, size, agents, seed=321 + rank, nx=nx, rank=rank, world=world)
    if world > 1:
        # strip `rank` of an nx x (size * world) GridSpace: this rank's rows and sheep (global ids and y), ghost rows for the grids
        from abm_b200 import strips
        cfg.update(ny=size * world, strip_y0=size * rank, strip_ny=size, n_ranks=world, rank=rank, seed=321, device=local)
        eng = lib.create(lib.default_config(**cfg))
        pad = lambda a: np.concatenate([np.zeros((strips.GHOST, nx), a.dtype), a, np.zeros((strips.GHOST, nx), a.dtype)], 0)
        eng.ny = size + 2 * strips.GHOST
        eng.set_grid(pad(w["fully_grown"]), pad(w["countdown"]), w["walkmap"], w["elevation"])  # terrain comes with its ghost rows
        eng.set_agents(w["ids"] + rank * agents, w["x"], w["y"] + size * rank, w["energy"], agents * world + 1)
        if args.transport == "auto":
            # peer-memory mailboxes need CUDA IPC + peer access between all ranks of the node; every rank must agree on the choice
            try:
                strips.init_p2p(eng)
                ok = 1
            except Exception as exc:  # noqa: BLE001 -- any failure to map a peer means NCCL for everybody
                print(f"[bench rank {rank}] peer-memory transport unavailable ({exc}); trying NCCL", file=sys.stderr)
                ok = 0
            flag = torch.tensor([ok], device="cuda")
            dist.all_reduce(flag, op=dist.ReduceOp.MIN)
            args.transport = "p2p" if int(flag.item()) else "nccl"
            if args.transport == "nccl":
                strips.init_nccl(eng)
        else:
            (strips.init_p2p if args.transport == "p2p" else strips.init_nccl)(eng)
    else:
        eng = load(lib, cfg, w, device=local)
    eng.set_profiling(not args.no_kernel_events)
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

    sampler = ClockSampler(local)
    sampler.start()
    for _ in range(args.warmup):
        tick()
    barrier()
    t_wall0 = time.perf_counter()
    ms = upd = launches = rounds = 0.0
    kstats = {}
    for _ in range(args.steps):
        t, ks = tick()
        ms += t["ms_total"]; upd += t["agent_updates"]; launches += t["launches"]; rounds += t["rounds"]
        for k, v in ks.items():
            a = kstats.setdefault(k, dict(ms=0.0, launches=0, units=0))
            a["ms"] += v["ms"]; a["launches"] += v["launches"]; a["units"] += v["units"]
    barrier()
    t_wall1 = time.perf_counter()
    t_wall = t_wall1 - t_wall0

    tt = torch.tensor([ms, upd, launches], dtype=torch.float64, device="cuda")
    if world > 1:
        mx = tt.clone(); dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        sm = tt.clone(); dist.all_reduce(sm, op=dist.ReduceOp.SUM)
        ms_max, upd_all, launches_all = float(mx[0]), float(sm[1]), float(sm[2])
    else:
        ms_max, upd_all, launches_all = ms, upd, launches
    value = upd_all / (ms_max * 1e-3)

    # ---- end to end through the C ABI with pinned HOST buffers: what a host that keeps the model on its side pays per tick.
    # The narrow transfer forms of include/abm.h (1 grass byte per cell, 16 bytes per agent); the terrain (walkmap,
    # elevation) is a static model property and was uploaded once by abm_set_grid.
    def pinned_empty(n, dtype):
        t = torch.empty(int(n), dtype=dtype).pin_memory()
        return t.numpy(), t
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
    # the model as the host sees it after the timed ticks (pinned buffers, as a Julia host would hold its arrays)
    out_agents = (host["ids"], host["xy"], host["energy"])
    e2e_upd = h2d = d2h = 0
    e2e_dt = 0.0
    if args.e2e_steps > 0:
        ids, xy, en = eng.get_agents_packed(out_agents)
        eng.get_grass_packed(host["grass"])
        st = eng.get_global_state()
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.e2e_steps):
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
    clocks = sampler.finish(t_wall0, t_wall1)
    e2 = torch.tensor([e2e_dt, e2e_upd], dtype=torch.float64, device="cuda")
    if world > 1:
        mx = e2.clone(); dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        sm = e2.clone(); dist.all_reduce(sm, op=dist.ReduceOp.SUM)
        e2e_value = float(sm[1]) / max(float(mx[0]), 1e-12)
    else:
        e2e_value = e2e_upd / max(e2e_dt, 1e-12)

    if rank == 0:
        peak, peak_src = measured_peak()
        cell_b = C_BYTES_GREG if wl["walk_type"] == 2 else C_BYTES
        dom = max(kstats, key=lambda k: kstats[k]["ms"]) if kstats else None
        roof = None
        if dom:
            d = kstats[dom]
            per_cell = dom in ("regrow", "cell_scan", "memset_cell_counts")
            # the agent kernels jointly move A bytes per agent-update, the cell sweep C bytes per cell; the dominant kernel
            # is charged the whole per-unit figure of its kind (the other kernels of the tick are overhead on top)
            alg = (cell_b * G * args.steps) if per_cell else (A_BYTES * upd)
            ach = alg / (d["ms"] * 1e-3) / 1e9
            traffic = None
            tp = os.path.join(ROOT, "profiles", "traffic.json")
            if os.path.exists(tp):
                traffic = json.load(open(tp)).get(dom)
            roof = {"bound": "hbm", "kernel": dom, "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak, "traffic": traffic,
                    "peak_source": peak_src, "launches_per_tick": d["launches"] / args.steps, "avg_launch_us": 1e3 * d["ms"] / max(d["launches"], 1),
                    "algorithmic_bytes_per_tick": alg / args.steps}
        tick_bytes = A_BYTES * upd / args.steps + cell_b * G
        tick_ach = tick_bytes / (ms / args.steps * 1e-3) / 1e9
        out = {
            "metric": "agent_updates_per_s", "value": value, "unit": "agent-updates/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_max / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "int32 cells / f64 energy", "data": "synthetic",
            "config": {"workload": wl["desc"].format(n=nx, m=size, a=agents), "grid": [nx, size], "agents": agents, "ticks": args.steps,
                       "l2": "flushed between ticks (256 MiB device write)" if flush is not None else "not flushed",
                       "parallelism": "single GPU" if world == 1 else f"{world} row strips of a {nx}x{size * world} grid, one per GPU, halo exchange inside abm_step over " + ("NVLink peer memory (kernel stores into the neighbours' mailboxes)" if args.transport == "p2p" else "NCCL"),
                       "cells_per_s": G * world * args.steps / (ms_max * 1e-3), "resolution_rounds_per_tick": rounds / args.steps},
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": "agent-updates/s", "h2d_bytes_per_step": h2d // max(args.e2e_steps, 1), "d2h_bytes_per_step": d2h // max(args.e2e_steps, 1),
                    "steps": args.e2e_steps, "ms_per_step": 1e3 * e2e_dt / max(args.e2e_steps, 1),
                    "api": "abm_set_grass_packed + abm_set_agents_packed, abm_step(1), abm_get_agents_packed + abm_get_grass_packed (pinned host buffers; terrain uploaded once)"},
            "gpu_launches": int(launches_all),
            "roofline": roof,
            "roofline_tick": {"bound": "hbm", "achieved": tick_ach, "peak": peak, "unit": "GB/s", "frac": tick_ach / peak,
                              "algorithmic_bytes_per_tick": tick_bytes, "note": "whole tick (all kernels): A*N + C*G over the tick's device time"},
            "kernels": {k: dict(ms_per_tick=v["ms"] / args.steps, launches_per_tick=v["launches"] / args.steps) for k, v in sorted(kstats.items(), key=lambda kv: -kv[1]["ms"])},
            "wall_s_timed_region": t_wall,
        }
        if world == 1 and not args.no_cpu_baseline:
            rs = args.ref_size
            v, dt, u = time_oracle(args.workload, rs, rs * rs // 4, ticks=max(2, int(12e6 / (rs * rs // 4))))
            out["cpu_baseline"] = {"value": v, "unit": "agent-updates/s", "cores": 1, "kind": "port",
                                   "sample": f"{rs}x{rs} grid, {rs * rs // 4} sheep, same rules and density, {u} agent-updates in {dt:.1f} s",
                                   "note": "oracle/ C++ restatement, sequential like Agents.jl step!; Julia not installed on the box"}
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
