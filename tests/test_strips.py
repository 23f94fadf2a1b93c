"""Row strips: the strip decomposition must reproduce the single-handle / oracle trajectory bit for bit.

  * one strip that is its own ring neighbour (exercises pack / unpack / ghost rows / merge without processes);
  * world_size-2 and -3 rings of processes over gloo (CPU, emulated engine, host-memory transport callbacks);
  * on the GPU box: 2 strips with the NCCL transport (needs 2 GPUs; skipped otherwise).
"""
import os
import sys

import numpy as np
import pytest

from conftest import ROOT, assert_same_state, build, capi, load_case, make_case
from abm_b200 import strips

CASES = {
    "rw": dict(nx=64, ny=128, n=3000, wt=0, per=1, kw=dict(reproduction_prob=0.05, regrowth_time=3, delta_energy=3.0)),
    "rw_clamped": dict(nx=96, ny=96, n=3000, wt=0, per=0, kw=dict(reproduction_prob=0.05, regrowth_time=3, delta_energy=3.0)),
    "attractor": dict(nx=64, ny=192, n=3000, wt=1, per=0, kw=dict(reproduction_prob=0.05, regrowth_time=3, delta_energy=3.0, prob_random_walk=0.5)),
    "gregarious": dict(nx=64, ny=128, n=2500, wt=2, per=1, kw=dict(visual_distance=3, reproduction_prob=0.05, regrowth_time=3, delta_energy=3.0)),
    "gregarious_r7": dict(nx=64, ny=96, n=900, wt=2, per=1, kw=dict(visual_distance=7, prob_random_walk=0.2, reproduction_prob=0.05, regrowth_time=2, delta_energy=3.0)),
    "walkmap": dict(nx=96, ny=128, n=3000, wt=4, per=1, terrain=True, kw=dict(walkmap_regrowth=1, reproduction_prob=0.05, regrowth_time=3, delta_energy=3.0)),
    # Schedulers.randomly: order keys instead of ids in everything the strips exchange for ORDER (ghost summaries, parents)
    "rw_randomly": dict(nx=64, ny=128, n=3000, wt=0, per=1, kw=dict(scheduler=1, reproduction_prob=0.05, regrowth_time=3, delta_energy=3.0)),
    "gregarious_randomly": dict(nx=64, ny=128, n=2500, wt=2, per=1, kw=dict(scheduler=1, visual_distance=3, reproduction_prob=0.05, regrowth_time=3, delta_energy=3.0)),
    "walkmap_randomly": dict(nx=96, ny=128, n=3000, wt=4, per=1, terrain=True, kw=dict(scheduler=1, walkmap_regrowth=1, reproduction_prob=0.05, regrowth_time=3, delta_energy=3.0)),
}


def _fuzz_case(seed):
    """A random strip-able configuration, reproducible from the seed (the spawned workers rebuild it from the name)."""
    rng = np.random.default_rng(seed)
    nx, ny = [(64, 96), (96, 96), (64, 192), (128, 96)][rng.integers(4)]
    wt = int(rng.choice([0, 1, 2, 4]))
    per = 1 if wt == 2 else int(rng.integers(0, 2))
    n = max(1, int(nx * ny * [0.05, 0.3, 0.8, 1.5][rng.integers(4)]))
    kw = dict(reproduction_prob=float(rng.choice([0.0, 0.05, 0.3])), regrowth_time=int(rng.integers(1, 6)), delta_energy=float(rng.integers(1, 5)),
              prob_random_walk=float(rng.choice([0.0, 0.3, 0.9])), visual_distance=int(rng.integers(1, 8)))
    if wt == 4:
        kw["walkmap_regrowth"] = int(rng.integers(0, 2))
    return dict(nx=nx, ny=ny, n=n, wt=wt, per=per, terrain=wt == 4, kw=kw)


def _case(name):
    if name.startswith("crowd"):  # every sheep starts in the first 32 rows: the other strips begin (and may stay) empty
        cfgkw, w = _case("fuzz" + name[5:])
        w["y"] = 1 + (w["y"] - 1) % 32
        if w.get("walkmap") is not None:  # keep them on walkable cells of those rows
            ok = np.argwhere(w["walkmap"][:32] == 1)
            pick = ok[np.random.default_rng(5).integers(0, len(ok), w["y"].size)]
            w["y"], w["x"] = pick[:, 0] + 1, pick[:, 1] + 1
        return cfgkw, w
    c = _fuzz_case(int(name[4:])) if name.startswith("fuzz") else CASES[name]
    return make_case(c["nx"], c["ny"], c["n"], c["wt"], c["per"], terrain=c.get("terrain", False), seed=31, **c["kw"])


@pytest.mark.parametrize("name", ["rw", "gregarious", "walkmap", "gregarious_randomly", "walkmap_randomly"])
def test_single_strip_ring_of_one(emu, oracle, name):
    cfgkw, w = _case(name)
    a = load_case(oracle, cfgkw, w)
    b = strips.load_strip(emu, cfgkw, w, 1, 0)
    for t in range(10):
        a.step(1)
        b.step(1)
        assert_same_state(strips.merge_snapshots([strips.strip_snapshot(b)]), a.snapshot(), f"{name} tick {t}")


def _ring_worker(rank, world, name, ticks, port, out_dir, use_gpu, transport="nccl"):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch
    import torch.distributed as dist
    from conftest import build as b2, capi as c2
    if use_gpu:
        torch.cuda.set_device(rank)
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
        lib = c2.CLib(b2.LIBABM, "abm_")
    else:
        dist.init_process_group("gloo", rank=rank, world_size=world)
        lib = c2.CLib(b2.LIBEMU, "emu_")
    cfgkw, w = _case(name[5:] if name.startswith("init_") else name)
    if use_gpu:
        cfgkw = dict(cfgkw, device=rank)
    e = strips.load_strip(lib, cfgkw, w, world, rank)
    if name.startswith("init_"):  # the engine draws grass and sheep itself (abm_init_random on strips)
        if w.get("walkmap") is not None:
            e.set_global_walkmap(w["walkmap"])
        e.init_random(w["ids"].size, 99)
    if use_gpu and transport == "p2p":
        strips.init_p2p(e)
    elif use_gpu:
        strips.init_nccl(e)
    else:
        tr = strips.DistTransport()
        e.set_transport(tr.struct)
    snaps = []
    for t in range(ticks):
        e.step(1)
        snaps.append(strips.strip_snapshot(e))
        snaps[-1]["reduced"] = [e.reduce(k) for k in range(5)]  # collective: the values over the whole GridSpace
    np.save(os.path.join(out_dir, f"snap_{rank}.npy"), np.array(snaps, dtype=object), allow_pickle=True)
    dist.barrier()
    dist.destroy_process_group()


def _run_ring(world, name, ticks, tmp_path, use_gpu=False, transport="nccl"):
    import torch.multiprocessing as mp
    port = 29500 + (os.getpid() + world * 7 + len(name)) % 400
    mp.spawn(_ring_worker, args=(world, name, ticks, port, str(tmp_path), use_gpu, transport), nprocs=world, join=True)
    return [np.load(os.path.join(str(tmp_path), f"snap_{r}.npy"), allow_pickle=True) for r in range(world)]


@pytest.mark.parametrize("world,name", [(2, "rw"), (2, "gregarious"), (2, "rw_clamped"), (3, "attractor"), (2, "walkmap"), (3, "gregarious_r7"), (4, "walkmap"), (3, "fuzz11"), (2, "fuzz12"), (3, "crowd3"), (2, "crowd4"), (3, "crowd9"),
                                        (2, "rw_randomly"), (3, "gregarious_randomly"), (2, "walkmap_randomly")])
def test_ring_of_processes_over_gloo(oracle, tmp_path, world, name):
    build.build_emu()
    ticks = 8
    per_rank = _run_ring(world, name, ticks, tmp_path)
    cfgkw, w = _case(name)
    a = load_case(oracle, cfgkw, w)
    for t in range(ticks):
        a.step(1)
        assert_same_state(strips.merge_snapshots([per_rank[r][t] for r in range(world)]), a.snapshot(), f"{name} world {world} tick {t}")


def _check_init_and_reduce(oracle, per_rank, world, name, ticks):
    cfgkw, w = _case(name)
    a = load_case(oracle, cfgkw, w)
    a.init_random(w["ids"].size, 99)
    for t in range(ticks):
        a.step(1)
        assert_same_state(strips.merge_snapshots([per_rank[r][t] for r in range(world)]), a.snapshot(), f"init {name} world {world} tick {t}")
        want = [a.reduce(k) for k in range(5)]
        for r in range(world):
            assert per_rank[r][t]["reduced"] == want, (name, world, t, r)


@pytest.mark.parametrize("world,name", [(2, "rw"), (3, "walkmap"), (2, "gregarious")])
def test_init_random_and_reduce_on_strips(oracle, tmp_path, world, name):
    """abm_init_random on strip handles builds the world a single handle would build (every rank keeps the sheep that
    land on its rows), and abm_reduce returns the value over the whole GridSpace on every rank."""
    build.build_emu()
    ticks = 5
    per_rank = _run_ring(world, "init_" + name, ticks, tmp_path)
    _check_init_and_reduce(oracle, per_rank, world, name, ticks)


@pytest.mark.gpu
@pytest.mark.parametrize("world,name,transport", [(2, "walkmap", "p2p"), (2, "gregarious", "nccl")])
def test_init_random_and_reduce_on_gpu_strips(libabm, oracle, tmp_path, world, name, transport):
    import torch
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    ticks = 5
    per_rank = _run_ring(world, "init_" + name, ticks, tmp_path, use_gpu=True, transport=transport)
    _check_init_and_reduce(oracle, per_rank, world, name, ticks)


@pytest.mark.parametrize("world,name", [(1, "walkmap"), (3, "attractor")])
def test_offspring_ids_through_the_bitmap(emu, oracle, tmp_path, monkeypatch, world, name):
    """ABM_BIRTH_BITMAP=0: offspring ids of the strips are ranked through the id-space bitmap (the path taken when a tick
    has ~10^5 births or more) instead of the all-pairs rank kernel; spawned workers inherit the variable."""
    monkeypatch.setenv("ABM_BIRTH_BITMAP", "0")
    cfgkw, w = _case(name)
    a = load_case(oracle, cfgkw, w)
    ticks = 8
    if world == 1:
        b = strips.load_strip(emu, cfgkw, w, 1, 0)
        for t in range(ticks):
            a.step(1)
            b.step(1)
            assert_same_state(strips.merge_snapshots([strips.strip_snapshot(b)]), a.snapshot(), f"{name} tick {t}")
        return
    per_rank = _run_ring(world, name, ticks, tmp_path)
    for t in range(ticks):
        a.step(1)
        assert_same_state(strips.merge_snapshots([per_rank[r][t] for r in range(world)]), a.snapshot(), f"{name} world {world} tick {t}")


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["rw", "gregarious"])
def test_single_strip_on_the_gpu(libabm, oracle, name):
    cfgkw, w = _case(name)
    a = load_case(oracle, cfgkw, w)
    b = strips.load_strip(libabm, cfgkw, w, 1, 0)
    for t in range(10):
        a.step(1)
        b.step(1)
        assert_same_state(strips.merge_snapshots([strips.strip_snapshot(b)]), a.snapshot(), f"{name} tick {t}")


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["rw", "gregarious", "walkmap"])
def test_two_strips_over_nccl(libabm, oracle, tmp_path, name):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    ticks = 8
    per_rank = _run_ring(2, name, ticks, tmp_path, use_gpu=True)
    cfgkw, w = _case(name)
    a = load_case(oracle, cfgkw, w)
    for t in range(ticks):
        a.step(1)
        assert_same_state(strips.merge_snapshots([per_rank[r][t] for r in range(2)]), a.snapshot(), f"{name} tick {t}")


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["rw", "gregarious", "walkmap", "rw_clamped", "gregarious_randomly", "walkmap_randomly"])
def test_two_strips_over_peer_memory(libabm, oracle, tmp_path, name):
    """The NVLink mailbox transport (abm_p2p_export / abm_p2p_connect): kernels store into the neighbour's memory."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    ticks = 8
    per_rank = _run_ring(2, name, ticks, tmp_path, use_gpu=True, transport="p2p")
    cfgkw, w = _case(name)
    a = load_case(oracle, cfgkw, w)
    for t in range(ticks):
        a.step(1)
        assert_same_state(strips.merge_snapshots([per_rank[r][t] for r in range(2)]), a.snapshot(), f"{name} tick {t}")


@pytest.mark.gpu
def test_two_strips_offspring_bitmap_on_the_gpu(libabm, oracle, tmp_path, monkeypatch):
    """The bitmap ranking of offspring ids (ABM_BIRTH_BITMAP=0 forces it) over the peer-memory transport."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    monkeypatch.setenv("ABM_BIRTH_BITMAP", "0")
    ticks = 8
    per_rank = _run_ring(2, "walkmap", ticks, tmp_path, use_gpu=True, transport="p2p")
    cfgkw, w = _case("walkmap")
    a = load_case(oracle, cfgkw, w)
    for t in range(ticks):
        a.step(1)
        assert_same_state(strips.merge_snapshots([per_rank[r][t] for r in range(2)]), a.snapshot(), f"walkmap tick {t}")


@pytest.mark.gpu
@pytest.mark.parametrize("world,name,transport", [(4, "walkmap", "p2p"), (3, "gregarious_r7", "p2p"), (3, "attractor", "nccl")])
def test_rings_of_three_and_four_gpus(libabm, oracle, tmp_path, world, name, transport):
    """With three or more strips the two neighbours of a rank are different processes (at two they are the same one):
    the mailbox slots, the all-to-all counters and the offspring all-gather with more than one peer, against the oracle."""
    import torch
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    ticks = 6
    per_rank = _run_ring(world, name, ticks, tmp_path, use_gpu=True, transport=transport)
    cfgkw, w = _case(name)
    a = load_case(oracle, cfgkw, w)
    for t in range(ticks):
        a.step(1)
        assert_same_state(strips.merge_snapshots([per_rank[r][t] for r in range(world)]), a.snapshot(), f"{name} world {world} tick {t}")


# ------------------------------------------------------------------ single-process multi-device (abm_group)

def _check_group(lib, oracle, devices, names):
    """abm_group: one call fans out to one worker thread per strip; whole-GridSpace arrays in and out.  Against the oracle,
    tick by tick, including a world built by the engine itself and the reductions of run!."""
    for name in names:
        cfgkw, w = _case(name)
        a = load_case(oracle, cfgkw, w)
        g = lib.create_group(lib.default_config(**cfgkw), devices)
        g.set_grid(w["fully_grown"], w["countdown"], w.get("walkmap"), w.get("elevation"))
        g.set_agents(w["ids"], w["x"], w["y"], w["energy"])
        for t in range(6):
            a.step(1); g.step(1)
            assert_same_state(g.snapshot(), a.snapshot(), f"group {name} x{len(devices)} tick {t}")
        assert [g.reduce(k) for k in range(5)] == [a.reduce(k) for k in range(5)]
        g.init_random(w["ids"].size, 17); a.init_random(w["ids"].size, 17)
        a.step(3); g.step(3)  # several ticks per call
        assert_same_state(g.snapshot(), a.snapshot(), f"group {name} after init_random")
        g.close(); a.close()


@pytest.mark.parametrize("n", [1, 2, 3])
def test_group_of_strips_in_one_process_emulated(emu, oracle, n):
    _check_group(emu, oracle, [0] * n, ["walkmap", "gregarious", "rw_clamped"])


@pytest.mark.gpu
@pytest.mark.parametrize("n", [2, 3])
def test_group_of_strips_in_one_process_on_the_gpu(libabm, oracle, n):
    """One strip per device, one worker thread per strip, NVLink peer memory mapped inside the process."""
    import torch
    if torch.cuda.device_count() < n:
        pytest.skip(f"needs {n} GPUs (a strip per device)")
    _check_group(libabm, oracle, list(range(n)), ["walkmap", "gregarious", "attractor"])
