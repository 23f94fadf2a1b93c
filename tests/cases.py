"""Named parity cases: the reference scripts' configurations at sizes the CPU oracle finishes in seconds.

Each case: (config kwargs, world kwargs, snapshot ticks).  Script defaults: scripts/v0.1.jl:36-46 (80x80, 40 sheep,
regrowth 30, Δenergy 4, repro 0.001, cost 1, seed 321, non-periodic), v0.2 attractor = griddims ./ 2
(scripts/v0.2.jl:57), v0.3 periodic + visual_distance 5 (scripts/v0.3.jl:32-53), v0.4 counter 50 / repro 0.004
(scripts/v0.4.jl:47-55), v0.5 walkmap + walkmap regrowth (scripts/v0.5.jl:39-134).  BASELINE.json C1 = v0.1 at
100x100 / 100 sheep / 1000 ticks.  `dense_*` cases raise density and birth rate to stress the sequential
dependencies of SURVEY.md Appendix C.
"""
from conftest import make_case

RW, ATT, GREG, ALT, WMAP = 0, 1, 2, 3, 4

CASES = {
    # name: (nx, ny, n, walk_type, periodic, terrain, extra config, snapshot ticks)
    "v0.1_default": (80, 80, 40, RW, 0, False, dict(), (1, 10, 100)),
    "C1_v0.1_100x100": (100, 100, 100, RW, 0, False, dict(), (100, 1000)),
    "v0.2_attractor": (80, 80, 40, ATT, 0, False, dict(), (1, 10, 100)),
    "v0.3_gregarious": (80, 80, 40, GREG, 1, False, dict(visual_distance=5), (1, 10, 100)),
    "v0.4_alternated": (80, 80, 40, ALT, 1, True, dict(counter=50, reproduction_prob=0.004, astar_metric=0), (1, 10, 100)),
    "v0.5_walkmap": (80, 80, 40, WMAP, 1, True, dict(walkmap_regrowth=1, reproduction_prob=0.004), (1, 10, 100)),
    "v0.6_alternated_penalty": (80, 80, 40, ALT, 1, True, dict(counter=50, reproduction_prob=0.004, astar_metric=1, astar_penalty=1), (1, 10, 100)),
    "dense_rw": (48, 40, 900, RW, 1, False, dict(reproduction_prob=0.05, regrowth_time=3, delta_energy=3.0), (1, 5, 40)),
    "dense_rw_clamped": (37, 29, 500, RW, 0, False, dict(reproduction_prob=0.05, regrowth_time=3, delta_energy=3.0), (1, 5, 40)),
    "dense_attractor": (50, 45, 900, ATT, 0, False, dict(reproduction_prob=0.05, regrowth_time=3, delta_energy=3.0, prob_random_walk=0.5), (1, 5, 40)),
    "dense_gregarious_r3": (64, 40, 700, GREG, 1, False, dict(visual_distance=3, reproduction_prob=0.05, regrowth_time=3, delta_energy=3.0), (1, 5, 40)),
    "dense_gregarious_r1_p0": (33, 35, 400, GREG, 1, False, dict(visual_distance=1, prob_random_walk=0.0, reproduction_prob=0.05, regrowth_time=2, delta_energy=3.0), (1, 5, 40)),
    "dense_walkmap": (64, 64, 900, WMAP, 1, True, dict(walkmap_regrowth=1, reproduction_prob=0.05, regrowth_time=3, delta_energy=3.0), (1, 5, 40)),
    "dense_alternated": (40, 40, 300, ALT, 1, True, dict(counter=7, reproduction_prob=0.03, regrowth_time=3, delta_energy=3.0, astar_metric=1, astar_penalty=1), (1, 5, 40)),
    # grids made of whole 32x32 cores: the engine takes its tile path (abm_tile.h) here
    "tile_rw_96x64": (96, 64, 2500, RW, 1, False, dict(reproduction_prob=0.05, regrowth_time=3, delta_energy=3.0), (1, 5, 30)),
    "tile_rw_clamped_64x96": (64, 96, 2500, RW, 0, False, dict(reproduction_prob=0.05, regrowth_time=3, delta_energy=3.0), (1, 5, 30)),
    "tile_attractor_128x64": (128, 64, 3000, ATT, 0, False, dict(reproduction_prob=0.05, regrowth_time=3, delta_energy=3.0, prob_random_walk=0.5), (1, 5, 30)),
    "tile_gregarious_r3_96x96": (96, 96, 2500, GREG, 1, False, dict(visual_distance=3, reproduction_prob=0.05, regrowth_time=3, delta_energy=3.0), (1, 5, 30)),
    "tile_gregarious_r7_p0_64x64": (64, 64, 900, GREG, 1, False, dict(visual_distance=7, prob_random_walk=0.0, reproduction_prob=0.05, regrowth_time=2, delta_energy=3.0), (1, 5, 30)),
    "tile_gregarious_r1_dense_64x64": (64, 64, 3500, GREG, 1, False, dict(visual_distance=1, prob_random_walk=0.3, reproduction_prob=0.08, regrowth_time=2, delta_energy=3.0), (1, 5, 30)),
    "tile_walkmap_128x96": (128, 96, 3000, WMAP, 1, True, dict(walkmap_regrowth=1, reproduction_prob=0.05, regrowth_time=3, delta_energy=3.0), (1, 5, 30)),
    "tile_crowded_64x64": (64, 64, 12000, RW, 1, False, dict(reproduction_prob=0.1, regrowth_time=1, delta_energy=3.0), (1, 5, 20)),
    # Schedulers.randomly (the scheduler every reference script selects, scripts/v0.1.jl:63): the Philox-keyed order of include/abm.h
    "randomly_v0.1_default": (80, 80, 40, RW, 0, False, dict(scheduler=1), (1, 10, 100)),
    "randomly_dense_rw_clamped": (37, 29, 500, RW, 0, False, dict(scheduler=1, reproduction_prob=0.05, regrowth_time=3, delta_energy=3.0), (1, 5, 40)),
    "randomly_dense_alternated": (40, 40, 300, ALT, 1, True, dict(scheduler=1, counter=7, reproduction_prob=0.03, regrowth_time=3, delta_energy=3.0, astar_metric=1, astar_penalty=1), (1, 5, 40)),
    "randomly_tile_gregarious_r3_96x96": (96, 96, 2500, GREG, 1, False, dict(scheduler=1, visual_distance=3, reproduction_prob=0.05, regrowth_time=3, delta_energy=3.0), (1, 5, 30)),
    "randomly_tile_attractor_128x64": (128, 64, 3000, ATT, 0, False, dict(scheduler=1, reproduction_prob=0.05, regrowth_time=3, delta_energy=3.0, prob_random_walk=0.5), (1, 5, 30)),
    "randomly_tile_walkmap_128x96": (128, 96, 3000, WMAP, 1, True, dict(scheduler=1, walkmap_regrowth=1, reproduction_prob=0.05, regrowth_time=3, delta_energy=3.0), (1, 5, 30)),
    "tiny_ring_2x2": (2, 2, 3, RW, 1, False, dict(reproduction_prob=0.2, regrowth_time=1), (1, 5, 20)),
    "ragged_33x1": (33, 1, 12, RW, 1, False, dict(reproduction_prob=0.1, regrowth_time=2), (1, 5, 20)),
}


def build_case(name):
    nx, ny, n, wt, per, terr, extra, ticks = CASES[name]
    cfgkw, w = make_case(nx, ny, n, wt, per, terrain=terr, **extra)
    state = None
    if wt == ALT:  # scripts/v0.4.jl:97: model.behav_counter[1] = counter, behav = 0
        state = (0, 0, extra["counter"])
    return cfgkw, w, state, ticks
