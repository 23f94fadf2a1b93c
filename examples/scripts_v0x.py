"""The six grid scripts of the reference (scripts/v0.1.jl ... scripts/v0.6.jl) as they read after the switch to libabm.so.

Each `vNN()` below keeps the script's own keywords and call sequence — `initialize_model(; ...)`, the two factories
`herbivore_dynamics` / `grass_growth_dynamics`, `step!` for the frames the script renders (`abmvideo(...; frames = 100)`,
scripts/v0.1.jl:98-106) and, for v0.6, `run!` with `adata = [(sheep, count)]`, `mdata = [count_grass]` (scripts/v0.6.jl:204-221)
— through the Python twin of the Julia shim (`abm_b200.model`).  What the scripts draw with Makie is returned
as arrays by `plot_data` (sheep positions + the grass heat array of `plot_abm_model`, src/plotting.jl:27-57).

    python examples/scripts_v0x.py            # on a B200: every script, 100 ticks each, population and grass printed
    python examples/scripts_v0x.py v0.3 v0.6  # some of them

`lib=` lets the tests run the same functions on the emulated engine (tests/test_examples.py); the default is the CUDA library
and there is no CPU path.
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import abm_b200 as A  # noqa: E402

FRAMES = 100  # `frames = 100` of every abmvideo call


def _frames(model, agent_step, model_step, frames, every=25):
    """abmvideo's loop: one step! per frame; plot_data is what plot_abm_model would draw."""
    shown = []
    for f in range(1, frames + 1):
        A.step(model, agent_step, model_step, 1)
        if f % every == 0 or f == frames:
            d = A.plot_data(model)
            shown.append(dict(frame=f, sheep=int(d["x"].size), grass=A.count_grass(model)))
    return shown


def v01(lib=None, frames=FRAMES, scheduler="randomly", **kw):
    """scripts/v0.1.jl: 80 x 80, 40 sheep, random walk, eat, reproduce, grass regrowth; not periodic."""
    model = A.initialize_model(n_sheep=40, griddims=(80, 80), regrowth_time=30, delta_energy_sheep=4, sheep_reproduce=0.001,
                               movement_cost=1, seed=321, periodic=False, scheduler=scheduler, lib=lib, **kw)
    return model, _frames(model, A.herbivore_dynamics(), A.grass_growth_dynamics(), frames)


def v02(lib=None, frames=FRAMES, scheduler="randomly", **kw):
    """scripts/v0.2.jl: + home-range attractor in the middle of the grid (walk_type = RANDOM_WALK_ATTRACTOR)."""
    model = A.initialize_model(periodic=False, scheduler=scheduler, lib=lib, **kw)
    return model, _frames(model, A.herbivore_dynamics(walk_type=A.WalkType.RANDOM_WALK_ATTRACTOR), A.grass_growth_dynamics(), frames)


def v03(lib=None, frames=FRAMES, scheduler="randomly", **kw):
    """scripts/v0.3.jl: periodic space, gregarious walk weighted by the closest sheep within visual_distance = 5."""
    model = A.initialize_model(periodic=True, visual_distance=5, scheduler=scheduler, lib=lib, **kw)
    return model, _frames(model, A.herbivore_dynamics(walk_type=A.WalkType.RANDOM_WALK_GREGARIOUS), A.grass_growth_dynamics(), frames)


def v04(lib=None, frames=FRAMES, scheduler="randomly", **kw):
    """scripts/v0.4.jl: the model-wide behaviour counter (50 agent steps per behaviour: random walk / A* route / rest)."""
    model = A.initialize_model(periodic=True, sheep_reproduce=0.004, counter=50, use_walkmap=True, water_level=-1, mountain_level=10 ** 6,
                               astar_metric=A.capi.ASTAR_DIRECT, scheduler=scheduler, lib=lib, **kw)  # every cell walkable: AStar(space)
    return model, _frames(model, A.herbivore_dynamics(walk_type=A.WalkType.ALTERNATED_WALK), A.grass_growth_dynamics(), frames)


def v05(lib=None, frames=FRAMES, scheduler="randomly", **kw):
    """scripts/v0.5.jl: elevation map, water and mountains are not walkable, grass regrows on land only."""
    model = A.initialize_model(periodic=True, sheep_reproduce=0.004, use_walkmap=True, water_level=1, mountain_level=18,
                               scheduler=scheduler, lib=lib, **kw)
    return model, _frames(model, A.herbivore_dynamics(walk_type=A.WalkType.RANDOM_WALKMAP), A.grass_growth_dynamics(walkmap=True), frames)


def v06(lib=None, frames=FRAMES, scheduler="randomly", **kw):
    """scripts/v0.6.jl: v0.4's behaviours on v0.5's terrain, routes by A* over PenaltyMap(elevation, MaxDistance); run! with
    adata = [(sheep, count)] and mdata = [count_grass] (:204-221), then the heat array the script plots (:179)."""
    model = A.initialize_model(periodic=True, sheep_reproduce=0.004, counter=50, use_walkmap=True, astar_metric=A.capi.ASTAR_MAX,
                               astar_penalty=True, scheduler=scheduler, lib=lib, **kw)
    agent_step = A.herbivore_dynamics(walk_type=A.WalkType.ALTERNATED_WALK)
    model_step = A.grass_growth_dynamics(walkmap=True)
    adf, mdf = A.run(model, agent_step, model_step, frames, adata=[(A.sheep, A.count)], mdata=[A.count_grass])
    heat = A.penaltymap_heat(model)
    return model, dict(agent_df=adf, model_df=mdf, heatarray_shape=tuple(heat.shape))


SCRIPTS = {"v0.1": v01, "v0.2": v02, "v0.3": v03, "v0.4": v04, "v0.5": v05, "v0.6": v06}

if __name__ == "__main__":
    for name in (sys.argv[1:] or list(SCRIPTS)):
        model, out = SCRIPTS[name]()
        if isinstance(out, dict):
            print(name, "run! ->", out["agent_df"].tail(1).to_dict("records"), out["model_df"].tail(1).to_dict("records"), "heat", out["heatarray_shape"])
        else:
            print(name, out[-1])
