"""Host-side mirror of the reference's operator interface for the stepping path.

The reference's boundary is functional: two factories return the callbacks that Agents.jl's
`step!` / `run!` accept —

    agent_step! = herbivore_dynamics(; walk_type, eat, reproduce)          reference src/agent_actions.jl:24-71
    model_step! = grass_growth_dynamics(; walkmap)                         reference src/model_actions.jl:15-32
    step!(model, agent_step!, model_step!, n)                              Agents.jl, used by every scripts/v0.*.jl
    run!(model, agent_step!, model_step!, n; adata, mdata)                 scripts/v0.6.jl:204-221

Here the factories keep their names and keywords but return tokens; `step` turns the whole call into ONE
`abm_step(h, n)` of libabm.so, and `run` asks the device for the aggregated columns.  This file is the Python
twin of the Julia shim in INTEGRATION.md (Julia is not installed in the build image).  There is no CPU path: the
default library is the CUDA product and creation fails loudly without a B200; tests may inject another library
object exporting the same ABI (`lib=`).
"""
from __future__ import annotations

import enum
from dataclasses import dataclass

import numpy as np

from . import capi, worlds


class WalkType(enum.IntEnum):
    """`@enum WalkType ...` (reference src/agent_actions.jl:1), same order and values."""
    RANDOM_WALK = 0
    RANDOM_WALK_ATTRACTOR = 1
    RANDOM_WALK_GREGARIOUS = 2
    ALTERNATED_WALK = 3
    RANDOM_WALKMAP = 4


RANDOM_WALK, RANDOM_WALK_ATTRACTOR, RANDOM_WALK_GREGARIOUS, ALTERNATED_WALK, RANDOM_WALKMAP = list(WalkType)


@dataclass(frozen=True)
class EngineAgentStep:
    """What `herbivore_dynamics` returns: the parameters of `custom_agent_step!` (src/agent_actions.jl:29-69)."""
    walk_type: WalkType = WalkType.RANDOM_WALK
    eat: bool = True
    reproduce: bool = True
    prob_random_walk: float = 0.9  # keyword of the closure, src/agent_actions.jl:32


@dataclass(frozen=True)
class EngineModelStep:
    """What `grass_growth_dynamics` returns (src/model_actions.jl:15-17)."""
    walkmap: bool = False


def herbivore_dynamics(*, walk_type: WalkType = WalkType.RANDOM_WALK, eat: bool = True, reproduce: bool = True) -> EngineAgentStep:
    if not isinstance(walk_type, WalkType):
        raise TypeError("walk_type must be a WalkType")  # the reference's keyword is typed `::WalkType`
    return EngineAgentStep(walk_type, bool(eat), bool(reproduce))


def grass_growth_dynamics(*, walkmap: bool = False) -> EngineModelStep:
    return EngineModelStep(bool(walkmap))


_PRODUCT_LIB = None


def product_library() -> capi.CLib:
    """libabm.so (CUDA, sm_100a).  Built in-tree by `__graft_entry__.build()`."""
    global _PRODUCT_LIB
    if _PRODUCT_LIB is None:
        from . import build
        _PRODUCT_LIB = capi.CLib(build.build_libabm(), "abm_")
    return _PRODUCT_LIB


SCHEDULERS = {"by_id": 0, "randomly": 1}  # Agents.Schedulers.by_id / Schedulers.randomly -> abm_config.scheduler


class SheepModel:
    """The `ABM(Sheep, GridSpace(griddims; periodic, metric = :chebyshev); properties, rng, scheduler)` of
    scripts/v0.1.jl:49-64 ... scripts/v0.6.jl:74-101, with the state kept on the device between calls.

    Host views (`fully_grown`, `countdown`, `agents()`) are fetched lazily, as the Julia shim does when a plot or
    `adata` touches them.  Scheduler: `scheduler="by_id"` (default) or `"randomly"` (what every reference script selects,
    scripts/v0.1.jl:63: a fresh Philox-keyed order every tick); randomness: Philox keyed by (seed, tick, id, draw) — include/abm.h.
    """

    def __init__(self, griddims, *, periodic=True, seed=321, regrowth_time=30, delta_energy=4.0, movement_cost=1.0,
                 reproduction_prob=0.001, visual_distance=5, counter=50, attractor=None, land_walkmap=None, elevation=None,
                 astar_metric=capi.ASTAR_DIRECT, astar_penalty=False, lib: capi.CLib | None = None, device=0, scheduler="by_id"):
        self.nx, self.ny = int(griddims[0]), int(griddims[1])
        self.periodic, self.seed = bool(periodic), int(seed)
        self.regrowth_time, self.delta_energy, self.movement_cost = int(regrowth_time), float(delta_energy), float(movement_cost)
        self.reproduction_prob, self.visual_distance, self.counter = float(reproduction_prob), int(visual_distance), int(counter)
        self.attractor = tuple(attractor) if attractor is not None else (self.nx / 2, self.ny / 2)  # griddims ./ 2, scripts/v0.2.jl:57
        self.land_walkmap = None if land_walkmap is None else np.ascontiguousarray(land_walkmap, np.uint8)
        self.elevation = None if elevation is None else np.ascontiguousarray(elevation, np.int64)
        self.astar_metric, self.astar_penalty = int(astar_metric), bool(astar_penalty)
        if scheduler not in SCHEDULERS:
            raise ValueError(f"scheduler must be one of {sorted(SCHEDULERS)} (Agents.Schedulers.by_id / Schedulers.randomly)")
        self.scheduler = scheduler
        self._lib, self._device = lib, int(device)
        self._engine, self._engine_key = None, None
        # host-side initial state (uploaded when the engine is created)
        self._fg = np.zeros((self.ny, self.nx), np.uint8)
        self._cd = np.zeros((self.ny, self.nx), np.int64)
        self._agents = dict(ids=np.zeros(0, np.int64), x=np.zeros(0, np.int64), y=np.zeros(0, np.int64), energy=np.zeros(0, np.float64))
        self._next_id = 1
        self._state = dict(tick=0, behav=0, behav_counter=0)
        self._host_fresh = True
        self._device_init = None  # (n_sheep, init_seed): the engine builds grass and sheep itself (abm_init_random)

    # ---- construction helpers (the host keeps `initialize_model`)
    def set_grass(self, fully_grown, countdown):
        self._sync_to_host()
        self._fg = np.ascontiguousarray(fully_grown, np.uint8).reshape(self.ny, self.nx)
        self._cd = np.ascontiguousarray(countdown, np.int64).reshape(self.ny, self.nx)
        self._drop_engine()

    def add_agents(self, x, y, energy, ids=None):
        """`add_agent_pos!(Sheep(nextid(model), pos, energy, ...), model)` for a batch (scripts/v0.1.jl:67-70)."""
        self._sync_to_host()
        x, y, e = np.asarray(x, np.int64), np.asarray(y, np.int64), np.asarray(energy, np.float64)
        if ids is None:
            ids = np.arange(self._next_id, self._next_id + x.size, dtype=np.int64)
        a = self._agents
        self._agents = dict(ids=np.concatenate([a["ids"], np.asarray(ids, np.int64)]), x=np.concatenate([a["x"], x]),
                            y=np.concatenate([a["y"], y]), energy=np.concatenate([a["energy"], e]))
        self._next_id = max(self._next_id, int(np.max(ids)) + 1) if x.size else self._next_id
        self._drop_engine()

    def init_on_device(self, n_sheep, init_seed=None):
        """The random part of `initialize_model` built by the engine itself (`abm_init_random`): nothing is generated or
        uploaded by the host except the terrain.  Takes effect when the first `step`/`run` creates the engine; the host
        views (`agents()`, `fully_grown`, ...) are then fetched lazily like after any step."""
        self._device_init = (int(n_sheep), int(self.seed if init_seed is None else init_seed))
        self._agents = dict(ids=np.zeros(0, np.int64), x=np.zeros(0, np.int64), y=np.zeros(0, np.int64), energy=np.zeros(0, np.float64))
        self._next_id = int(n_sheep) + 1
        self._drop_engine()

    def set_behaviour(self, behav=0, behav_counter=None):
        """model.behav[1], model.behav_counter[1] (scripts/v0.4.jl:69-71, :97)."""
        self._sync_to_host()
        self._state["behav"] = int(behav)
        self._state["behav_counter"] = self.counter if behav_counter is None else int(behav_counter)
        self._drop_engine()

    # ---- engine lifecycle
    def _drop_engine(self):
        if self._engine is not None:
            self._engine.close()
        self._engine, self._engine_key = None, None

    def _sync_to_host(self):
        if self._engine is not None and not self._host_fresh:
            ids, x, y, e = self._engine.get_agents()
            self._agents = dict(ids=ids, x=x, y=y, energy=e)
            self._fg, self._cd = self._engine.get_grid()
            st = self._engine.get_global_state()
            self._next_id = st.pop("next_id")
            self._state = st
            self._host_fresh = True

    def _ensure_engine(self, agent_step: EngineAgentStep, model_step: EngineModelStep):
        if not isinstance(agent_step, EngineAgentStep) or not isinstance(model_step, EngineModelStep):
            raise TypeError("step/run need the tokens returned by herbivore_dynamics / grass_growth_dynamics")
        key = (agent_step, model_step)
        if self._engine is not None and key == self._engine_key:
            return self._engine
        self._sync_to_host()
        self._drop_engine()
        lib = self._lib or product_library()
        cfg = lib.default_config(
            nx=self.nx, ny=self.ny, periodic=int(self.periodic), walk_type=int(agent_step.walk_type), eat=int(agent_step.eat),
            reproduce=int(agent_step.reproduce), walkmap_regrowth=int(model_step.walkmap), regrowth_time=self.regrowth_time,
            visual_distance=self.visual_distance, counter=self.counter, astar_metric=self.astar_metric, astar_penalty=int(self.astar_penalty),
            device=self._device, prob_random_walk=agent_step.prob_random_walk, delta_energy=self.delta_energy,
            movement_cost=self.movement_cost, reproduction_prob=self.reproduction_prob, attractor_x=self.attractor[0],
            attractor_y=self.attractor[1], seed=self.seed, scheduler=SCHEDULERS[self.scheduler])
        e = lib.create(cfg)
        if self._device_init is None or self.land_walkmap is not None or self.elevation is not None:
            e.set_grid(self._fg, self._cd, self.land_walkmap, self.elevation)  # with a device init only the terrain matters
        if self._device_init is not None:
            e.init_random(*self._device_init)
            self._device_init = None
            self._host_fresh = False
        else:
            a = self._agents
            e.set_agents(a["ids"], a["x"], a["y"], a["energy"], self._next_id)
        e.set_global_state(self._state["tick"], self._state["behav"], self._state["behav_counter"])
        if agent_step.walk_type == WalkType.RANDOM_WALK_GREGARIOUS:
            e.set_weight_lut(worlds.weight_lut(self.visual_distance))  # the host's exp (parity note in include/abm.h)
        self._engine, self._engine_key = e, key
        return e

    # ---- lazily materialised host views
    @property
    def fully_grown(self):
        self._sync_to_host()
        return self._fg

    @property
    def countdown(self):
        self._sync_to_host()
        return self._cd

    def agents(self):
        """ids (ascending), x, y (1-based, `agent.pos`), energy."""
        self._sync_to_host()
        a = self._agents
        return a["ids"], a["x"], a["y"], a["energy"]

    def nagents(self) -> int:
        if self._engine is None and self._device_init is not None:
            return self._device_init[0]
        return self._engine.count_agents() if self._engine is not None else int(self._agents["ids"].size)

    @property
    def tick(self):
        return self._engine.get_global_state()["tick"] if self._engine is not None else self._state["tick"]


def step(model: SheepModel, agent_step: EngineAgentStep, model_step: EngineModelStep, n: int = 1) -> SheepModel:
    """`Agents.step!(model, agent_step!, model_step!, n)`: n ticks of {every agent in id order; model_step!}."""
    e = model._ensure_engine(agent_step, model_step)
    e.step(int(n))
    model._host_fresh = False
    return model


# ---- run! with adata / mdata -------------------------------------------------------------------------------------

def sheep(agent=None):
    """`sheep(a) = a isa Sheep` (scripts/v0.6.jl:204): every agent of this model is a Sheep."""
    return True


def count(values) -> int:
    """Julia `count` over the agents' values."""
    return int(np.count_nonzero(values))


def count_grass(model: SheepModel) -> int:
    """`count_grass(model) = count(model.fully_grown)` (scripts/v0.6.jl:205)."""
    if model._engine is not None:
        return int(model._engine.reduce(capi.REDUCE_COUNT_GROWN))
    return int(np.count_nonzero(model._fg))


class Where:
    """A filter of an adata tuple `(property, aggregator, filter)` that the engine evaluates on the device
    (abm_reduce_where): agents with energy, x, y inside the given inclusive ranges.  Also callable on a host-side agent
    record, so it works wherever a plain predicate does."""

    def __init__(self, energy=(-np.inf, np.inf), x=None, y=None):
        self.energy, self.x, self.y = energy, x, y

    def __call__(self, a):
        ex, (px, py) = a["energy"], a["pos"]
        return (self.energy[0] <= ex <= self.energy[1] and (self.x is None or self.x[0] <= px <= self.x[1])
                and (self.y is None or self.y[0] <= py <= self.y[1]))


_FAST_AGENT = {("energy", "sum"): capi.REDUCE_SUM_ENERGY, ("energy", "maximum"): capi.REDUCE_MAX_ENERGY, ("energy", "max"): capi.REDUCE_MAX_ENERGY,
               ("energy", "minimum"): capi.REDUCE_MIN_ENERGY, ("energy", "min"): capi.REDUCE_MIN_ENERGY}


def _agg_name(f):
    return getattr(f, "__name__", str(f))


def _collect_agents(model: SheepModel, adata):
    row = {}
    for entry in adata:
        if not (isinstance(entry, tuple) and len(entry) in (2, 3)):
            raise ValueError("adata entries are (property_or_function, aggregator[, filter]) tuples")
        prop, agg = entry[0], entry[1]
        flt = entry[2] if len(entry) == 3 else None
        pname = prop if isinstance(prop, str) else _agg_name(prop)
        col = f"{_agg_name(agg)}_{pname}"  # Agents.jl's naming, relied on at src/agent_actions.jl:964
        e = model._engine
        if flt is None and e is not None and prop is sheep and agg is count:
            row[col] = int(e.reduce(capi.REDUCE_COUNT_AGENTS))
        elif flt is None and e is not None and (pname, _agg_name(agg)) in _FAST_AGENT:
            row[col] = e.reduce(_FAST_AGENT[(pname, _agg_name(agg))])
        elif flt is None and e is not None and pname == "energy" and _agg_name(agg) == "mean":
            n = e.reduce(capi.REDUCE_COUNT_AGENTS)
            row[col] = e.reduce(capi.REDUCE_SUM_ENERGY) / n if n else float("nan")
        elif isinstance(flt, Where) and e is not None and ((prop is sheep and agg is count) or (pname, _agg_name(agg)) in _FAST_AGENT or (pname, _agg_name(agg)) == ("energy", "mean")):
            # filtered tuple on the device: no agent leaves the GPU
            kw = dict(energy=flt.energy, x=flt.x, y=flt.y)
            if prop is sheep:
                row[col] = int(e.reduce_where(capi.REDUCE_COUNT_AGENTS, **kw))
            elif _agg_name(agg) == "mean":
                n = e.reduce_where(capi.REDUCE_COUNT_AGENTS, **kw)
                row[col] = e.reduce_where(capi.REDUCE_SUM_ENERGY, **kw) / n if n else float("nan")
            else:
                row[col] = e.reduce_where(_FAST_AGENT[(pname, _agg_name(agg))], **kw)
        else:  # general case: materialise the agents on the host
            ids, x, y, en = model.agents()
            fields = dict(id=ids, energy=en, x=x, y=y)
            if isinstance(prop, str):
                vals = fields[prop] if prop != "pos" else np.stack([x, y], 1)
            else:
                vals = np.array([prop(dict(id=int(i), pos=(int(a), int(b)), energy=float(c))) for i, a, b, c in zip(ids, x, y, en)])
            if flt is not None:
                keep = np.array([bool(flt(dict(id=int(i), pos=(int(a), int(b)), energy=float(c)))) for i, a, b, c in zip(ids, x, y, en)], bool)
                vals = vals[keep]
            row[col] = agg(vals)
    return row


def run(model: SheepModel, agent_step: EngineAgentStep, model_step: EngineModelStep, n: int, *, adata=None, mdata=None, when=None):
    """`run!(model, agent_step!, model_step!, n; adata, mdata)`: returns (agent_df, model_df) with a `step` column;
    data are collected at step 0 and after every step for which `when(model, s)` holds (default: all)."""
    import pandas as pd
    model._ensure_engine(agent_step, model_step)
    arows, mrows = [], []

    def collect(s):
        if adata:
            arows.append(dict(step=s, **_collect_agents(model, adata)))
        if mdata:
            mrows.append(dict(step=s, **{_agg_name(f): f(model) for f in mdata}))

    collect(0)
    for s in range(1, int(n) + 1):
        step(model, agent_step, model_step, 1)
        if when is None or when(model, s):
            collect(s)
    return pd.DataFrame(arows), pd.DataFrame(mrows)


# ---- what plot_abm_model reads ------------------------------------------------------------------------------------

def grasscolor(model: SheepModel):
    """`grasscolor(model) = model.countdown ./ model.regrowth_time` — the `heatarray` of `plot_abm_model`
    (src/plotting.jl:33, :27-57), shape (ny, nx), 1.0 = just eaten / fully grown, towards 0 = about to regrow."""
    if model._engine is not None:  # computed on the device (abm_get_heat): 4 bytes per cell come back instead of 9
        return model._engine.get_heat(0)
    return model.countdown.astype(np.float64) / float(model.regrowth_time)


def penaltymap_heat(model: SheepModel):
    """`heatarray = model -> penaltymap(model.pathfinder)` (scripts/v0.6.jl:179): the elevation the PenaltyMap was built from."""
    if model._engine is not None:
        return model._engine.get_heat(1)
    return np.zeros((model.ny, model.nx), np.float32) if model.elevation is None else np.asarray(model.elevation, np.float32)


def plot_data(model: SheepModel):
    """Everything `plot_abm_model` draws, fetched from the device in one go: sheep positions (x, y; 1-based, for the
    scatter) and the grass heat array.  The GUI itself (Makie) stays on the host side and out of scope."""
    _, x, y, _ = model.agents()
    return dict(x=x, y=y, heatarray=grasscolor(model), colorrange=(0.0, 1.0))


# ---- initialize_model --------------------------------------------------------------------------------------------

def initialize_model(*, n_sheep=40, griddims=(80, 80), regrowth_time=30, delta_energy_sheep=4, sheep_reproduce=0.001, movement_cost=1,
                     visual_distance=5, seed=321, periodic=False, counter=50, water_level=1, mountain_level=18, elevation=None,
                     use_walkmap=False, synthetic_terrain_seed=7, astar_metric=capi.ASTAR_DIRECT, astar_penalty=False, lib=None, device=0, on_device=False, scheduler="by_id") -> SheepModel:
    """The scripts' `initialize_model(; kwargs)` (scripts/v0.1.jl:36-81 ... scripts/v0.6.jl:36-135), same keywords
    (Δenergy_sheep spelled delta_energy_sheep).  v0.1/v0.2 are non-periodic, v0.3-v0.6 periodic; v0.5/v0.6 build
    `land_walkmap = water_level .< elevation .< mountain_level` from a heightmap — the scripts download theirs, here
    a seeded synthetic one stands in unless `elevation` is given.  Initial draws come from numpy's Philox stream (the
    scripts use MersenneTwister(seed); initialisation is host-side and not part of the parity contract).
    `scheduler="randomly"` is the scripts' `scheduler = Schedulers.randomly` (scripts/v0.1.jl:63); the default steps by id."""
    nx, ny = int(griddims[0]), int(griddims[1])
    wm = None
    if use_walkmap or elevation is not None:
        if elevation is None:
            elevation = worlds.make_terrain(nx, ny, seed=synthetic_terrain_seed)
        wm = worlds.walkmap_from_elevation(np.asarray(elevation), water_level, mountain_level)
    w = None if on_device else worlds.make_world(nx, ny, n_sheep, seed=seed, regrowth_time=regrowth_time, delta_energy=int(delta_energy_sheep), walkmap=wm, periodic=bool(periodic))
    m = SheepModel((nx, ny), periodic=periodic, seed=seed, regrowth_time=regrowth_time, delta_energy=delta_energy_sheep, movement_cost=movement_cost,
                   reproduction_prob=sheep_reproduce, visual_distance=visual_distance, counter=counter, land_walkmap=wm, elevation=elevation,
                   astar_metric=astar_metric, astar_penalty=astar_penalty, lib=lib, device=device, scheduler=scheduler)
    if on_device:  # grass and sheep are drawn by the engine (abm_init_random): a 2^28-sheep world needs no host arrays
        m.init_on_device(n_sheep, seed)
    else:
        m.set_grass(w["fully_grown"], w["countdown"])
        m.add_agents(w["x"], w["y"], w["energy"], w["ids"])
    m.set_behaviour(0, counter)  # model.behav_counter[1] = counter, scripts/v0.4.jl:97
    return m
