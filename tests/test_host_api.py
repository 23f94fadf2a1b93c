"""Host-side mirror of the reference's operator interface (abm_b200.model): factories, step, run with adata/mdata.
The library object is injected (`lib=`): the emulated engine here, the product library under -m gpu."""
import numpy as np
import pytest

import abm_b200 as A
from conftest import assert_same_state


def _script_model(lib, version, **kw):
    """initialize_model with the defaults of scripts/v0.N.jl"""
    if version == "v0.1":
        return A.initialize_model(periodic=False, lib=lib, **kw), A.herbivore_dynamics(), A.grass_growth_dynamics()
    if version == "v0.3":
        return (A.initialize_model(periodic=True, visual_distance=5, lib=lib, **kw),
                A.herbivore_dynamics(walk_type=A.WalkType.RANDOM_WALK_GREGARIOUS, eat=True, reproduce=True), A.grass_growth_dynamics())
    if version == "v0.5":
        return (A.initialize_model(periodic=True, use_walkmap=True, sheep_reproduce=0.004, lib=lib, **kw),
                A.herbivore_dynamics(walk_type=A.WalkType.RANDOM_WALKMAP), A.grass_growth_dynamics(walkmap=True))
    if version == "v0.6":
        return (A.initialize_model(periodic=True, use_walkmap=True, sheep_reproduce=0.004, counter=50, astar_metric=A.capi.ASTAR_MAX, astar_penalty=True, lib=lib, **kw),
                A.herbivore_dynamics(walk_type=A.WalkType.ALTERNATED_WALK), A.grass_growth_dynamics(walkmap=True))
    raise KeyError(version)


def test_factories_keep_the_reference_signature():
    a = A.herbivore_dynamics()
    assert (a.walk_type, a.eat, a.reproduce, a.prob_random_walk) == (A.WalkType.RANDOM_WALK, True, True, 0.9)
    assert A.grass_growth_dynamics().walkmap is False and A.grass_growth_dynamics(walkmap=True).walkmap is True
    assert [w.name for w in A.WalkType] == ["RANDOM_WALK", "RANDOM_WALK_ATTRACTOR", "RANDOM_WALK_GREGARIOUS", "ALTERNATED_WALK", "RANDOM_WALKMAP"]
    with pytest.raises(TypeError):
        A.herbivore_dynamics(walk_type=2)
    with pytest.raises(TypeError):
        A.herbivore_dynamics(True)  # keyword-only, like the Julia factory


def _check_step_and_run(lib, oracle):
    for version in ("v0.1", "v0.3", "v0.5", "v0.6"):
        m, astep, mstep = _script_model(lib, version)
        ref, _, _ = _script_model(oracle, version)
        n0 = m.nagents()
        A.step(m, astep, mstep, 7)
        A.step(ref, astep, mstep, 7)
        assert m.tick == 7
        ids, x, y, e = m.agents()
        rids, rx, ry, re = ref.agents()
        assert np.array_equal(ids, rids) and np.array_equal(x, rx) and np.array_equal(y, ry) and np.array_equal(e, re)
        assert np.array_equal(m.fully_grown, ref.fully_grown) and np.array_equal(m.countdown, ref.countdown)
        assert n0 == 40
        # run!(model, agent_step!, model_step!, n; adata = [(sheep, count)], mdata = [count_grass]), scripts/v0.6.jl:204-221
        adf, mdf = A.run(m, astep, mstep, 5, adata=[(A.sheep, A.count), ("energy", sum), ("energy", max)], mdata=[A.count_grass])
        assert list(adf.columns) == ["step", "count_sheep", "sum_energy", "max_energy"] and list(mdf.columns) == ["step", "count_grass"]
        assert list(adf["step"]) == [0, 1, 2, 3, 4, 5]
        ids, x, y, e = m.agents()
        assert adf["count_sheep"].iloc[-1] == ids.size and adf["sum_energy"].iloc[-1] == e.sum() and adf["max_energy"].iloc[-1] == e.max()
        assert mdf["count_grass"].iloc[-1] == int(m.fully_grown.sum())


def test_step_and_run_on_the_emulated_engine(emu, oracle):
    _check_step_and_run(emu, oracle)


@pytest.mark.gpu
def test_step_and_run_on_the_gpu(libabm, oracle):
    _check_step_and_run(libabm, oracle)


def test_scheduler_randomly_through_the_mirror(emu, oracle):
    """`scheduler = Schedulers.randomly` (scripts/v0.1.jl:63): same trajectory from engine and oracle, and not the by-id one."""
    runs = {}
    for sched in ("by_id", "randomly"):
        m, astep, mstep = _script_model(emu, "v0.3", scheduler=sched, n_sheep=900)
        ref, _, _ = _script_model(oracle, "v0.3", scheduler=sched, n_sheep=900)
        A.step(m, astep, mstep, 6)
        A.step(ref, astep, mstep, 6)
        for a, b in zip(m.agents(), ref.agents()):
            assert np.array_equal(a, b)
        assert np.array_equal(m.fully_grown, ref.fully_grown)
        runs[sched] = m.agents()
    assert not (runs["by_id"][1].shape == runs["randomly"][1].shape and np.array_equal(runs["by_id"][1], runs["randomly"][1]))
    with pytest.raises(ValueError):
        A.initialize_model(scheduler="fastest", lib=emu)


def test_changing_the_rules_keeps_the_state(emu):
    m, astep, mstep = _script_model(emu, "v0.3")
    A.step(m, astep, mstep, 3)
    ids = m.agents()[0].copy()
    A.step(m, A.herbivore_dynamics(walk_type=A.WalkType.RANDOM_WALK, reproduce=False), mstep, 1)  # new engine, same world
    assert m.tick == 4 and set(m.agents()[0]) <= set(ids)


def test_product_library_is_the_default_and_has_no_cpu_path():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    m = A.initialize_model()
    with pytest.raises(A.AbmError) as ei:
        A.step(m, A.herbivore_dynamics(), A.grass_growth_dynamics(), 1)
    assert ei.value.code == A.capi.ABM_E_CUDA


def test_initialize_model_on_the_device(emu, oracle):
    """`initialize_model(...; on_device=True)`: the engine draws grass and sheep itself (abm_init_random); the emulated
    engine and the oracle, driven through the same mirror, build and step the same world; changing the rules afterwards
    keeps it."""
    models = [A.initialize_model(n_sheep=120, griddims=(64, 64), periodic=True, use_walkmap=True, seed=5, on_device=True, lib=lib) for lib in (emu, oracle)]
    astep, mstep = A.herbivore_dynamics(walk_type=A.WalkType.RANDOM_WALKMAP), A.grass_growth_dynamics(walkmap=True)
    assert all(m.nagents() == 120 for m in models)
    for m in models:
        A.step(m, astep, mstep, 4)
    (ia, xa, ya, ea), (ib, xb, yb, eb) = models[0].agents(), models[1].agents()
    assert np.array_equal(ia, ib) and np.array_equal(xa, xb) and np.array_equal(ya, yb) and np.array_equal(ea, eb) and ia.size > 0
    assert np.array_equal(models[0].fully_grown, models[1].fully_grown) and models[0].tick == 4
    m = models[0]
    wm = m.land_walkmap.astype(bool)
    assert wm[ya - 1, xa - 1].all()
    A.step(m, A.herbivore_dynamics(walk_type=A.WalkType.RANDOM_WALK), A.grass_growth_dynamics(), 1)  # new engine, same world
    assert m.tick == 5 and m.nagents() > 0


def _check_filtered_adata_and_heat(lib, oracle):
    """Filtered adata tuples `(property, aggregator, filter)` whose filter is a box are reduced on the device
    (abm_reduce_where) and equal the host-side evaluation of the same predicate; the heat arrays of the plots come from the
    device (abm_get_heat) and equal `countdown ./ regrowth_time` / the elevation."""
    m, astep, mstep = _script_model(lib, "v0.6")
    A.step(m, astep, mstep, 4)
    w = A.Where(energy=(2.0, 9.0), x=(10, 60), y=(5, 70))
    adf, _ = A.run(m, astep, mstep, 3, adata=[(A.sheep, A.count, w), ("energy", sum, w), ("energy", max, w), ("energy", min, w)])
    ids, x, y, e = m.agents()
    keep = (e >= 2.0) & (e <= 9.0) & (x >= 10) & (x <= 60) & (y >= 5) & (y <= 70)
    assert keep.any() and not keep.all()
    last = adf.iloc[-1]
    assert last["count_sheep"] == keep.sum() and last["sum_energy"] == e[keep].sum() and last["max_energy"] == e[keep].max() and last["min_energy"] == e[keep].min()
    # the same numbers from the oracle's restatement of the entry point
    ref, _, _ = _script_model(oracle, "v0.6")
    A.step(ref, astep, mstep, 7)
    assert ref._engine.reduce_where(A.capi.REDUCE_SUM_ENERGY, energy=(2.0, 9.0), x=(10, 60), y=(5, 70)) == last["sum_energy"]
    heat = A.grasscolor(m)
    assert heat.dtype == np.float32 and np.allclose(heat, m.countdown / m.regrowth_time, rtol=1e-6)
    assert np.array_equal(A.penaltymap_heat(m), np.asarray(m.elevation, np.float32))


def test_filtered_adata_and_heat_arrays_emulated(emu, oracle):
    _check_filtered_adata_and_heat(emu, oracle)


@pytest.mark.gpu
def test_filtered_adata_and_heat_arrays_on_the_gpu(libabm, oracle):
    _check_filtered_adata_and_heat(libabm, oracle)


def test_plot_data_matches_the_reference_heat_array(emu):
    """`plot_abm_model` draws `countdown ./ regrowth_time` and the sheep positions (src/plotting.jl:27-57)."""
    m, astep, mstep = _script_model(emu, "v0.1")
    A.step(m, astep, mstep, 3)
    d = A.plot_data(m)
    assert d["heatarray"].shape == (m.ny, m.nx) and d["heatarray"].min() >= 0 and d["heatarray"].max() <= 1
    assert np.allclose(d["heatarray"], m.countdown / m.regrowth_time, rtol=1e-6)
    assert d["x"].size == m.nagents() and d["x"].min() >= 1 and d["y"].max() <= m.ny
