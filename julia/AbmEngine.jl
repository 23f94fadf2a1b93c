# AbmEngine.jl — Julia host shim for libabm.so (include/abm.h), the B200 stepping engine of the GridSpace sheep model.
#
# What a maintainer of tatakof/Animal-Movement-ABM adds: `include(".../julia/AbmEngine.jl"); using .AbmEngine` next to
# `include(srcdir("agent_actions.jl"))`.  The factories keep their names and keywords (reference src/agent_actions.jl:24-71,
# src/model_actions.jl:15-32) but return tokens; `Agents.step!(model, agent_step!, model_step!, n)` specialised on the tokens
# is ONE `ccall(:abm_step)`; `Agents.run!` gets a fast path whose `adata` / `mdata` aggregations come from `abm_reduce`.
# Julia is not installed in the build image, so this file is syntax-reviewed, not executed there; its Python twin
# (abm_b200/model.py: same factories, same call order) is what the tests run.
module AbmEngine

using Agents
import Agents: step!, run!

export herbivore_dynamics, grass_growth_dynamics, WalkType, RANDOM_WALK, RANDOM_WALK_ATTRACTOR, RANDOM_WALK_GREGARIOUS,
       ALTERNATED_WALK, RANDOM_WALKMAP, sync_host!, count_grass, sheep, engine_handle, release_engine!

const libabm = get(ENV, "LIBABM", "libabm")          # libabm.so on the library path, or its full path in ENV["LIBABM"]

# reference: `@enum WalkType ...`, src/agent_actions.jl:1 (same order, same values as include/abm.h)
@enum WalkType RANDOM_WALK RANDOM_WALK_ATTRACTOR RANDOM_WALK_GREGARIOUS ALTERNATED_WALK RANDOM_WALKMAP

# ---- struct abm_config (include/abm.h), field for field -------------------------------------------------------------
Base.@kwdef mutable struct AbmConfig
    struct_size::Int32 = 0
    nx::Int32 = 80;  ny::Int32 = 80;  periodic::Int32 = 0
    walk_type::Int32 = 0;  eat::Int32 = 1;  reproduce::Int32 = 1
    walkmap_regrowth::Int32 = 0;  regrowth_time::Int32 = 30;  visual_distance::Int32 = 5;  counter::Int32 = 50
    randomwalk_ifempty::Int32 = 1;  astar_metric::Int32 = 0;  astar_penalty::Int32 = 0
    device::Int32 = 0;  reserved0::Int32 = 0
    prob_random_walk::Float64 = 0.9;  delta_energy::Float64 = 4;  movement_cost::Float64 = 1
    reproduction_prob::Float64 = 0.001;  attractor_x::Float64 = 40;  attractor_y::Float64 = 40
    seed::UInt64 = 321
    strip_y0::Int32 = 0;  strip_ny::Int32 = 0;  n_ranks::Int32 = 0;  rank::Int32 = 0
    scheduler::Int32 = 0;  reserved1::Int32 = 0   # 0 = Schedulers.by_id, 1 = Schedulers.randomly (Philox-keyed permutation per tick)
end

function __init__()
    # the mirror above must be the library's struct: abm_config_size() is sizeof(abm_config) of the loaded libabm.so
    want = ccall((:abm_config_size, libabm), Cint, ())
    @assert sizeof(AbmConfig) == want "AbmConfig is $(sizeof(AbmConfig)) bytes, libabm's abm_config is $want: update julia/AbmEngine.jl"
end

last_error(h) = unsafe_string(ccall((:abm_last_error, libabm), Cstring, (Ptr{Cvoid},), h))
check(h, rc) = rc == 0 ? nothing : error("libabm error $rc: ", last_error(h))

# ---- the factories: same names, same keywords, tokens instead of closures --------------------------------------------
struct EngineAgentStep
    walk_type::WalkType; eat::Bool; reproduce::Bool; prob_random_walk::Float64
end
struct EngineModelStep
    walkmap::Bool
end
"herbivore_dynamics(; walk_type, eat, reproduce) — reference src/agent_actions.jl:24-33; prob_random_walk is the closure's 0.9 (:32)"
herbivore_dynamics(; walk_type::WalkType = RANDOM_WALK, eat::Bool = true, reproduce::Bool = true) =
    EngineAgentStep(walk_type, eat, reproduce, 0.9)
"grass_growth_dynamics(; walkmap) — reference src/model_actions.jl:15-17"
grass_growth_dynamics(; walkmap::Bool = false) = EngineModelStep(walkmap)

# ---- per-model engine state (kept outside the model's NamedTuple properties, which are immutable) --------------------
mutable struct Engine
    h::Ptr{Cvoid}
    rules::Tuple{EngineAgentStep,EngineModelStep}
    host_stale::Bool            # the device holds newer positions / energies / grass than the Julia objects
    device_stale::Bool          # the Julia side changed the model since the last upload
end
const ENGINES = IdDict{Any,Engine}()

engine_handle(model) = ENGINES[model].h
function release_engine!(model)
    e = pop!(ENGINES, model, nothing)
    e === nothing || ccall((:abm_destroy, libabm), Cvoid, (Ptr{Cvoid},), e.h)
    nothing
end

prop(model, name, default) = hasproperty(model, name) ? getproperty(model, name) : default

function create_handle(model, a::EngineAgentStep, m::EngineModelStep)
    nx, ny = size(model.space)
    s = first(allagents(model))          # sheep parameters are model-uniform in every script (scripts/v0.1.jl:67-70)
    att = prop(model, :attractor, (nx / 2, ny / 2))
    cfg = AbmConfig(nx = nx, ny = ny, periodic = all(Agents.spacesize(model) .> 0) && model.space.periodic ? 1 : 0,
                    walk_type = Int32(a.walk_type), eat = a.eat, reproduce = a.reproduce, walkmap_regrowth = m.walkmap,
                    regrowth_time = model.regrowth_time, visual_distance = floor(Int32, s.visual_distance),
                    counter = prop(model, :counter, 50), prob_random_walk = a.prob_random_walk,
                    delta_energy = s.Δenergy, movement_cost = s.movement_cost, reproduction_prob = s.reproduction_prob,
                    attractor_x = att[1], attractor_y = att[2], seed = prop(model, :seed, 321),
                    scheduler = model.scheduler === Schedulers.randomly ? 1 : 0,   # scripts/v0.1.jl:63; anything else is stepped by id
                    astar_metric = prop(model, :astar_max_distance, false) ? 1 : 0, astar_penalty = hasproperty(model, :elevation) ? 1 : 0)
    cfg.struct_size = sizeof(AbmConfig)
    h = Ref{Ptr{Cvoid}}(C_NULL)
    rc = ccall((:abm_create, libabm), Cint, (Ref{AbmConfig}, Ref{Ptr{Cvoid}}), cfg, h)
    rc == 0 || error("abm_create ($rc): ", last_error(C_NULL))
    if a.walk_type == RANDOM_WALK_GREGARIOUS     # the HOST's exp decides the wsample index (include/abm.h, src/agent_actions.jl:319)
        r = cfg.visual_distance
        lut = Float64[exp(-sqrt(k) / 3) for k in 0:2(r + 1)^2]
        check(h[], ccall((:abm_set_weight_lut, libabm), Cint, (Ptr{Cvoid}, Ptr{Float64}, Int32), h[], lut, length(lut)))
    end
    return h[]
end

# model -> device.  Grid arrays are column-major with x fastest: exactly Julia's Matrix / an unpacked BitMatrix.
function upload!(model, e::Engine)
    h = e.h
    fg = Matrix{UInt8}(model.fully_grown);  cd = Matrix{Int64}(model.countdown)
    wm = hasproperty(model, :land_walkmap) ? Matrix{UInt8}(model.land_walkmap) : nothing
    el = hasproperty(model, :elevation) ? Matrix{Int64}(model.elevation) : nothing
    GC.@preserve fg cd wm el check(h, ccall((:abm_set_grid, libabm), Cint, (Ptr{Cvoid}, Ptr{UInt8}, Ptr{Int64}, Ptr{UInt8}, Ptr{Int64}),
        h, fg, cd, wm === nothing ? C_NULL : pointer(wm), el === nothing ? C_NULL : pointer(el)))
    ids = sort!(collect(allids(model)))                      # Schedulers.by_id order
    xs = Int64[model[i].pos[1] for i in ids];  ys = Int64[model[i].pos[2] for i in ids]
    es = Float64[model[i].energy for i in ids]
    check(h, ccall((:abm_set_agents, libabm), Cint, (Ptr{Cvoid}, Int64, Ptr{Int64}, Ptr{Int64}, Ptr{Int64}, Ptr{Float64}, Int64),
                   h, length(ids), Int64.(ids), xs, ys, es, nextid(model)))
    if e.rules[1].walk_type == ALTERNATED_WALK               # model.behav / model.behav_counter (scripts/v0.4.jl:69-71)
        tick = Ref{Int64}(0)
        ccall((:abm_get_global_state, libabm), Cint, (Ptr{Cvoid}, Ref{Int64}, Ptr{Int64}, Ptr{Int64}, Ptr{Int64}), h, tick, C_NULL, C_NULL, C_NULL)
        check(h, ccall((:abm_set_global_state, libabm), Cint, (Ptr{Cvoid}, Int64, Int64, Int64), h, tick[], model.behav[1], model.behav_counter[1]))
    end
    e.device_stale = false;  e.host_stale = false
end

function engine!(model, a::EngineAgentStep, m::EngineModelStep)
    e = get(ENGINES, model, nothing)
    if e !== nothing && e.rules != (a, m)                    # new rules: new handle, same world (the tick carries over)
        sync_host!(model)
        tick = Ref{Int64}(0)
        ccall((:abm_get_global_state, libabm), Cint, (Ptr{Cvoid}, Ref{Int64}, Ptr{Int64}, Ptr{Int64}, Ptr{Int64}), e.h, tick, C_NULL, C_NULL, C_NULL)
        release_engine!(model)
        e = Engine(create_handle(model, a, m), (a, m), false, true)
        check(e.h, ccall((:abm_set_global_state, libabm), Cint, (Ptr{Cvoid}, Int64, Int64, Int64), e.h, tick[], 0, 0))
        ENGINES[model] = e
    elseif e === nothing
        e = Engine(create_handle(model, a, m), (a, m), false, true)
        ENGINES[model] = e
    end
    e.device_stale && upload!(model, e)
    return e
end

# ---- Agents.step! specialised on the tokens: ONE ccall for n ticks (reference call: scripts/v0.3.jl:124 via abmvideo) ----
function step!(model::ABM, a::EngineAgentStep, m::EngineModelStep, n::Integer = 1, agents_first::Bool = true)
    agents_first || error("the engine steps agents first, like every script of the reference")
    e = engine!(model, a, m)
    check(e.h, ccall((:abm_step, libabm), Cint, (Ptr{Cvoid}, Int64), e.h, n))
    e.host_stale = true                   # positions / energies / grass are fetched lazily (sync_host!)
    return model
end

# ---- device -> model: rebuild the Julia objects when a plot or a generic adata entry needs them -----------------------
function sync_host!(model)
    e = get(ENGINES, model, nothing)
    (e === nothing || !e.host_stale) && return model
    h = e.h
    n = Ref{Int64}(0)
    check(h, ccall((:abm_count_agents, libabm), Cint, (Ptr{Cvoid}, Ref{Int64}), h, n))
    ids = Vector{Int64}(undef, n[]); xs = similar(ids); ys = similar(ids); es = Vector{Float64}(undef, n[])
    check(h, ccall((:abm_get_agents, libabm), Cint, (Ptr{Cvoid}, Ref{Int64}, Ptr{Int64}, Ptr{Int64}, Ptr{Int64}, Ptr{Float64}),
                   h, n, ids, xs, ys, es))
    alive = Set(ids)
    proto = first(allagents(model))       # offspring copy the parent's parameters (src/agent_actions.jl:175): model-uniform
    for id in collect(allids(model))      # deaths (reduce_energy!, src/agent_actions.jl:85-92)
        id in alive || kill_agent!(model[id], model)
    end
    for k in eachindex(ids)
        id, pos = ids[k], (Int(xs[k]), Int(ys[k]))
        if haskey(model.agents, id)       # survivors: new cell, new energy
            ag = model[id]
            ag.pos == pos || move_agent!(ag, pos, model)
            ag.energy = es[k]
        else                              # births (reproduce!, :172-178): same type and parameters, the engine's id
            ag = deepcopy(proto)
            ag.id = id;  ag.pos = pos;  ag.energy = es[k]
            add_agent_pos!(ag, model)
        end
    end
    next = Ref{Int64}(0)
    ccall((:abm_get_global_state, libabm), Cint, (Ptr{Cvoid}, Ptr{Int64}, Ptr{Int64}, Ptr{Int64}, Ref{Int64}), h, C_NULL, C_NULL, C_NULL, next)
    model.maxid[] = next[] - 1            # nextid(model) = maxid + 1 (Agents.jl 5.14)
    fg = Matrix{UInt8}(undef, size(model.space)...);  cd = Matrix{Int64}(undef, size(model.space)...)
    check(h, ccall((:abm_get_grid, libabm), Cint, (Ptr{Cvoid}, Ptr{UInt8}, Ptr{Int64}), h, fg, cd))
    model.fully_grown .= fg .== 1;  model.countdown .= cd
    if e.rules[1].walk_type == ALTERNATED_WALK
        b = Ref{Int64}(0);  bc = Ref{Int64}(0)
        ccall((:abm_get_global_state, libabm), Cint, (Ptr{Cvoid}, Ptr{Int64}, Ref{Int64}, Ref{Int64}, Ptr{Int64}), h, C_NULL, b, bc, C_NULL)
        model.behav[1] = b[];  model.behav_counter[1] = bc[]
    end
    e.host_stale = false
    return model
end

# ---- run!: adata / mdata aggregations from the device (scripts/v0.6.jl:204-221; column names as Agents.jl builds them,
#      which the reference's plotting fork relies on: src/agent_actions.jl:964,977) ------------------------------------------
const REDUCE = Dict(:count_agents => 0, :count_grown => 1, :sum_energy => 2, :max_energy => 3, :min_energy => 4)
function device_reduce(model, what::Symbol)
    e = ENGINES[model]
    out = Ref{Float64}(0)
    check(e.h, ccall((:abm_reduce, libabm), Cint, (Ptr{Cvoid}, Int32, Ref{Float64}), e.h, REDUCE[what], out))
    out[]
end
sheep(a) = true                                              # scripts/v0.6.jl:204: adata = [(sheep, count)]
count_grass(model) = haskey(ENGINES, model) && ENGINES[model].host_stale ? Int(device_reduce(model, :count_grown)) : count(model.fully_grown)

# (property_or_function, aggregator) pairs the engine can answer without downloading the agents
function device_adata(model, entry)
    entry isa Tuple && length(entry) == 2 || return nothing
    f, agg = entry
    n = device_reduce(model, :count_agents)
    f === sheep && agg === count && return Int(n)
    f === :energy && agg === sum && return device_reduce(model, :sum_energy)
    f === :energy && agg === maximum && return device_reduce(model, :max_energy)
    f === :energy && agg === minimum && return device_reduce(model, :min_energy)
    f === :energy && nameof(agg) === :mean && return n == 0 ? NaN : device_reduce(model, :sum_energy) / n
    return nothing
end

function run!(model::ABM, a::EngineAgentStep, m::EngineModelStep, n::Integer; adata = nothing, mdata = nothing, when = true, kwargs...)
    fast = (adata === nothing || all(en -> device_adata_supported(en), adata)) && (mdata === nothing || all(f -> f === count_grass, mdata))
    if !fast                              # anything else: Agents.jl's own run! on top of step! + sync_host!
        return invoke(run!, Tuple{ABM,Any,Any,Any}, model, (ag, mo) -> nothing, mo -> (step!(mo, a, m, 1); sync_host!(mo)), n; adata, mdata, when, kwargs...)
    end
    engine!(model, a, m)
    acols = adata === nothing ? String[] : [string(nameof(en[2]), "_", en[1] isa Symbol ? en[1] : nameof(en[1])) for en in adata]
    mcols = mdata === nothing ? String[] : [string(nameof(f)) for f in mdata]
    arows = Vector{Any}[];  mrows = Vector{Any}[]
    collect_now(s) = when === true || (when isa AbstractVector ? s in when : when(model, s))
    for s in 0:n
        s > 0 && step!(model, a, m, 1)
        if collect_now(s)
            adata === nothing || push!(arows, Any[s; [device_adata(model, en) for en in adata]])
            mdata === nothing || push!(mrows, Any[s; [Int(device_reduce(model, :count_grown)) for _ in mdata]])
        end
    end
    todf(cols, rows) = Agents.DataFrames.DataFrame(permutedims(reduce(hcat, rows)), ["step"; cols])
    return (adata === nothing ? Agents.DataFrames.DataFrame() : todf(acols, arows)), (mdata === nothing ? Agents.DataFrames.DataFrame() : todf(mcols, mrows))
end
device_adata_supported(en) = en isa Tuple && length(en) == 2 &&
    ((en[1] === sheep && en[2] === count) || (en[1] === :energy && (en[2] in (sum, maximum, minimum) || nameof(en[2]) === :mean)))

# ---- the narrow per-tick round trip (abm_*_packed): for hosts that draw every tick (abmplot / abmvideo) -----------------
"positions as (x, y) tuples and energies, in id order, through the 16-byte agent records"
function agents_packed(model)
    h = ENGINES[model].h
    n = Ref{Int64}(0);  ccall((:abm_count_agents, libabm), Cint, (Ptr{Cvoid}, Ref{Int64}), h, n)
    ids = Vector{UInt32}(undef, n[]); xy = similar(ids); es = Vector{Float64}(undef, n[])
    check(h, ccall((:abm_get_agents_packed, libabm), Cint, (Ptr{Cvoid}, Ref{Int64}, Ptr{UInt32}, Ptr{UInt32}, Ptr{Float64}), h, n, ids, xy, es))
    return ids, [(Int(q & 0xffff), Int(q >> 16)) for q in xy], es
end
"`countdown ./ regrowth_time`, the heat array of plot_abm_model (src/plotting.jl:27-57), from the one-byte grass state"
function heatarray(model)
    h = ENGINES[model].h
    st = Matrix{UInt8}(undef, size(model.space)...)
    check(h, ccall((:abm_get_grass_packed, libabm), Cint, (Ptr{Cvoid}, Ptr{UInt8}), h, st))
    return (st .& 0x7f) ./ model.regrowth_time
end

# ---- several GPUs from this ONE Julia process (abm_group_*, include/abm.h): the GridSpace is cut into row strips, one per
#      device, and stepped by one ccall; arrays go in and come out whole.  group_step!(model, a, m, n; devices = 0:7) ---------
mutable struct GroupEngine
    g::Ptr{Cvoid}
    rules::Tuple{EngineAgentStep,EngineModelStep}
    host_stale::Bool
end
const GROUPS = IdDict{Any,GroupEngine}()
gcheck(g, rc) = rc == 0 || error("libabm group error $rc: ", unsafe_string(ccall((:abm_group_last_error, libabm), Cstring, (Ptr{Cvoid},), g)))

function group_engine!(model, a::EngineAgentStep, m::EngineModelStep, devices)
    ge = get(GROUPS, model, nothing)
    ge !== nothing && ge.rules == (a, m) && return ge
    ge === nothing || (group_sync_host!(model); ccall((:abm_group_destroy, libabm), Cvoid, (Ptr{Cvoid},), ge.g))
    nx, ny = size(model.space);  s = first(allagents(model));  att = prop(model, :attractor, (nx / 2, ny / 2))
    cfg = AbmConfig(nx = nx, ny = ny, periodic = model.space.periodic ? 1 : 0, walk_type = Int32(a.walk_type), eat = a.eat,
                    reproduce = a.reproduce, walkmap_regrowth = m.walkmap, regrowth_time = model.regrowth_time,
                    visual_distance = floor(Int32, s.visual_distance), counter = prop(model, :counter, 50),
                    prob_random_walk = a.prob_random_walk, delta_energy = s.Δenergy, movement_cost = s.movement_cost,
                    reproduction_prob = s.reproduction_prob, attractor_x = att[1], attractor_y = att[2], seed = prop(model, :seed, 321),
                    scheduler = model.scheduler === Schedulers.randomly ? 1 : 0)
    cfg.struct_size = sizeof(AbmConfig)
    devs = Int32.(collect(devices));  g = Ref{Ptr{Cvoid}}(C_NULL)
    rc = ccall((:abm_group_create, libabm), Cint, (Ref{AbmConfig}, Int32, Ptr{Int32}, Ref{Ptr{Cvoid}}), cfg, length(devs), devs, g)
    rc == 0 || error("abm_group_create ($rc): ", last_error(C_NULL))
    if a.walk_type == RANDOM_WALK_GREGARIOUS                 # every member takes the host's exp table
        lut = Float64[exp(-sqrt(k) / 3) for k in 0:2(cfg.visual_distance + 1)^2]
        for i in 0:length(devs)-1
            hm = ccall((:abm_group_member, libabm), Ptr{Cvoid}, (Ptr{Cvoid}, Int32), g[], i)
            check(hm, ccall((:abm_set_weight_lut, libabm), Cint, (Ptr{Cvoid}, Ptr{Float64}, Int32), hm, lut, length(lut)))
        end
    end
    fg = Matrix{UInt8}(model.fully_grown);  cd = Matrix{Int64}(model.countdown)
    wm = hasproperty(model, :land_walkmap) ? Matrix{UInt8}(model.land_walkmap) : nothing
    el = hasproperty(model, :elevation) ? Matrix{Int64}(model.elevation) : nothing
    GC.@preserve fg cd wm el gcheck(g[], ccall((:abm_group_set_grid, libabm), Cint, (Ptr{Cvoid}, Ptr{UInt8}, Ptr{Int64}, Ptr{UInt8}, Ptr{Int64}),
        g[], fg, cd, wm === nothing ? C_NULL : pointer(wm), el === nothing ? C_NULL : pointer(el)))
    ids = sort!(collect(allids(model)))
    xs = Int64[model[i].pos[1] for i in ids];  ys = Int64[model[i].pos[2] for i in ids];  es = Float64[model[i].energy for i in ids]
    gcheck(g[], ccall((:abm_group_set_agents, libabm), Cint, (Ptr{Cvoid}, Int64, Ptr{Int64}, Ptr{Int64}, Ptr{Int64}, Ptr{Float64}, Int64),
                      g[], length(ids), Int64.(ids), xs, ys, es, nextid(model)))
    GROUPS[model] = GroupEngine(g[], (a, m), false)
end

function group_step!(model::ABM, a::EngineAgentStep, m::EngineModelStep, n::Integer = 1; devices = 0:0)
    ge = group_engine!(model, a, m, devices)
    gcheck(ge.g, ccall((:abm_group_step, libabm), Cint, (Ptr{Cvoid}, Int64), ge.g, n))
    ge.host_stale = true
    return model
end

"the all-reduced adata / mdata scalars of a group (`what` as in REDUCE)"
function group_reduce(model, what::Symbol)
    out = Ref{Float64}(0);  g = GROUPS[model].g
    gcheck(g, ccall((:abm_group_reduce, libabm), Cint, (Ptr{Cvoid}, Int32, Ref{Float64}), g, REDUCE[what], out))
    out[]
end

function group_sync_host!(model)
    ge = get(GROUPS, model, nothing)
    (ge === nothing || !ge.host_stale) && return model
    n = Ref{Int64}(0);  gcheck(ge.g, ccall((:abm_group_count_agents, libabm), Cint, (Ptr{Cvoid}, Ref{Int64}), ge.g, n))
    ids = Vector{Int64}(undef, n[]); xs = similar(ids); ys = similar(ids); es = Vector{Float64}(undef, n[])
    gcheck(ge.g, ccall((:abm_group_get_agents, libabm), Cint, (Ptr{Cvoid}, Ref{Int64}, Ptr{Int64}, Ptr{Int64}, Ptr{Int64}, Ptr{Float64}), ge.g, n, ids, xs, ys, es))
    alive = Set(ids);  proto = first(allagents(model))
    for id in collect(allids(model)); id in alive || kill_agent!(model[id], model); end
    for k in eachindex(ids)
        id, pos = ids[k], (Int(xs[k]), Int(ys[k]))
        if haskey(model.agents, id)
            ag = model[id];  ag.pos == pos || move_agent!(ag, pos, model);  ag.energy = es[k]
        else
            ag = deepcopy(proto);  ag.id = id;  ag.pos = pos;  ag.energy = es[k];  add_agent_pos!(ag, model)
        end
    end
    next = Ref{Int64}(0)
    ccall((:abm_group_get_global_state, libabm), Cint, (Ptr{Cvoid}, Ptr{Int64}, Ptr{Int64}, Ptr{Int64}, Ref{Int64}), ge.g, C_NULL, C_NULL, C_NULL, next)
    model.maxid[] = next[] - 1
    fg = Matrix{UInt8}(undef, size(model.space)...);  cd = Matrix{Int64}(undef, size(model.space)...)
    gcheck(ge.g, ccall((:abm_group_get_grid, libabm), Cint, (Ptr{Cvoid}, Ptr{UInt8}, Ptr{Int64}), ge.g, fg, cd))
    model.fully_grown .= fg .== 1;  model.countdown .= cd
    ge.host_stale = false
    return model
end

end # module
