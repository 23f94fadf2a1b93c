# tools/dump_reference.jl — run the UNMODIFIED rules of tatakof/Animal-Movement-ABM under the randomness contract of
# include/abm.h and dump every tick, so that the CPU oracle (and through it the CUDA engine) can be pinned against the
# real reference.
#
# NOT EXECUTED in the build image (no Julia there): written against Julia 1.9.1 + Agents 5.14.0 + StatsBase 0.33.21 as
# pinned by the reference's Manifest.toml.  Usage, on a box that has them:
#
#   python tools/export_cases.py cases_dir                      # initial worlds of tests/cases.py, plain CSV
#   julia --project=/path/to/Animal-Movement-ABM tools/dump_reference.jl /path/to/Animal-Movement-ABM cases_dir
#   python tools/compare_reference_dump.py cases_dir            # oracle vs the dumps, tick by tick, bit-exact
#
# What is changed relative to the reference, and nothing else (SURVEY.md §0.7, INTEGRATION.md §2):
#   1. `model.rng` is a Philox4x32-10 stream keyed (seed; tick, agent id, positional draw index), reset before every
#      agent step; the scheduler is `Schedulers.ByID()`;
#   2. the two global-RNG call sites are routed through `model.rng` (text substitution below, on the in-memory copy of
#      src/agent_actions.jl: `sample(1:3)` -> `rand(model.rng, 1:3)`, `wsample(` -> `wsample(model.rng, `).
# Only lines 1-336 of src/agent_actions.jl (the GridSpace rules) and 1-32 of src/model_actions.jl are loaded: the rest is
# ContinuousSpace / GUI code that needs Makie.

using Agents, Random, StatsBase, Distributions, LinearAlgebra, DelimitedFiles

# ------------------------------------------------------------------------------------------------ Philox4x32-10
const PHILOX_M0 = 0xD2511F53
const PHILOX_M1 = 0xCD9E8D57
const PHILOX_W0 = 0x9E3779B9
const PHILOX_W1 = 0xBB67AE85

@inline function mulhilo(a::UInt32, b::UInt32)
    p = UInt64(a) * UInt64(b)
    return (p >> 32) % UInt32, p % UInt32
end

function philox4x32_10(c::NTuple{4,UInt32}, k::NTuple{2,UInt32})
    c0, c1, c2, c3 = c
    k0, k1 = k
    for _ in 1:10
        hi0, lo0 = mulhilo(PHILOX_M0, c0)
        hi1, lo1 = mulhilo(PHILOX_M1, c2)
        c0, c1, c2, c3 = hi1 ⊻ c1 ⊻ k0, lo1, hi0 ⊻ c3 ⊻ k1, lo0
        k0 += PHILOX_W0   # UInt32 arithmetic wraps
        k1 += PHILOX_W1
    end
    return (c0, c1, c2, c3)
end

# known answer (Random123 kat_vectors): ctr = key = all ones -> 408f276d 41c83b0e a20bc7c6 6d5451fd
@assert philox4x32_10((0xffffffff, 0xffffffff, 0xffffffff, 0xffffffff), (0xffffffff, 0xffffffff)) ==
        (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)

"""u64 draw(seed, tick, id, n) of include/abm.h: words (2h, 2h+1), h = n & 1, of the block with counter n >> 1."""
mutable struct PhiloxRNG <: AbstractRNG
    seed::UInt64
    tick::UInt64
    id::UInt64
    n::UInt64
end
reset!(r::PhiloxRNG, tick, id) = (r.tick = UInt64(tick); r.id = UInt64(id); r.n = 0; r)

function Random.rand(r::PhiloxRNG, ::Random.SamplerType{UInt64})
    ctr = (((r.n >> 1) % UInt32), r.tick % UInt32, r.id % UInt32, (r.id >> 32) % UInt32)
    key = (r.seed % UInt32, (r.seed >> 32) % UInt32)
    out = philox4x32_10(ctr, key)
    h = Int(r.n & 1)
    r.n += 1
    return UInt64(out[2h + 1]) | (UInt64(out[2h + 2]) << 32)
end
# Float64 and 1:k go through Julia's generic AbstractRNG paths: low 52 bits as [1, 2) mantissa minus 1, SamplerRangeNDL.
Random.rng_native_52(::PhiloxRNG) = UInt64

# ------------------------------------------------------------------------------------------------ the reference's rules
refdir = ARGS[1]
casesdir = ARGS[2]

function load_rules(path, lastline)
    lines = readlines(path)[1:lastline]
    text = join(lines, "\n")
    text = replace(text, "sample(1:3)" => "rand(model.rng, 1:3)")
    text = replace(text, "wsample(" => "wsample(model.rng, ")
    include_string(Main, text, path)
end
load_rules(joinpath(refdir, "src", "agent_actions.jl"), 336)
load_rules(joinpath(refdir, "src", "model_actions.jl"), 32)

@agent Sheep GridAgent{2} begin   # scripts/v0.1.jl:28-34 (v0.3+ add visual_distance)
    energy::Float64
    reproduction_prob::Float64
    Δenergy::Float64
    movement_cost::Float64
    visual_distance::Float64
end

# ------------------------------------------------------------------------------------------------ one case
readgrid(path) = permutedims(readdlm(path, ',', Int))   # file: ny rows of nx values -> Julia (nx, ny), x fastest

function read_config(path)
    cfg = Dict{String,Float64}()
    for line in eachline(path)
        isempty(strip(line)) && continue
        k, v = split(line, '=')
        cfg[strip(k)] = parse(Float64, strip(v))
    end
    return cfg
end

function run_case(dir)
    cfg = read_config(joinpath(dir, "config.txt"))
    geti(k) = Int(cfg[k])
    nx, ny = geti("nx"), geti("ny")
    periodic = geti("periodic") == 1
    space = GridSpace((nx, ny); periodic, metric = :chebyshev)
    fully_grown = BitMatrix(readgrid(joinpath(dir, "fully_grown.csv")) .== 1)
    countdown = readgrid(joinpath(dir, "countdown.csv"))
    properties = Dict{Symbol,Any}(
        :fully_grown => fully_grown, :countdown => countdown, :regrowth_time => geti("regrowth_time"),
        :attractor => (cfg["attractor_x"], cfg["attractor_y"]), :counter => geti("counter"),
        :behav => [geti("behav")], :behav_counter => [geti("behav_counter")])
    if isfile(joinpath(dir, "walkmap.csv"))
        walkmap = BitMatrix(readgrid(joinpath(dir, "walkmap.csv")) .== 1)
        elevation = readgrid(joinpath(dir, "elevation.csv"))
        properties[:land_walkmap] = walkmap
        properties[:elevation] = elevation
        base = geti("astar_metric") == 1 ? MaxDistance{2}() : DirectDistance{2}()
        metric = geti("astar_penalty") == 1 ? PenaltyMap(elevation, base) : base
        properties[:pathfinder] = AStar(space; walkmap, cost_metric = metric)   # scripts/v0.4.jl:67, v0.6.jl:87-93
    end
    rng = PhiloxRNG(UInt64(cfg["seed"]), 0, 0, 0)
    model = ABM(Sheep, space; properties, rng, scheduler = Schedulers.ByID())
    agents = readdlm(joinpath(dir, "agents.csv"), ',', Float64; skipstart = 1)
    for r in eachrow(agents)   # id, x, y, energy — in id order, as abm_set_agents requires
        add_agent_pos!(Sheep(Int(r[1]), (Int(r[2]), Int(r[3])), r[4], cfg["reproduction_prob"], cfg["delta_energy"],
                             cfg["movement_cost"], Float64(geti("visual_distance"))), model)
    end

    if geti("walk_type") == 2   # the HOST's exp decides the wsample index (include/abm.h, abm_set_weight_lut): hand Julia's to the oracle
        r = geti("visual_distance")
        open(joinpath(dir, "lut.csv"), "w") do io
            for k in 0:2(r + 1)^2
                println(io, repr(exp(-sqrt(k) / 3)))
            end
        end
    end

    tick = Ref(0)
    inner = herbivore_dynamics(; walk_type = WalkType(geti("walk_type")), eat = geti("eat") == 1, reproduce = geti("reproduce") == 1)
    prob = cfg["prob_random_walk"]
    agent_step!(agent, model) = (reset!(model.rng, tick[], agent.id); inner(agent, model; prob_random_walk = prob))
    grass! = grass_growth_dynamics(; walkmap = geti("walkmap_regrowth") == 1)
    model_step!(model) = (grass!(model); tick[] += 1)

    for t in 1:geti("ticks")
        step!(model, agent_step!, model_step!, 1)
        ids = sort!(collect(allids(model)))
        open(joinpath(dir, "ref_tick_$(lpad(t, 4, '0'))_agents.csv"), "w") do io
            println(io, "id,x,y,energy")
            for i in ids
                a = model[i]
                println(io, i, ",", a.pos[1], ",", a.pos[2], ",", repr(a.energy))
            end
        end
        writedlm(joinpath(dir, "ref_tick_$(lpad(t, 4, '0'))_fully_grown.csv"), permutedims(Int.(model.fully_grown)), ',')
        writedlm(joinpath(dir, "ref_tick_$(lpad(t, 4, '0'))_countdown.csv"), permutedims(model.countdown), ',')
        open(joinpath(dir, "ref_tick_$(lpad(t, 4, '0'))_state.txt"), "w") do io
            println(io, "behav=", model.behav[1]); println(io, "behav_counter=", model.behav_counter[1])
        end
    end
end

for name in sort(readdir(casesdir))
    dir = joinpath(casesdir, name)
    isfile(joinpath(dir, "config.txt")) || continue
    println("case ", name)
    run_case(dir)
end
