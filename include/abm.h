/*
 * abm.h — C ABI of libabm.so, the B200-native stepping engine for the GridSpace sheep model of
 * tatakof/Animal-Movement-ABM.
 *
 * The reference has no FFI of its own; the boundary it defines is functional: two factories return the
 * callbacks that Agents.jl's step!/run! accept (reference src/agent_actions.jl:24-71 `herbivore_dynamics`,
 * src/model_actions.jl:15-32 `grass_growth_dynamics`).  A host shim (Julia `ccall`, Python `ctypes`) keeps
 * those factory names and turns `step!(model, agent_step!, model_step!, n)` into ONE `abm_step(h, n)`.
 * Each entry point below names the reference interface it replaces.
 *
 * Conventions
 *  - every function returns 0 (ABM_OK) or a negative ABM_E_* code; nothing throws or aborts;
 *  - all arrays are caller-owned; the library copies and never retains a host pointer;
 *  - grid arrays are nx*ny, column-major with x fastest (Julia `Array`): element (x,y) 1-based is
 *    a[(y-1)*nx + (x-1)];
 *  - agent coordinates are 1-based (Julia `agent.pos`);
 *  - a handle is not thread-safe; distinct handles are independent;
 *  - there is no CPU fallback: abm_create fails with ABM_E_CUDA when no sm_100 device is usable.
 *
 * Randomness contract (north star: counter-based Philox keyed by (seed, tick, agent id, draw index)):
 *   u64 draw(seed, tick, id, n) = words (2h, 2h+1), h = n & 1, of
 *       Philox4x32-10(key = (seed_lo, seed_hi), ctr = (n >> 1, tick_lo32, id_lo32, id_hi32))
 *   as lo | hi << 32.  `n` is POSITIONAL: the n-th `rand` call the reference's agent_step! makes for this
 *   (tick, agent), with the two global-RNG call sites (src/agent_actions.jl:213 `sample(1:3)`, :328
 *   `wsample`) routed through the injected RNG.  Float64 = low 52 bits as [1,2) mantissa minus 1
 *   (Julia Random, generic AbstractRNG); `rand(1:k)` = Lemire nearly-division-less (Julia
 *   SamplerRangeNDL), redraws advance n.  Scheduler order: abm_config.scheduler (ABM_SCHED_* below) — by agent id, or
 *   Schedulers.randomly as a Philox-keyed permutation of the ids, fresh every tick.
 */
#ifndef ABM_H_
#define ABM_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ABM_ABI_VERSION 3

/* error codes */
enum {
  ABM_OK = 0,
  ABM_E_BADARG = -1,     /* invalid argument / unsupported parameter combination */
  ABM_E_CUDA = -2,       /* CUDA runtime error or no usable device */
  ABM_E_SMALL = -3,      /* caller buffer too small; needed size returned through the count pointer */
  ABM_E_STATE = -4,      /* call order violated (e.g. abm_step before abm_set_grid) */
  ABM_E_NOWALKABLE = -5, /* random_walkmap found no walkable neighbour: the reference throws (rand(1:0)) */
  ABM_E_CAPACITY = -6,   /* id space (2^32-1) or device memory exhausted */
  ABM_E_NOMEM = -7,
  ABM_E_INTERNAL = -8
};

/* reference: `@enum WalkType ...`, src/agent_actions.jl:1 (same order, same values) */
enum {
  ABM_RANDOM_WALK = 0,
  ABM_RANDOM_WALK_ATTRACTOR = 1,
  ABM_RANDOM_WALK_GREGARIOUS = 2,
  ABM_ALTERNATED_WALK = 3,
  ABM_RANDOM_WALKMAP = 4
};

/* Agents.Schedulers: the order in which the agents of a tick are stepped.  Every reference script selects
 * Schedulers.randomly (scripts/v0.1.jl:63 ... scripts/v0.6.jl:100).  BY_ID is ascending id.  RANDOMLY is a fresh
 * pseudo-random order every tick: ascending ORDER KEY, key = P_t(id), where P_t is a keyed permutation of [0, 2^b): a
 * 4-round balanced Feistel network over b = 2 hb bits, b the smallest even number >= 2 with 2^b >= nextid(model) at the
 * start of the tick; round i maps (L, R) -> (R, L xor (mix(R xor k_i) mod 2^hb)) with mix(z) = z * 0x9E3779B1, z ^= z >> 15,
 * z *= 0x85EBCA77, z ^= z >> 13 (32-bit), and (k_0..k_3) = Philox4x32-10(counter (0, tick, 2^32-1, 2^32-1); key = seed), the
 * stream of the reserved agent id 2^64-1.  (The reference shuffles with its MersenneTwister; the north star replaces that
 * RNG, so the permutation is part of the Philox contract above.)  Offspring receive nextid(model) in the order their
 * parents are stepped; the first agent stepped onto a grown cell eats it. */
enum { ABM_SCHED_BY_ID = 0, ABM_SCHED_RANDOMLY = 1 };

/* Agents.Pathfinding cost metrics used by the scripts (v0.4/v0.5: DirectDistance default,
 * v0.6: PenaltyMap(elevation, MaxDistance{2}()), scripts/v0.6.jl:87-93) */
enum { ABM_ASTAR_DIRECT = 0, ABM_ASTAR_MAX = 1 };

/* what abm_reduce computes (reference: adata/mdata aggregations of run!, scripts/v0.6.jl:204-210) */
enum {
  ABM_REDUCE_COUNT_AGENTS = 0, /* adata = [(sheep, count)] */
  ABM_REDUCE_COUNT_GROWN = 1,  /* mdata = [count_grass] = count(model.fully_grown) */
  ABM_REDUCE_SUM_ENERGY = 2,
  ABM_REDUCE_MAX_ENERGY = 3,
  ABM_REDUCE_MIN_ENERGY = 4
};

/*
 * Flat mirror of the kwargs of `initialize_model` (scripts/v0.1.jl:36-46 ... scripts/v0.6.jl:36-47), of the two
 * factories (src/agent_actions.jl:24-33, src/model_actions.jl:15-17) and of the model `properties`
 * NamedTuple (scripts/v0.6.jl:78-94).  Sheep parameters are model-uniform: every script gives all sheep the
 * same values and offspring copy the parent's (src/agent_actions.jl:175).
 */
typedef struct abm_config {
  int32_t struct_size; /* = sizeof(abm_config); ABI versioning */
  int32_t nx, ny;      /* GridSpace dims (griddims) */
  int32_t periodic;    /* GridSpace(...; periodic) */
  int32_t walk_type;   /* herbivore_dynamics(; walk_type) */
  int32_t eat;         /* herbivore_dynamics(; eat) */
  int32_t reproduce;   /* herbivore_dynamics(; reproduce) */
  int32_t walkmap_regrowth; /* grass_growth_dynamics(; walkmap) */
  int32_t regrowth_time;    /* model.regrowth_time, 1..126 */
  int32_t visual_distance;  /* floor(agent.visual_distance), 1..7 (gregarious only) */
  int32_t counter;          /* model.counter (alternated walk) */
  int32_t randomwalk_ifempty; /* Agents.walk!(...; ifempty) default used by randomwalk!; 1 in Agents 5.14 */
  int32_t astar_metric;     /* ABM_ASTAR_DIRECT / ABM_ASTAR_MAX */
  int32_t astar_penalty;    /* 1: PenaltyMap(elevation, base metric) */
  int32_t device;           /* CUDA device ordinal of this handle */
  int32_t reserved0;        /* 1: force the general stepping path (tests); 0: automatic */
  double prob_random_walk;  /* custom_agent_step!(...; prob_random_walk = 0.9), src/agent_actions.jl:32 */
  double delta_energy;      /* agent.Δenergy */
  double movement_cost;     /* agent.movement_cost */
  double reproduction_prob; /* agent.reproduction_prob */
  double attractor_x, attractor_y; /* model.attractor (griddims ./ 2, scripts/v0.2.jl:57); 2*a must be integral */
  uint64_t seed;            /* Philox key */
  /* Row-strip decomposition (no reference counterpart: the reference is one process).  n_ranks <= 1: this handle holds the
   * whole GridSpace.  n_ranks > 1 (or strip_ny > 0): `ny` is the GLOBAL height and this handle owns the rows
   * strip_y0 .. strip_y0 + strip_ny - 1 (0-based, multiples of 32); rank r's ring neighbours are r-1 (previous rows) and
   * r+1.  In strip mode every grid array passed to / returned by abm_set_grid / abm_get_grid has strip_ny + 64 rows: 32
   * ghost rows of the neighbouring strips above and below (wrapped when periodic; only walkmap and elevation are read
   * there), and agents are exchanged with global 1-based y. */
  int32_t strip_y0, strip_ny, n_ranks, rank;
  int32_t scheduler;        /* ABM_SCHED_BY_ID (default) / ABM_SCHED_RANDOMLY: AgentBasedModel(...; scheduler), scripts/v0.1.jl:63 */
  int32_t reserved1;
} abm_config;

typedef struct abm_handle abm_handle;

/* fills cfg with the defaults of scripts/v0.1.jl:36-46 (80x80, regrowth 30, Δenergy 4, repro 0.001, cost 1) */
void abm_config_default(abm_config* cfg);
/* sizeof(abm_config) of this library: a binding that mirrors the struct by hand (julia/AbmEngine.jl, ctypes) checks its
 * own layout against it before the first abm_create */
int abm_config_size(void);

/* replaces: AgentBasedModel(Sheep, GridSpace(dims; periodic, metric=:chebyshev); properties, rng, scheduler)
 * scripts/v0.1.jl:49-64.
 * ABM_E_BADARG (text through abm_last_error(NULL)) for what the engine does not take: regrowth_time outside 0..126 (the
 * packed grass byte); RANDOM_WALK_GREGARIOUS on a non-periodic space (the reference itself throws there,
 * src/agent_actions.jl:334), with visual_distance outside 1..7, or on a grid narrower than 2*visual_distance + 3 (the
 * neighbour box would wrap onto itself); attractor coordinates that are not multiples of 1/2; counter < 1; a grid whose
 * 32x32-tiled cell count does not fit 32 bits; strip mode with rows or nx that are not multiples of 32, with nx < 64, or
 * with ALTERNATED_WALK. */
int abm_create(const abm_config* cfg, abm_handle** out);
void abm_destroy(abm_handle* h);

/* replaces: model.fully_grown / model.countdown / model.land_walkmap / model.elevation property arrays
 * (scripts/v0.6.jl:78-94).  fully_grown, walkmap: one byte per cell (0/1); walkmap and elevation may be NULL
 * (all walkable / flat); fully_grown and countdown may both be NULL (terrain only: bare grass, countdown 0 — e.g. before
 * abm_init_random).  countdown must be in 0..126, elevation in -32768..32767. */
int abm_set_grid(abm_handle* h, const uint8_t* fully_grown, const int64_t* countdown, const uint8_t* walkmap,
                 const int64_t* elevation);

/* replaces: add_agent!/add_agent_pos! loops (scripts/v0.1.jl:67-70, scripts/v0.6.jl:104-118).  ids unique,
 * 1 <= id < next_id <= 2^32-1; next_id = model.maxid[] + 1 (Agents.nextid). */
int abm_set_agents(abm_handle* h, int64_t n, const int64_t* id, const int64_t* x, const int64_t* y,
                   const double* energy, int64_t next_id);

/* replaces: the random part of `initialize_model` (scripts/v0.1.jl:66-79: `n_sheep` sheep, each with
 * `energy = rand(1:Δenergy*5) - 1` at a random position; every cell `fully_grown = rand(Bool)`,
 * `countdown = fully_grown ? regrowth_time : rand(1:regrowth_time) - 1`; scripts/v0.5.jl:100-122: sheep at
 * `random_walkable` positions, grass only on walkable cells) — built on the device, so a 2^28-sheep world needs no host
 * arrays.  Terrain, if any, must have been given before through abm_set_grid (its grass arguments are overwritten
 * here); without a previous abm_set_grid every cell is walkable.  Sheep get ids 1..n_sheep, next_id = n_sheep + 1.
 * The reference draws all of this from one sequential MersenneTwister stream; here every sheep and every cell has its
 * own stream of the contract above — sheep i: draw(init_seed, tick = 2^32-1, id = i, n) with n = 0 the energy, then
 * (x, y) pairs until the cell is walkable; cell (x, y): draw(init_seed, tick = 2^32-2, id = (y-1)*nx + (x-1), n) with
 * n = 0 the grown bit (lowest bit of the word), n = 1.. the countdown — so the world does not depend on who builds it.
 * Strip handles: every rank calls it with the same arguments (n_sheep = the population of the WHOLE GridSpace); each rank
 * goes through every sheep's draws and keeps those that land on its rows, so the strips together hold exactly the world a
 * single handle would build.  random_walkable rejects positions anywhere in the space, so a strip with terrain must have
 * been given the whole walkmap before (abm_set_global_walkmap_bits); without terrain nothing else is needed. */
int abm_init_random(abm_handle* h, int64_t n_sheep, uint64_t init_seed);

/* Strip handles only: the walkmap of the whole GridSpace, one bit per cell — bit (l & 31) of word (l >> 5),
 * l = (y-1)*nx + (x-1) — for abm_init_random (32768^2 cells are 128 MiB).  NULL forgets it. */
int abm_set_global_walkmap_bits(abm_handle* h, const uint32_t* bits);

/* replaces: model.behav[1], model.behav_counter[1] (scripts/v0.4.jl:69-71,97) and the tick the RNG is keyed by */
int abm_set_global_state(abm_handle* h, int64_t tick, int64_t behav, int64_t behav_counter);
int abm_get_global_state(abm_handle* h, int64_t* tick, int64_t* behav, int64_t* behav_counter, int64_t* next_id);

/* w[k] = exp(-sqrt(k)/3), k = 0 .. 2*(visual_distance+1)^2, computed by the HOST's exp so that the device
 * only does IEEE add/div/compare (reference: f(x) = exp(-x / 3), src/agent_actions.jl:319).  When never
 * called, the library fills the table with the C library's exp. */
int abm_set_weight_lut(abm_handle* h, const double* w, int32_t n);

/* replaces: Agents.step!(model, agent_step!, model_step!, n) — n ticks of { agent_step! for every agent in id
 * order (src/agent_actions.jl:29-69); model_step! (src/model_actions.jl:18-30) }.  Synchronous. */
int abm_step(abm_handle* h, int64_t n_ticks);

/* replaces: nagents(model) */
int abm_count_agents(abm_handle* h, int64_t* n);

/* replaces: reading model[id].pos / .energy for all agents, in id order.  *n_inout: capacity in, count out. */
int abm_get_agents(abm_handle* h, int64_t* n_inout, int64_t* id, int64_t* x, int64_t* y, double* energy);

/* replaces: reading model.fully_grown / model.countdown (e.g. src/plotting.jl:27-57).  Either may be NULL. */
int abm_get_grid(abm_handle* h, uint8_t* fully_grown, int64_t* countdown);

/* Narrow forms of the four state transfers above, for hosts that round-trip the model every tick (abmplot/abmvideo step
 * one tick per frame, scripts/v0.3.jl:124): 1 byte per cell and 16 bytes per agent instead of 9 and 32.
 *   grass byte  = fully_grown << 7 | countdown (0..126)  — model.fully_grown[x,y] and model.countdown[x,y] in one byte;
 *                 the terrain (walkmap, elevation) given to abm_set_grid is kept; without a previous abm_set_grid every
 *                 cell is walkable and flat;
 *   agent       = id (uint32), xy = x | y << 16 (1-based agent.pos; grid dims <= 65535), energy (double).
 * Same layouts (x fastest; strip handles: strip_ny + 64 rows, global y), same validation, same id order on the way out. */
int abm_set_grass_packed(abm_handle* h, const uint8_t* state);
int abm_get_grass_packed(abm_handle* h, uint8_t* state);
int abm_set_agents_packed(abm_handle* h, int64_t n, const uint32_t* id, const uint32_t* xy, const double* energy,
                          int64_t next_id);
int abm_get_agents_packed(abm_handle* h, int64_t* n_inout, uint32_t* id, uint32_t* xy, double* energy);

/* replaces: adata/mdata aggregation inside Agents.run! (scripts/v0.6.jl:204-210).  On a strip handle the call is
 * collective (every rank makes it) and returns the value over the WHOLE GridSpace, combined in rank order. */
int abm_reduce(abm_handle* h, int32_t what, double* out);

/* replaces: filtered adata tuples of run!, `(property, aggregator, filter)` (Agents.jl; naming at src/agent_actions.jl:964),
 * for filters that are boxes: agents with energy in [energy_lo, energy_hi] standing on x in [x_lo, x_hi], y in [y_lo, y_hi]
 * (1-based, inclusive; use -inf / +inf and 1 / nx, ny for "any").  `what` as for abm_reduce (COUNT_GROWN is not an agent
 * reduction: ABM_E_BADARG); min / max of an empty selection give 0.  Collective on strip handles, like abm_reduce. */
typedef struct abm_filter { double energy_lo, energy_hi; int32_t x_lo, x_hi, y_lo, y_hi; } abm_filter;
int abm_reduce_where(abm_handle* h, int32_t what, const abm_filter* f, double* out);

/* replaces: the heat arrays the reference's plots draw, computed on the device and returned as nx*ny floats (x fastest):
 *   which = 0: `model.countdown ./ model.regrowth_time` (grasscolor / heatarray of plot_abm_model, src/plotting.jl:27-57)
 *   which = 1: `penaltymap(model.pathfinder)` = the elevation the PenaltyMap was built from (scripts/v0.6.jl:179)
 * Strip handles: strip_ny + 64 rows, like abm_get_grid. */
int abm_get_heat(abm_handle* h, int32_t which, float* out);

/* replaces: plan_route!(agent, dest, model.pathfinder) for a batch of agents (src/agent_actions.jl:217-221;
 * Agents.Pathfinding.find_path).  goal coordinates 1-based.  Existing routes of these agents are overwritten;
 * an unreachable goal leaves the agent without a route.  A search whose open list outgrows its slot is run again with a
 * larger one; beyond 2^20 open entries the call fails with ABM_E_CAPACITY.  Periodic grids need nx, ny >= 3 for A* (below,
 * two of the 8 neighbour offsets wrap onto one cell): ABM_E_BADARG here, and from abm_create for ALTERNATED_WALK. */
int abm_plan_routes(abm_handle* h, int64_t b, const int64_t* agent_id, const int64_t* goal_x,
                    const int64_t* goal_y);

/* replaces: model.pathfinder.agent_paths[id] (Agents.Pathfinding).  CSR output: path of agent_id[i] is
 * (px,py)[offset[i] .. offset[i+1]), 1-based cells, first element = next cell to move to; the start cell is
 * excluded.  offset has b+1 entries; *cells_inout: capacity of px/py in, total cells out.  cost (nullable)
 * receives the g-cost of the goal at planning time, or -1 when there is no route. */
int abm_get_paths(abm_handle* h, int64_t b, const int64_t* agent_id, int64_t* offset, int64_t* cells_inout,
                  int64_t* px, int64_t* py, int64_t* cost);

/* device timing of the last abm_step in milliseconds: out[0] total, out[1] agent phase, out[2] grass phase,
 * out[3] resolution rounds (count), out[4] kernels launched, out[5] A* expansions performed by the last plan call
 * (a goal outside the start's walkmap component is answered without searching and costs none),
 * out[6] agent-updates performed (sum over ticks of the population at tick start) */
int abm_get_timing(abm_handle* h, double* out, int32_t n);

/* no reference counterpart (the reference has no profiler): when on, every kernel of abm_step is bracketed by CUDA
 * events on the handle's stream; abm_get_kernel_stat(idx = 0, 1, ...) then returns per-kernel totals of the last
 * abm_step (ABM_E_SMALL past the last kernel).  `units` = agents or cells the launches processed. */
int abm_set_profiling(abm_handle* h, int32_t on);
int abm_get_kernel_stat(abm_handle* h, int32_t idx, char* name, int32_t name_cap, double* ms, int64_t* launches,
                        int64_t* units);

/* ---- multi-GPU: one process per GPU, one handle per process; abm_step performs the per-tick halo exchange itself ---- */

/* Transport supplied by the host (MPI, gloo, ...): all pointers are HOST memory, calls are collective over the ring of
 * n_ranks handles and return 0 on success.  sendrecv: `to_prev` goes to rank-1 (who receives it as `from_next`), `to_next`
 * to rank+1 (who receives it as `from_prev`); with n_ranks == 2 both neighbours are the same rank and the `to_prev`
 * message must be matched first. */
typedef struct abm_transport {
  void* user;
  int (*sendrecv)(void* user, const void* to_prev, uint64_t to_prev_bytes, const void* to_next, uint64_t to_next_bytes,
                  void* from_prev, uint64_t from_prev_bytes, void* from_next, uint64_t from_next_bytes);
  int (*allreduce_sum_u64)(void* user, uint64_t* values, int32_t n);
  int (*allgather_u32)(void* user, const uint32_t* send, uint32_t* recv /* n_ranks * n */, uint32_t n);
} abm_transport;
int abm_set_transport(abm_handle* h, const abm_transport* t);

/* NCCL transport over NVLink (device buffers, no host staging): rank 0 calls abm_nccl_unique_id, the host ships the 128
 * bytes to every rank (e.g. torch.distributed.broadcast), every rank calls abm_comm_init_nccl. */
int abm_nccl_unique_id(void* out128);
int abm_comm_init_nccl(abm_handle* h, const void* unique_id128);

/* Peer-memory transport for the GPUs of one node (NVLink / NVSwitch): kernels store the boundary rows, decision bytes and
 * the small per-round counters straight into the neighbours' "mailboxes" and publish sequence numbers there — no NCCL
 * call, no host staging.  After abm_set_agents every rank calls abm_p2p_export (a 64-byte CUDA IPC handle of its
 * mailbox), the host all-gathers the handles (n_ranks x 64 bytes, rank order) and every rank calls abm_p2p_connect.
 * A neighbour that does not deliver within a few seconds makes abm_step fail with ABM_E_CUDA instead of hanging. */
int abm_p2p_export(abm_handle* h, void* handle64);
int abm_p2p_connect(abm_handle* h, const void* handles, int32_t n_ranks);

/* ---- multi-GPU from ONE process: the reference's caller is a single Julia process (scripts/v0.3.jl:124, abmvideo ->
 * step!(model, agent_step!, model_step!, n)) and cannot start a process per GPU.  An abm_group owns n_devices strip handles
 * (whole 32-row core rows, as even as possible) of the GridSpace described by cfg (strip fields ignored) and drives them
 * from the calling thread: each call below fans out to one worker thread per device and returns when all are done.  The
 * strips exchange their boundaries over NVLink peer memory mapped inside the process (cudaDeviceEnablePeerAccess), or
 * through an in-process host mailbox when the devices cannot reach each other.  All arrays are those of the WHOLE
 * GridSpace, exactly as for a single handle; device_ids == NULL means devices 0 .. n_devices-1 (all distinct). ---- */
typedef struct abm_group abm_group;
int abm_group_create(const abm_config* cfg, int32_t n_devices, const int32_t* device_ids, abm_group** out);
void abm_group_destroy(abm_group* g);
int abm_group_set_grid(abm_group* g, const uint8_t* fully_grown, const int64_t* countdown, const uint8_t* walkmap,
                       const int64_t* elevation);
int abm_group_set_agents(abm_group* g, int64_t n, const int64_t* id, const int64_t* x, const int64_t* y,
                         const double* energy, int64_t next_id);
int abm_group_init_random(abm_group* g, int64_t n_sheep, uint64_t init_seed);
int abm_group_set_global_state(abm_group* g, int64_t tick, int64_t behav, int64_t behav_counter);
int abm_group_get_global_state(abm_group* g, int64_t* tick, int64_t* behav, int64_t* behav_counter, int64_t* next_id);
int abm_group_step(abm_group* g, int64_t n_ticks);            /* replaces: Agents.step!(model, agent_step!, model_step!, n) */
int abm_group_count_agents(abm_group* g, int64_t* n);
int abm_group_get_agents(abm_group* g, int64_t* n_inout, int64_t* id, int64_t* x, int64_t* y, double* energy); /* id order */
int abm_group_get_grid(abm_group* g, uint8_t* fully_grown, int64_t* countdown);
int abm_group_reduce(abm_group* g, int32_t what, double* out);
abm_handle* abm_group_member(abm_group* g, int32_t i);        /* strip i's handle (profiling, timing); owned by the group */
const char* abm_group_last_error(abm_group* g);

/* message for the last failing call on h (or on creation when h == NULL); valid until the next call */
const char* abm_last_error(abm_handle* h);

#ifdef __cplusplus
}
#endif
#endif /* ABM_H_ */
