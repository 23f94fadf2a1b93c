/*
 * abm_oracle.cpp — TEST INFRASTRUCTURE ONLY.  Sequential CPU restatement of the per-tick hot path of
 * tatakof/Animal-Movement-ABM (reference src/agent_actions.jl:1-336, src/model_actions.jl:15-32) and of the
 * Agents.jl 5.14.0 / StatsBase 0.33.21 / Julia 1.9.1 Random pieces those rules call.
 *
 * PARITY UNPINNED: the reference's only test is `@test 1 == 1` (test/runtests.jl:11-13), there are no golden
 * vectors, Julia is not installed here, and Agents.jl is not vendored under /root/reference.  The third-party
 * semantics below are restated from their published behaviour; each restated function names its upstream.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may use this file.
 * The shipped library (libabm.so) never links or calls it.
 *
 * The data structures deliberately mirror Agents.jl (id-keyed agent table, per-cell id vectors in insertion
 * order, strictly sequential agent loop) so that every ordering effect of the reference is reproduced rather
 * than re-derived.  Nothing here is shared with the CUDA engine — not even the Philox code — so that the two
 * implementations cross-check each other.
 */
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <deque>
#include <map>
#include <queue>
#include <set>
#include <string>
#include <tuple>
#include <unordered_map>
#include <unordered_set>
#include <vector>

#include "../include/abm.h"

namespace {

typedef int64_t i64;
typedef uint64_t u64;
typedef uint32_t u32;

/* ---- Philox4x32-10 (Salmon et al. 2011; constants as in curand_philox4x32_x.h:88-91) ------------------- */
static inline void philox_round(u32 c[4], const u32 k[2]) {
  const u64 p0 = (u64)0xD2511F53u * c[0];
  const u64 p1 = (u64)0xCD9E8D57u * c[2];
  const u32 n0 = (u32)(p1 >> 32) ^ c[1] ^ k[0];
  const u32 n1 = (u32)p1;
  const u32 n2 = (u32)(p0 >> 32) ^ c[3] ^ k[1];
  const u32 n3 = (u32)p0;
  c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
}
static inline void philox4x32_10(const u32 ctr[4], const u32 key[2], u32 out[4]) {
  u32 c[4] = {ctr[0], ctr[1], ctr[2], ctr[3]};
  u32 k[2] = {key[0], key[1]};
  for (int r = 0; r < 10; ++r) {
    philox_round(c, k);
    k[0] += 0x9E3779B9u;
    k[1] += 0xBB67AE85u;
  }
  out[0] = c[0]; out[1] = c[1]; out[2] = c[2]; out[3] = c[3];
}

/* The injected RNG of the parity harness: an AbstractRNG that counts `rand(rng, UInt64)` calls since it was
 * reset for (tick, agent id).  include/abm.h states the contract. */
struct AgentRng {
  u64 seed = 0, tick = 0, id = 0, n = 0;
  void reset(u64 t, u64 i) { tick = t; id = i; n = 0; }
  u64 next_u64() {
    const u32 ctr[4] = {(u32)(n >> 1), (u32)tick, (u32)id, (u32)(id >> 32)};
    const u32 key[2] = {(u32)seed, (u32)(seed >> 32)};
    u32 o[4];
    philox4x32_10(ctr, key, o);
    const int h = (int)(n & 1);
    ++n;
    return (u64)o[2 * h] | ((u64)o[2 * h + 1] << 32);
  }
  /* Julia 1.9.1 Random: rand(r::AbstractRNG, ::SamplerTrivial{CloseOpen12_64}) =
   * reinterpret(Float64, 0x3ff0000000000000 | rand(r, UInt64) & 0x000fffffffffffff); CloseOpen01 = that - 1.0 */
  double rand_f64() {
    const u64 bits = 0x3ff0000000000000ull | (next_u64() & 0x000fffffffffffffull);
    double d;
    std::memcpy(&d, &bits, 8);
    return d - 1.0;
  }
  /* Julia 1.9.1 Random: rand(rng, 1:s) for Int64 ranges -> SamplerRangeNDL (Lemire nearly-division-less). */
  i64 rand_1_to(u64 s) {
    unsigned __int128 m = (unsigned __int128)next_u64() * s;
    u64 l = (u64)m;
    if (l < s) {
      const u64 t = (0 - s) % s;
      while (l < t) {
        m = (unsigned __int128)next_u64() * s;
        l = (u64)m;
      }
    }
    return (i64)(u64)(m >> 64) + 1;
  }
};

struct Sheep {
  i64 id;
  int x, y; /* 1-based, as agent.pos */
  double energy;
};

typedef std::pair<int, int> Pos;

struct Model {
  abm_config cfg;
  int nx = 0, ny = 0;
  bool periodic = false;
  std::map<i64, Sheep> agents;            /* model.agents :: Dict{Int,Sheep}; Schedulers.by_id = sorted keys */
  std::vector<std::vector<i64>> stored;   /* GridSpace.stored_ids[x,y] :: Vector{Int}, insertion order */
  std::vector<uint8_t> fully_grown;       /* model.fully_grown */
  std::vector<i64> countdown;             /* model.countdown */
  std::vector<uint8_t> walkmap;           /* model.land_walkmap (all true when absent) */
  std::vector<i64> elevation;             /* model.elevation (all 0 when absent) */
  i64 behav = 0, behav_counter = 0;       /* model.behav[1], model.behav_counter[1] */
  i64 maxid = 0;                          /* model.maxid[] ; nextid = maxid + 1 */
  u64 tick = 0;
  std::map<i64, std::deque<Pos>> paths;   /* pathfinder.agent_paths */
  std::map<i64, i64> path_cost;           /* bookkeeping for tests: g(goal) at planning time */
  std::vector<double> lut;                /* optional host-supplied exp(-sqrt(k)/3) table */
  AgentRng rng;
  bool grid_set = false;
  /* alternative readings of the third-party semantics that SURVEY.md Appendix A could not settle without running Julia
   * (confidence M / L): bit mask set through oracle_set_variants (tests, tools/compare_reference_dump.py --variants); 0 = the
   * reading the engine implements */
  uint32_t variants = 0;
  u64 astar_expansions = 0;
  std::string err;

  size_t lin(int x, int y) const { return (size_t)(y - 1) * nx + (size_t)(x - 1); }
};

/* ---- Agents.jl 5.14.0 GridSpace primitives ---------------------------------------------------------------- */

static inline int mod1(int a, int n) { /* Julia mod1 */
  int r = a % n;
  if (r <= 0) r += n;
  return r;
}

/* offsets_at_radius(model, 1) == offsets_within_radius_no_0(model, 1): Iterators.product(-1:1, -1:1), x fastest,
 * (0,0) removed. */
static const int OFF8[8][2] = {{-1, -1}, {0, -1}, {1, -1}, {-1, 0}, {1, 0}, {-1, 1}, {0, 1}, {1, 1}};
/* the other reading of A-5 / A-8 (order marked M): the second coordinate fastest */
static const int OFF8_YFAST[8][2] = {{-1, -1}, {-1, 0}, {-1, 1}, {0, -1}, {0, 1}, {1, -1}, {1, 0}, {1, 1}};
enum { VAR_OFFSETS_Y_FASTEST = 1, VAR_RANDOM_POSITION_Y_FIRST = 2, VAR_ASTAR_NEIGHBOURS_Y_FASTEST = 4, VAR_ALL = 7 };
static const int (*offsets8(const Model& m))[2] { return (m.variants & VAR_OFFSETS_Y_FASTEST) ? OFF8_YFAST : OFF8; }

static void remove_from_cell(Model& m, i64 id, int x, int y) {
  std::vector<i64>& v = m.stored[m.lin(x, y)];
  std::vector<i64>::iterator it = std::find(v.begin(), v.end(), id); /* findfirst + deleteat! */
  if (it != v.end()) v.erase(it);
}

/* move_agent!(agent, pos, model) */
static void move_agent(Model& m, Sheep& a, int x, int y) {
  remove_from_cell(m, a.id, a.x, a.y);
  a.x = x;
  a.y = y;
  m.stored[m.lin(x, y)].push_back(a.id);
}

/* walk!(agent, direction, model; ifempty) */
static void walk(Model& m, Sheep& a, int dx, int dy, bool ifempty) {
  int tx, ty;
  if (m.periodic) {
    tx = mod1(a.x + dx, m.nx);
    ty = mod1(a.y + dy, m.ny);
  } else {
    tx = std::min(std::max(a.x + dx, 1), m.nx);
    ty = std::min(std::max(a.y + dy, 1), m.ny);
  }
  if (ifempty) {
    if (m.stored[m.lin(tx, ty)].empty()) move_agent(m, a, tx, ty); /* isempty(target, model) && move_agent! */
  } else {
    move_agent(m, a, tx, ty);
  }
}

/* randomwalk!(agent, model) = walk!(agent, rand(abmrng(model), offsets_at_radius(model, 1)), model) */
static void randomwalk(Model& m, Sheep& a) {
  const i64 k = m.rng.rand_1_to(8);
  walk(m, a, offsets8(m)[k - 1][0], offsets8(m)[k - 1][1], m.cfg.randomwalk_ifempty != 0);
}

/* nearby_positions(pos, model, 1): wrapped when periodic, bounds-filtered otherwise */
static std::vector<Pos> nearby_positions(const Model& m, int x, int y) {
  std::vector<Pos> out;
  for (int i = 0; i < 8; ++i) {
    int px = x + offsets8(m)[i][0], py = y + offsets8(m)[i][1];
    if (m.periodic) {
      out.push_back(Pos(mod1(px, m.nx), mod1(py, m.ny)));
    } else if (px >= 1 && px <= m.nx && py >= 1 && py <= m.ny) {
      out.push_back(Pos(px, py));
    }
  }
  return out;
}

/* euclidean_distance(a, b, model): squared, as an exact integer; sqrt applied by the caller */
static i64 dist2(const Model& m, int ax, int ay, int bx, int by) {
  i64 dx = std::abs(ax - bx), dy = std::abs(ay - by);
  if (m.periodic) {
    dx = std::min<i64>(dx, m.nx - dx);
    dy = std::min<i64>(dy, m.ny - dy);
  }
  return dx * dx + dy * dy;
}

/* ---- Agents.Pathfinding 5.14.0 ----------------------------------------------------------------------------- */

static i64 base_metric(const Model& m, Pos a, Pos b) {
  i64 dx = std::abs(a.first - b.first), dy = std::abs(a.second - b.second);
  if (m.periodic) {
    dx = std::min<i64>(dx, m.nx - dx);
    dy = std::min<i64>(dy, m.ny - dy);
  }
  if (m.cfg.astar_metric == ABM_ASTAR_MAX) return std::max(dx, dy);      /* MaxDistance */
  return 10 * (std::max(dx, dy) - std::min(dx, dy)) + 14 * std::min(dx, dy); /* DirectDistance, costs [10,14] */
}
/* delta_cost: PenaltyMap(pmap, base) = base + |pmap[a] - pmap[b]| — edge cost AND heuristic */
static i64 delta_cost(const Model& m, Pos a, Pos b) {
  i64 c = base_metric(m, a, b);
  if (m.cfg.astar_penalty) c += std::llabs(m.elevation[m.lin(a.first, a.second)] - m.elevation[m.lin(b.first, b.second)]);
  return c;
}

/* find_path(pathfinder, from, to): binary min-heap of (f, pos) tuples with lazy duplicates, closed set, parent
 * updated on strict g improvement; path excludes the start cell.  Returns false when unreachable. */
static bool find_path(Model& m, Pos from, Pos to, std::deque<Pos>& path, i64* cost) {
  typedef std::tuple<i64, int, int> Node; /* (f, x, y): tuple order = Julia's isless on (Int, Dims{2}) */
  std::priority_queue<Node, std::vector<Node>, std::greater<Node>> open;
  const size_t G = (size_t)m.nx * m.ny;
  std::unordered_map<size_t, i64> g;       /* grid :: DefaultDict{Dims,GridCell}; missing = typemax */
  std::unordered_map<size_t, size_t> parent;
  std::unordered_set<size_t> closed;
  (void)G;
  const size_t lf = m.lin(from.first, from.second), lt = m.lin(to.first, to.second);
  g[lf] = 0;
  /* GridCell(g, h, admissibility = 0.0).f = round(Int, g + (1 + 0.0) * h) = g + h */
  open.push(Node(0 + delta_cost(m, from, to), from.first, from.second));
  while (!open.empty()) {
    const Node top = open.top();
    open.pop();
    const Pos cur(std::get<1>(top), std::get<2>(top));
    if (cur == to) break;
    const size_t lc = m.lin(cur.first, cur.second);
    closed.insert(lc);
    ++m.astar_expansions;
    for (int i = 0; i < 8; ++i) { /* moore_neighborhood, product order */
      const int (*nb8)[2] = (m.variants & VAR_ASTAR_NEIGHBOURS_Y_FASTEST) ? OFF8_YFAST : OFF8;
      int px = cur.first + nb8[i][0], py = cur.second + nb8[i][1];
      if (m.periodic) {
        px = mod1(px, m.nx);
        py = mod1(py, m.ny);
      } else if (px < 1 || px > m.nx || py < 1 || py > m.ny) {
        continue;
      }
      const size_t ln = m.lin(px, py);
      if (!m.walkmap[ln] || closed.count(ln)) continue;
      const i64 ng = g[lc] + delta_cost(m, cur, Pos(px, py));
      std::unordered_map<size_t, i64>::iterator it = g.find(ln);
      if (it == g.end() || ng < it->second) {
        parent[ln] = lc;
        g[ln] = ng;
        open.push(Node(ng + delta_cost(m, Pos(px, py), to), px, py));
      }
    }
  }
  path.clear();
  size_t cur = lt;
  while (true) {
    std::unordered_map<size_t, size_t>::iterator it = parent.find(cur);
    if (it == parent.end()) break;
    path.push_front(Pos((int)(cur % m.nx) + 1, (int)(cur / m.nx) + 1));
    cur = it->second;
  }
  if (cur != lf) {
    path.clear();
    return false;
  }
  if (cost) *cost = g.count(lt) ? g[lt] : 0;
  return true;
}

/* plan_route!(agent, dest, pathfinder): stores (overwrites) the path; `nothing` deletes nothing in 5.14 — the
 * old entry is simply replaced only when a path exists */
static void plan_route(Model& m, const Sheep& a, Pos dest) {
  std::deque<Pos> p;
  i64 c = 0;
  if (find_path(m, Pos(a.x, a.y), dest, p, &c)) {
    m.paths[a.id] = p;
    m.path_cost[a.id] = c;
  } else {
    /* plan_route!: `isnothing(path) && return`; the previous route (if any) stays */
  }
}

/* random_walkable(model, pathfinder): rejection-sample random_position(model) = (rand(1:nx), rand(1:ny)) */
static Pos random_walkable(Model& m) {
  while (true) {
    int x, y;
    if (m.variants & VAR_RANDOM_POSITION_Y_FIRST) { y = (int)m.rng.rand_1_to((u64)m.ny); x = (int)m.rng.rand_1_to((u64)m.nx); }
    else { x = (int)m.rng.rand_1_to((u64)m.nx); y = (int)m.rng.rand_1_to((u64)m.ny); }
    if (m.walkmap[m.lin(x, y)]) return Pos(x, y);
  }
}

/* move_along_route!(agent, model::ABM{<:GridSpace}, pathfinder) */
static void move_along_route(Model& m, Sheep& a) {
  std::map<i64, std::deque<Pos>>::iterator it = m.paths.find(a.id);
  if (it == m.paths.end() || it->second.empty()) return;
  const Pos nxt = it->second.front();
  it->second.pop_front();
  move_agent(m, a, nxt.first, nxt.second);
}

/* ---- the reference's own rules ----------------------------------------------------------------------------- */

/* random_walk_to_attractor, src/agent_actions.jl:263-267.  round.(Int, normalize(d)) restated as the exact
 * integer predicate of SURVEY.md A-6, evaluated on 2d (attractor may be a half-integer). */
static void random_walk_to_attractor(Model& m, Sheep& a) {
  const i64 dx2 = (i64)std::llround(2.0 * m.cfg.attractor_x) - 2 * (i64)a.x;
  const i64 dy2 = (i64)std::llround(2.0 * m.cfg.attractor_y) - 2 * (i64)a.y;
  int dirx = 0, diry = 0;
  if (dx2 != 0 || dy2 != 0) {
    if (3 * dx2 * dx2 > dy2 * dy2) dirx = dx2 > 0 ? 1 : -1;
    if (3 * dy2 * dy2 > dx2 * dx2) diry = dy2 > 0 ? 1 : -1;
  }
  walk(m, a, dirx, diry, true); /* walk!(agent, (direction...,), model, ifempty=true) */
}

static double weight_of(const Model& m, i64 k) {
  if ((size_t)k < m.lut.size()) return m.lut[(size_t)k];
  return std::exp(-std::sqrt((double)k) / 3.0); /* f(x) = exp(-x / 3), :319 */
}

/* random_walk_gregarious, src/agent_actions.jl:297-336 */
static int random_walk_gregarious(Model& m, Sheep& a) {
  const std::vector<Pos> possible = nearby_positions(m, a.x, a.y); /* :299 */
  /* nearby_ids(agent, model, r): cells pos .+ β, β over offsets_within_radius (x fastest), in-cell order,
   * own id dropped; :302 */
  const int r = m.cfg.visual_distance;
  bool have = false;
  i64 best = 0;
  Pos closest(0, 0);
  const int side = 2 * r + 1;
  for (int q = 0; q < side * side; ++q) {
    {
      /* offsets_within_radius order: first coordinate fastest (or, VAR_OFFSETS_Y_FASTEST, the second) */
      const int bx = (m.variants & VAR_OFFSETS_Y_FASTEST) ? q / side - r : q % side - r, by = (m.variants & VAR_OFFSETS_Y_FASTEST) ? q % side - r : q / side - r;
      int px = a.x + bx, py = a.y + by;
      if (m.periodic) {
        px = mod1(px, m.nx);
        py = mod1(py, m.ny);
      } else if (px < 1 || px > m.nx || py < 1 || py > m.ny) {
        continue;
      }
      const std::vector<i64>& ids = m.stored[m.lin(px, py)];
      for (size_t i = 0; i < ids.size(); ++i) {
        if (ids[i] == a.id) continue;
        const Sheep& nb = m.agents.at(ids[i]);
        const i64 d2 = dist2(m, nb.x, nb.y, a.x, a.y); /* distance < min_distance on sqrt == on d2 */
        if (!have || d2 < best) {
          have = true;
          best = d2;
          closest = Pos(nb.x, nb.y);
        }
      }
    }
  }
  if (have) {
    const size_t n = possible.size();
    std::vector<double> w(n), p(n);
    for (size_t i = 0; i < n; ++i) w[i] = weight_of(m, dist2(m, possible[i].first, possible[i].second, closest.first, closest.second));
    double s1 = w[0]; /* sum(f.(d)): left to right for n < 16 */
    for (size_t i = 1; i < n; ++i) s1 += w[i];
    for (size_t i = 0; i < n; ++i) p[i] = w[i] / s1;
    double s2 = p[0]; /* Weights(p).sum */
    for (size_t i = 1; i < n; ++i) s2 += p[i];
    /* StatsBase.sample(rng, wv): t = rand(rng) * sum(wv); linear scan */
    const double t = m.rng.rand_f64() * s2;
    size_t i = 0;
    double cw = p[0];
    while (cw < t && i + 1 < n) {
      ++i;
      cw += p[i];
    }
    move_agent(m, a, possible[i].first, possible[i].second); /* :331, unconditional */
  } else {
    const i64 k = m.rng.rand_1_to(8); /* :334 */
    if ((size_t)k > possible.size()) return ABM_E_BADARG; /* BoundsError in the reference (non-periodic edge) */
    move_agent(m, a, possible[k - 1].first, possible[k - 1].second);
  }
  return 0;
}

/* random_walkmap, src/agent_actions.jl:182-186 */
static int random_walkmap(Model& m, Sheep& a) {
  std::vector<Pos> nb = nearby_positions(m, a.x, a.y);
  std::vector<Pos> ok;
  for (size_t i = 0; i < nb.size(); ++i)
    if (m.walkmap[m.lin(nb[i].first, nb[i].second)]) ok.push_back(nb[i]);
  if (ok.empty()) return ABM_E_NOWALKABLE; /* rand(rng, 1:0) throws */
  const i64 k = m.rng.rand_1_to((u64)ok.size());
  move_agent(m, a, ok[k - 1].first, ok[k - 1].second);
  return 0;
}

/* alternated_walk, src/agent_actions.jl:211-242 */
static void alternated_walk(Model& m, Sheep& a) {
  if (m.behav_counter == m.cfg.counter) {
    m.behav = m.rng.rand_1_to(3); /* sample(1:3), routed through the injected RNG */
    if (m.behav == 2) {
      const Pos dest = random_walkable(m);
      plan_route(m, a, dest);
    }
  }
  if (0 < m.behav_counter && m.behav_counter <= m.cfg.counter) {
    if (m.behav == 1) {
      randomwalk(m, a);
    } else if (m.behav == 2) {
      move_along_route(m, a);
    } else if (m.behav == 3) {
      move_agent(m, a, a.x, a.y);
    }
    m.behav_counter -= 1;
  }
  if (m.behav_counter == 0) m.behav_counter = m.cfg.counter;
}

/* execute_walk!, src/agent_actions.jl:109-117 */
static int execute_walk(Model& m, Sheep& a, int walk_type, double prob_random_walk) {
  if (m.rng.rand_f64() < prob_random_walk) { /* rand(model.rng, Uniform(0, 1)) < prob */
    randomwalk(m, a);
    return 0;
  }
  switch (walk_type) {
    case ABM_RANDOM_WALK_ATTRACTOR: random_walk_to_attractor(m, a); return 0;
    case ABM_RANDOM_WALK_GREGARIOUS: return random_walk_gregarious(m, a);
    case ABM_ALTERNATED_WALK: alternated_walk(m, a); return 0;
    case ABM_RANDOM_WALKMAP: return random_walkmap(m, a);
  }
  return ABM_E_BADARG;
}

/* custom_agent_step!, src/agent_actions.jl:29-69.  `a` is a copy of the struct fields the reference mutates in
 * place; like the reference's `agent` binding it stays usable after kill_agent!. */
static int agent_step(Model& m, i64 id) {
  Sheep& a = m.agents.at(id);
  int rc = 0;
  switch (m.cfg.walk_type) { /* :36-52 */
    case ABM_RANDOM_WALK: randomwalk(m, a); break;
    case ABM_RANDOM_WALK_ATTRACTOR:
    case ABM_RANDOM_WALK_GREGARIOUS: rc = execute_walk(m, a, m.cfg.walk_type, m.cfg.prob_random_walk); break;
    case ABM_ALTERNATED_WALK:
    case ABM_RANDOM_WALKMAP: rc = execute_walk(m, a, m.cfg.walk_type, 0.0); break;
    default: return ABM_E_BADARG;
  }
  if (rc) return rc;
  /* reduce_energy!, :85-92 — walk_ocurred is always true */
  Sheep dead_copy;
  Sheep* ag = &a;
  a.energy -= m.cfg.movement_cost;
  if (a.energy < 0) { /* kill_agent!: the struct outlives its removal; the step continues (:59-67) */
    dead_copy = a;
    remove_from_cell(m, a.id, a.x, a.y);
    m.agents.erase(id);
    ag = &dead_copy;
  }
  if (m.cfg.eat) { /* eat!, :139-147 */
    const size_t c = m.lin(ag->x, ag->y);
    if (m.fully_grown[c]) {
      ag->energy += m.cfg.delta_energy;
      m.fully_grown[c] = 0;
    }
  }
  if (m.cfg.reproduce) {
    if (m.rng.rand_f64() <= m.cfg.reproduction_prob) { /* :64, <= */
      ag->energy /= 2; /* reproduce!, :172-178 */
      Sheep kid;
      kid.id = m.maxid + 1; /* nextid(model) */
      kid.x = ag->x;
      kid.y = ag->y;
      kid.energy = ag->energy;
      m.maxid = kid.id;
      m.agents[kid.id] = kid; /* add_agent_pos! */
      m.stored[m.lin(kid.x, kid.y)].push_back(kid.id);
    }
  }
  return 0;
}

/* custom_model_step!, src/model_actions.jl:18-30 */
static void model_step(Model& m) {
  const size_t G = (size_t)m.nx * m.ny;
  const bool wm = m.cfg.walkmap_regrowth != 0;
  for (size_t p = 0; p < G; ++p) {
    if ((!wm || m.walkmap[p] == 1) && !m.fully_grown[p]) {
      if (m.countdown[p] <= 0) {
        m.fully_grown[p] = 1;
        m.countdown[p] = m.cfg.regrowth_time;
      } else {
        m.countdown[p] -= 1;
      }
    }
  }
}

/* Schedulers.randomly under the Philox contract (include/abm.h, ABM_SCHED_RANDOMLY): the order key of `id` in tick `tick`
 * — a 4-round balanced Feistel permutation of [0, 2^(2 hb)), 2^(2 hb) >= nextid, round keys from the Philox block of the
 * reserved agent id 2^64-1 */
static u32 sched_mix(u32 z) { z *= 0x9E3779B1u; z ^= z >> 15; z *= 0x85EBCA77u; z ^= z >> 13; return z; }
static std::vector<i64> schedule(const Model& m) {
  std::vector<i64> ids;
  ids.reserve(m.agents.size());
  for (std::map<i64, Sheep>::const_iterator it = m.agents.begin(); it != m.agents.end(); ++it) ids.push_back(it->first);
  if (m.cfg.scheduler != ABM_SCHED_RANDOMLY) return ids; /* Schedulers.by_id = sorted keys of model.agents */
  int hb = 1;
  while (hb < 16 && (1ull << (2 * hb)) < (u64)(m.maxid + 1)) ++hb;
  const u32 ctr[4] = {0u, (u32)m.tick, 0xFFFFFFFFu, 0xFFFFFFFFu}, key[2] = {(u32)m.cfg.seed, (u32)(m.cfg.seed >> 32)};
  u32 k[4];
  philox4x32_10(ctr, key, k);
  const u32 mask = (1u << hb) - 1u;
  std::vector<std::pair<u32, i64>> keyed;
  keyed.reserve(ids.size());
  for (size_t i = 0; i < ids.size(); ++i) {
    u32 L = (u32)ids[i] >> hb, R = (u32)ids[i] & mask;
    for (int r = 0; r < 4; ++r) { const u32 t = L ^ (sched_mix(R ^ k[r]) & mask); L = R; R = t; }
    keyed.push_back(std::make_pair((L << hb) | R, ids[i]));
  }
  std::sort(keyed.begin(), keyed.end());
  for (size_t i = 0; i < ids.size(); ++i) ids[i] = keyed[i].second;
  return ids;
}

/* Agents.step!(model, agent_step!, model_step!, n), agents_first = true: the scheduler's ids are collected before the first
 * agent steps (newborns are not scheduled until the next tick), dead agents are skipped */
static int step(Model& m, i64 n) {
  for (i64 t = 0; t < n; ++t) {
    const std::vector<i64> ids = schedule(m);
    for (size_t i = 0; i < ids.size(); ++i) {
      if (!m.agents.count(ids[i])) continue;
      m.rng.reset(m.tick, (u64)ids[i]);
      const int rc = agent_step(m, ids[i]);
      if (rc) return rc;
    }
    model_step(m);
    ++m.tick;
  }
  return 0;
}

static std::string g_err;

}  // namespace

struct oracle_handle {
  Model m;
};

extern "C" {

void oracle_config_default(abm_config* c) {
  std::memset(c, 0, sizeof(*c));
  c->struct_size = (int32_t)sizeof(abm_config);
  c->nx = 80; c->ny = 80; c->periodic = 0; c->walk_type = ABM_RANDOM_WALK; c->eat = 1; c->reproduce = 1;
  c->walkmap_regrowth = 0; c->regrowth_time = 30; c->visual_distance = 5; c->counter = 50;
  c->randomwalk_ifempty = 1; c->astar_metric = ABM_ASTAR_DIRECT; c->astar_penalty = 0; c->device = 0;
  c->prob_random_walk = 0.9; c->delta_energy = 4; c->movement_cost = 1; c->reproduction_prob = 0.001;
  c->attractor_x = 40; c->attractor_y = 40; c->seed = 321;
}

int oracle_config_size(void) { return (int)sizeof(abm_config); }

int oracle_create(const abm_config* cfg, oracle_handle** out) {
  if (!cfg || !out || cfg->struct_size != (int32_t)sizeof(abm_config) || cfg->nx < 1 || cfg->ny < 1) return ABM_E_BADARG;
  oracle_handle* h = new oracle_handle();
  Model& m = h->m;
  m.cfg = *cfg;
  m.nx = cfg->nx; m.ny = cfg->ny; m.periodic = cfg->periodic != 0;
  const size_t G = (size_t)m.nx * m.ny;
  m.stored.assign(G, std::vector<i64>());
  m.fully_grown.assign(G, 0);
  m.countdown.assign(G, 0);
  m.walkmap.assign(G, 1);
  m.elevation.assign(G, 0);
  m.rng.seed = cfg->seed;
  *out = h;
  return 0;
}
void oracle_destroy(oracle_handle* h) { delete h; }

int oracle_set_grid(oracle_handle* h, const uint8_t* fg, const int64_t* cd, const uint8_t* wm, const int64_t* el) {
  Model& m = h->m;
  const size_t G = (size_t)m.nx * m.ny;
  if ((fg == nullptr) != (cd == nullptr)) return ABM_E_BADARG;
  for (size_t i = 0; i < G; ++i) {
    m.fully_grown[i] = (fg && fg[i]) ? 1 : 0;
    m.countdown[i] = cd ? cd[i] : 0;
    m.walkmap[i] = wm ? (wm[i] ? 1 : 0) : 1;
    m.elevation[i] = el ? el[i] : 0;
  }
  m.grid_set = true;
  return 0;
}

int oracle_set_agents(oracle_handle* h, int64_t n, const int64_t* id, const int64_t* x, const int64_t* y,
                      const double* e, int64_t next_id) {
  Model& m = h->m;
  m.agents.clear();
  m.paths.clear();
  m.path_cost.clear();
  for (size_t i = 0; i < m.stored.size(); ++i) m.stored[i].clear();
  for (i64 i = 0; i < n; ++i) {
    if (x[i] < 1 || x[i] > m.nx || y[i] < 1 || y[i] > m.ny || id[i] < 1 || id[i] >= next_id || m.agents.count(id[i])) return ABM_E_BADARG;
    Sheep s;
    s.id = id[i]; s.x = (int)x[i]; s.y = (int)y[i]; s.energy = e[i];
    m.agents[s.id] = s;
  }
  /* cell vectors in id order: the order add_agent! would produce for ascending ids */
  for (std::map<i64, Sheep>::iterator it = m.agents.begin(); it != m.agents.end(); ++it)
    m.stored[m.lin(it->second.x, it->second.y)].push_back(it->first);
  m.maxid = next_id - 1;
  return 0;
}

/* the random part of initialize_model (scripts/v0.1.jl:66-79, scripts/v0.5.jl:100-122) under the per-sheep / per-cell
 * streams that include/abm.h states for abm_init_random */
int oracle_init_random(oracle_handle* h, int64_t n_sheep, uint64_t init_seed) {
  Model& m = h->m;
  const i64 span = (i64)(m.cfg.delta_energy * 5);
  if (n_sheep < 0 || n_sheep >= 0xFFFFFFF0ll || span < 1 || m.cfg.n_ranks > 1 || m.cfg.strip_ny > 0) return ABM_E_BADARG;
  AgentRng r;
  r.seed = init_seed;
  bool any = false;
  for (int y = 1; y <= m.ny; ++y)
    for (int x = 1; x <= m.nx; ++x) { /* for p in positions(model): grass on walkable cells only */
      const size_t c = m.lin(x, y);
      m.fully_grown[c] = 0;
      m.countdown[c] = 0;
      if (!m.walkmap[c]) continue;
      any = true;
      r.reset(0xFFFFFFFEull, (u64)c);
      const bool grown = (r.next_u64() & 1ull) != 0;
      m.fully_grown[c] = grown ? 1 : 0;
      m.countdown[c] = grown ? m.cfg.regrowth_time : (m.cfg.regrowth_time > 0 ? r.rand_1_to((u64)m.cfg.regrowth_time) - 1 : 0);
    }
  m.grid_set = true;
  if (n_sheep > 0 && !any) return ABM_E_NOWALKABLE;
  m.agents.clear();
  m.paths.clear();
  m.path_cost.clear();
  for (size_t i = 0; i < m.stored.size(); ++i) m.stored[i].clear();
  for (i64 i = 1; i <= n_sheep; ++i) {
    r.reset(0xFFFFFFFFull, (u64)i);
    Sheep s;
    s.id = i;
    s.energy = (double)(r.rand_1_to((u64)span) - 1);
    int tries = 0;
    while (true) { /* random_position / random_walkable */
      s.x = (int)r.rand_1_to((u64)m.nx);
      s.y = (int)r.rand_1_to((u64)m.ny);
      if (m.walkmap[m.lin(s.x, s.y)]) break;
      if (++tries >= 4096) return ABM_E_NOWALKABLE;
    }
    m.agents[i] = s;
    m.stored[m.lin(s.x, s.y)].push_back(i);
  }
  m.maxid = n_sheep;
  return 0;
}

int oracle_set_global_state(oracle_handle* h, int64_t tick, int64_t behav, int64_t bc) {
  h->m.tick = (u64)tick; h->m.behav = behav; h->m.behav_counter = bc;
  return 0;
}
int oracle_get_global_state(oracle_handle* h, int64_t* tick, int64_t* behav, int64_t* bc, int64_t* next_id) {
  if (tick) *tick = (i64)h->m.tick;
  if (behav) *behav = h->m.behav;
  if (bc) *bc = h->m.behav_counter;
  if (next_id) *next_id = h->m.maxid + 1;
  return 0;
}
int oracle_set_weight_lut(oracle_handle* h, const double* w, int32_t n) {
  h->m.lut.assign(w, w + n);
  return 0;
}
int oracle_step(oracle_handle* h, int64_t n) { return step(h->m, n); }

/* test infrastructure only (no abm_ counterpart): selects alternative readings of Appendix-A items, see Model::variants */
int oracle_set_variants(oracle_handle* h, uint32_t mask) {
  if (!h || (mask & ~(uint32_t)VAR_ALL)) return ABM_E_BADARG;
  h->m.variants = mask;
  return ABM_OK;
}
int oracle_count_agents(oracle_handle* h, int64_t* n) { *n = (i64)h->m.agents.size(); return 0; }

int oracle_get_agents(oracle_handle* h, int64_t* n_inout, int64_t* id, int64_t* x, int64_t* y, double* e) {
  Model& m = h->m;
  const i64 n = (i64)m.agents.size();
  if (*n_inout < n) { *n_inout = n; return ABM_E_SMALL; }
  i64 i = 0;
  for (std::map<i64, Sheep>::iterator it = m.agents.begin(); it != m.agents.end(); ++it, ++i) {
    if (id) id[i] = it->first;
    if (x) x[i] = it->second.x;
    if (y) y[i] = it->second.y;
    if (e) e[i] = it->second.energy;
  }
  *n_inout = n;
  return 0;
}
int oracle_get_grid(oracle_handle* h, uint8_t* fg, int64_t* cd) {
  Model& m = h->m;
  const size_t G = (size_t)m.nx * m.ny;
  for (size_t i = 0; i < G; ++i) {
    if (fg) fg[i] = m.fully_grown[i];
    if (cd) cd[i] = m.countdown[i];
  }
  return 0;
}
/* the narrow transfer forms of include/abm.h, as conversions around the wide ones */
int oracle_set_grass_packed(oracle_handle* h, const uint8_t* state) {
  Model& m = h->m;
  if (!state) return ABM_E_BADARG;
  const size_t G = (size_t)m.nx * m.ny;
  for (size_t i = 0; i < G; ++i) if ((state[i] & 0x7F) == 0x7F) return ABM_E_BADARG;
  for (size_t i = 0; i < G; ++i) { m.fully_grown[i] = state[i] >> 7; m.countdown[i] = state[i] & 0x7F; }
  m.grid_set = true;
  return 0;
}
int oracle_get_grass_packed(oracle_handle* h, uint8_t* state) {
  Model& m = h->m;
  const size_t G = (size_t)m.nx * m.ny;
  for (size_t i = 0; i < G; ++i) state[i] = (uint8_t)((m.fully_grown[i] ? 0x80 : 0) | (m.countdown[i] & 0x7F));
  return 0;
}
int oracle_set_agents_packed(oracle_handle* h, int64_t n, const uint32_t* id, const uint32_t* xy, const double* e, int64_t next_id) {
  if (h->m.nx > 65535 || h->m.ny > 65535) return ABM_E_BADARG;
  std::vector<int64_t> a((size_t)n), x((size_t)n), y((size_t)n);
  for (i64 i = 0; i < n; ++i) { a[i] = id[i]; x[i] = xy[i] & 0xFFFFu; y[i] = xy[i] >> 16; }
  return oracle_set_agents(h, n, a.data(), x.data(), y.data(), e, next_id);
}
int oracle_get_agents_packed(oracle_handle* h, int64_t* n_inout, uint32_t* id, uint32_t* xy, double* e) {
  Model& m = h->m;
  if (m.nx > 65535 || m.ny > 65535) return ABM_E_BADARG;
  const i64 n = (i64)m.agents.size();
  if (*n_inout < n) { *n_inout = n; return ABM_E_SMALL; }
  i64 i = 0;
  for (std::map<i64, Sheep>::iterator it = m.agents.begin(); it != m.agents.end(); ++it, ++i) {
    if (id) id[i] = (uint32_t)it->first;
    if (xy) xy[i] = (uint32_t)it->second.x | ((uint32_t)it->second.y << 16);
    if (e) e[i] = it->second.energy;
  }
  *n_inout = n;
  return 0;
}
int oracle_reduce(oracle_handle* h, int32_t what, double* out) {
  Model& m = h->m;
  double v = 0;
  bool first = true;
  switch (what) {
    case ABM_REDUCE_COUNT_AGENTS: v = (double)m.agents.size(); break;
    case ABM_REDUCE_COUNT_GROWN: for (size_t i = 0; i < m.fully_grown.size(); ++i) v += m.fully_grown[i]; break;
    case ABM_REDUCE_SUM_ENERGY: for (std::map<i64, Sheep>::iterator it = m.agents.begin(); it != m.agents.end(); ++it) v += it->second.energy; break;
    case ABM_REDUCE_MAX_ENERGY:
    case ABM_REDUCE_MIN_ENERGY:
      for (std::map<i64, Sheep>::iterator it = m.agents.begin(); it != m.agents.end(); ++it) {
        const double e = it->second.energy;
        if (first || (what == ABM_REDUCE_MAX_ENERGY ? e > v : e < v)) v = e;
        first = false;
      }
      break;
    default: return ABM_E_BADARG;
  }
  *out = v;
  return 0;
}
int oracle_reduce_where(oracle_handle* h, int32_t what, const abm_filter* f, double* out) {
  Model& m = h->m;
  if (!f || !out || what == ABM_REDUCE_COUNT_GROWN || what < 0 || what > ABM_REDUCE_MIN_ENERGY) return ABM_E_BADARG;
  double v = 0;
  bool first = true;
  for (std::map<i64, Sheep>::iterator it = m.agents.begin(); it != m.agents.end(); ++it) {
    const Sheep& s = it->second;
    if (!(s.energy >= f->energy_lo && s.energy <= f->energy_hi && s.x >= f->x_lo && s.x <= f->x_hi && s.y >= f->y_lo && s.y <= f->y_hi)) continue;
    if (what == ABM_REDUCE_COUNT_AGENTS) v += 1;
    else if (what == ABM_REDUCE_SUM_ENERGY) v += s.energy;
    else if (first || (what == ABM_REDUCE_MAX_ENERGY ? s.energy > v : s.energy < v)) v = s.energy;
    first = false;
  }
  *out = v;
  return 0;
}
int oracle_get_heat(oracle_handle* h, int32_t which, float* out) {
  Model& m = h->m;
  if (!out || which < 0 || which > 1) return ABM_E_BADARG;
  const size_t G = (size_t)m.nx * m.ny;
  const float inv = m.cfg.regrowth_time > 0 ? 1.0f / (float)m.cfg.regrowth_time : 0.0f;
  for (size_t i = 0; i < G; ++i) out[i] = which == 0 ? (float)m.countdown[i] * inv : (float)m.elevation[i];
  return 0;
}
int oracle_plan_routes(oracle_handle* h, int64_t b, const int64_t* aid, const int64_t* gx, const int64_t* gy) {
  Model& m = h->m;
  m.astar_expansions = 0;
  for (i64 i = 0; i < b; ++i) {
    if (!m.agents.count(aid[i]) || gx[i] < 1 || gx[i] > m.nx || gy[i] < 1 || gy[i] > m.ny) return ABM_E_BADARG;
    plan_route(m, m.agents.at(aid[i]), Pos((int)gx[i], (int)gy[i]));
  }
  return 0;
}
int oracle_get_paths(oracle_handle* h, int64_t b, const int64_t* aid, int64_t* off, int64_t* cells_inout,
                     int64_t* px, int64_t* py, int64_t* cost) {
  Model& m = h->m;
  i64 total = 0;
  for (i64 i = 0; i < b; ++i) {
    std::map<i64, std::deque<Pos>>::iterator it = m.paths.find(aid[i]);
    total += it == m.paths.end() ? 0 : (i64)it->second.size();
  }
  if (*cells_inout < total) { *cells_inout = total; return ABM_E_SMALL; }
  i64 o = 0;
  for (i64 i = 0; i < b; ++i) {
    off[i] = o;
    std::map<i64, std::deque<Pos>>::iterator it = m.paths.find(aid[i]);
    if (cost) cost[i] = it == m.paths.end() ? -1 : m.path_cost[aid[i]];
    if (it == m.paths.end()) continue;
    for (size_t k = 0; k < it->second.size(); ++k, ++o) {
      px[o] = it->second[k].first;
      py[o] = it->second[k].second;
    }
  }
  off[b] = o;
  *cells_inout = total;
  return 0;
}
int oracle_get_timing(oracle_handle* h, double* out, int32_t n) {
  for (int i = 0; i < n; ++i) out[i] = 0;
  if (n > 5) out[5] = (double)h->m.astar_expansions;
  return 0;
}
int oracle_set_global_walkmap_bits(oracle_handle*, const uint32_t*) { return ABM_E_STATE; } /* strips only */
int oracle_set_transport(oracle_handle*, const abm_transport*) { return ABM_E_STATE; } /* the oracle is one process, like the reference */
int oracle_nccl_unique_id(void*) { return ABM_E_STATE; }
int oracle_comm_init_nccl(oracle_handle*, const void*) { return ABM_E_STATE; }
int oracle_p2p_export(oracle_handle*, void*) { return ABM_E_STATE; }
int oracle_p2p_connect(oracle_handle*, const void*, int32_t) { return ABM_E_STATE; }
int oracle_set_profiling(oracle_handle*, int32_t) { return 0; }
int oracle_get_kernel_stat(oracle_handle*, int32_t, char*, int32_t, double*, int64_t*, int64_t*) { return ABM_E_SMALL; }
const char* oracle_last_error(oracle_handle*) { return ""; }

/* Philox known-answer access for tests */
void oracle_philox(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) { philox4x32_10(ctr, key, out); }
uint64_t oracle_draw(uint64_t seed, uint64_t tick, uint64_t id, uint64_t n) {
  AgentRng r; r.seed = seed; r.tick = tick; r.id = id; r.n = n;
  return r.next_u64();
}

}  // extern "C"
