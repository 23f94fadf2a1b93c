/*
 * abm_kernels.h — device-side bodies of the general ("G") stepping path.
 *
 * One tick of the reference (Agents.step!: every agent in id order runs custom_agent_step!,
 * src/agent_actions.jl:29-69, then custom_model_step!, src/model_actions.jl:18-30) becomes
 *
 *   propose   per agent: all state-independent work — RNG draws, walk rule, death and birth flags
 *   resolve   rounds over the agents whose move depends on what lower-id agents did before them
 *             (`ifempty` walks, gregarious view): exact rank-ordered fixed point
 *   commit    eat (lowest id on a grown cell wins), energy, slot of the agent inside its new cell
 *   scan      new per-cell offsets (cell-sorted SoA, counting sort)
 *   scatter   agents (minus deaths, plus newborns) into the new cell-sorted SoA
 *   regrow    grass sweep
 *
 * Sequential semantics.  With by-id scheduling, "agent k sees cell c occupied at its turn" is
 *     exists j > k with old(j) = c                      (j has not moved yet)
 *  or exists j < k with final(j) = c and present(j)     (j already moved; present = survived or left a newborn)
 * where final(j) itself may depend on lower ids.  Each agent owns one decision byte; a decision is written
 * once, and only from definite information, so any interleaving of threads yields the sequential result.
 *
 * Layout.  Cells are numbered tile-major: 32x32 tiles, row-major tiles, row-major inside a tile
 * (tcell = tile<<10 | ly<<5 | lx).  Agents are kept sorted by tcell; cell_start[c] .. cell_start[c+1] is the
 * run of agents standing on c (this replaces GridSpace.stored_ids).  Agents move at most one cell per tick
 * (route followers excepted, handled through the `remote` list), so everything that can reach a cell sits in
 * the 3x3 block of runs around it.
 */
#ifndef ABM_KERNELS_H_
#define ABM_KERNELS_H_

#include "abm_platform.h"
#include "abm_rng.h"
#include "../../include/abm.h"

namespace abm {

/* decision byte */
enum {
  ST_DIRMASK = 0x0F,
  ST_DECIDED = 0x10,
  ST_MOVED = 0x20, /* final cell = old cell + direction; else final = old */
  ST_DEAD = 0x40,  /* energy - movement_cost < 0 (reduce_energy!, src/agent_actions.jl:85-92) */
  ST_BIRTH = 0x80  /* rand(model.rng) <= reproduction_prob (src/agent_actions.jl:64) */
};
enum { DIR_STAY = 8, DIR_GREG = 9, DIR_REMOTE = 10 }; /* REMOTE: move_along_route!, final cell in remote_dest[] */
enum { REMOTE_NONE = 0xFFFFFFFFu };

/* grass byte: bit 7 = fully_grown, bits 0..6 = countdown (0..126); 0x7F = padding cell outside the grid */
enum { GRASS_GROWN = 0x80, GRASS_PAD = 0x7F };

enum { TILE_SHIFT = 5, TILE = 32, TILE_CELLS = 1024 };
enum { GREG_MAX_R = 7, GREG_WORDS = 4 };

struct Geo {
  int32_t nx, ny, tiles_x, tiles_y; /* the grid this handle stores: the whole GridSpace, or (strip mode) the rows of one
                                     * strip plus one 32-row ghost tile row above and below */
  uint32_t gt;      /* number of tile-major cells incl. padding */
  int32_t periodic; /* x axis wraps (and y too, unless a strip) */
  int32_t py;       /* y axis wraps inside this handle's grid (0 in strip mode: the ghost rows carry the wrap) */
  int32_t y_lo, y_hi; /* rows a non-wrapping move may land on (clamping bounds of Agents.walk! on a non-periodic space) */
  int32_t y_off;    /* global y = stored y + y_off */
  int32_t own_lo, own_hi; /* rows owned by this handle */
};

/* offsets_at_radius(model, 1) order (x fastest), SURVEY.md A-5 */
ABM_HD void off8(int d, int& dx, int& dy) {
  /* d: 0..7 -> (-1,-1),(0,-1),(1,-1),(-1,0),(1,0),(-1,1),(0,1),(1,1): two bits per entry, dx + 1 and dy + 1 */
  dx = (int)((0x9224u >> (2 * d)) & 3u) - 1;
  dy = (int)((0xA940u >> (2 * d)) & 3u) - 1;
}
ABM_HD int dir_code(int ex, int ey) {
  const int q = (ey + 1) * 3 + (ex + 1);
  return q < 4 ? q : (q == 4 ? (int)DIR_STAY : q - 1);
}
ABM_HD uint32_t enc(const Geo& g, int x, int y) {
  return ((uint32_t)((y >> TILE_SHIFT) * g.tiles_x + (x >> TILE_SHIFT)) << 10) | (uint32_t)((y & 31) << 5) | (uint32_t)(x & 31);
}
ABM_HD void dec(const Geo& g, uint32_t c, int& x, int& y) {
  const uint32_t t = c >> 10;
  const uint32_t ty = t / (uint32_t)g.tiles_x, tx = t - ty * (uint32_t)g.tiles_x;
  x = (int)(tx << TILE_SHIFT) | (int)(c & 31);
  y = (int)(ty << TILE_SHIFT) | (int)((c >> 5) & 31);
}
/* Julia mod1 on 0-based coordinates for |d| <= n */
ABM_HD int wrap(int a, int n) { return a < 0 ? a + n : (a >= n ? a - n : a); }
ABM_HD int wrap_far(int a, int n) { a %= n; return a < 0 ? a + n : a; }
ABM_HD int g_wrap(int a, int n, int periodic) { return periodic ? wrap_far(a, n) : a; }

/* pos + (dx,dy): wrapped when periodic; false when it leaves a non-periodic grid */
ABM_HD bool shift(const Geo& g, int x, int y, int dx, int dy, int& ox, int& oy) {
  ox = x + dx; oy = y + dy; /* |dx|, |dy| <= 1 at every call site */
  bool ok = true;
  if (g.periodic) ox = wrap(ox, g.nx); else ok = ox >= 0 && ox < g.nx;
  if (g.py) oy = wrap(oy, g.ny); else ok = ok && oy >= g.y_lo && oy < g.y_hi;
  return ok;
}

/* the distinct in-bounds cells of the 3x3 block around (x,y), centre first */
ABM_HD int block9(const Geo& g, int x, int y, int px[9], int py[9]) {
  int n = 0;
  px[n] = x; py[n] = y; ++n;
  for (int d = 0; d < 8; ++d) {
    int dx, dy, ox, oy;
    off8(d, dx, dy);
    if (!shift(g, x, y, dx, dy, ox, oy)) continue;
    bool dup = false;
    if (g.nx < 3 || g.ny < 3) { /* wrap can alias cells only on grids narrower than 3 */
      for (int i = 0; i < n; ++i) dup = dup || (px[i] == ox && py[i] == oy);
    }
    if (!dup) { px[n] = ox; py[n] = oy; ++n; }
  }
  return n;
}

struct Params {
  int32_t walk_type, eat, reproduce, ifempty, walkmap_regrowth, regrowth_time, r, counter;
  double prob_random_walk, delta_energy, movement_cost, reproduction_prob;
  int64_t ax2, ay2; /* 2 * attractor, in 1-based coordinates */
  uint64_t seed, tick;
  SchedKey sk;      /* the scheduler's order keys of this tick (abm_rng.h) */
};

/* everything a body needs; plain pointers, passed by value */
struct View {
  Geo g;
  Params p;
  uint32_t n;                /* agents at tick start */
  uint32_t* cell;            /* tcell of each agent, ascending */
  uint32_t* id;
  const uint32_t* okey;      /* order key of each agent: what decides who acts first (== id under Schedulers.by_id) */
  double* energy;
  uint8_t* state;
  uint32_t* slot;            /* position inside the new cell run */
  const uint32_t* cell_start; /* gt + 1 */
  uint8_t* grass;            /* gt */
  const uint32_t* walkbits;  /* gt / 32 words, bit = walkable */
  const double* lut;         /* exp(-sqrt(k)/3) */
  const uint8_t* greg_order; /* box cells by (d^2, β index) */
  const int8_t* greg_dxy;    /* the same order as (dx, dy) pairs */
  const double* greg_table;  /* per closest-neighbour offset: the 8 cumulative wsample weights (GregTableBody) */
  uint32_t* err;             /* sticky device error flags */
  /* ALTERNATED_WALK only (else null): agents that follow a route land on an arbitrary cell; they are chained per
   * destination cell so that `ifempty` probes and the commit see them */
  uint32_t* remote_head;     /* gt, REMOTE_NONE = empty */
  uint32_t* remote_next;     /* per agent */
  uint32_t* remote_dest;     /* per agent: final tcell */
};

enum { ERR_NOWALKABLE = 1, ERR_INTERNAL = 2, ERR_STRIP_CAP = 4, ERR_P2P_TIMEOUT = 8, ERR_LIST_CAP = 16 };

ABM_HD bool walkable(const View& v, uint32_t c) { return (v.walkbits[c >> 5] >> (c & 31)) & 1u; }
ABM_HD bool present(uint8_t s) { return !(s & ST_DEAD) || (s & ST_BIRTH); }

/* ------------------------------------------------------------------------------------------------ propose */

/* effective direction code after GridSpace clamping / wrapping (Agents.walk!) */
ABM_HD int effective_dir(const Geo& g, int x, int y, int dx, int dy) {
  int tx = x + dx, ty = y + dy;
  if (g.periodic) {
    tx = wrap(tx, g.nx);
    if (g.nx < 3 && tx != x) dx = 1; /* both x-neighbours alias on a 2-wide ring: canonical direction */
    if (tx == x) dx = 0;
  } else {
    tx = tx < 0 ? 0 : (tx >= g.nx ? g.nx - 1 : tx);
    dx = tx - x;
  }
  if (g.py) {
    ty = wrap(ty, g.ny);
    if (g.ny < 3 && ty != y) dy = 1;
    if (ty == y) dy = 0;
  } else {
    ty = ty < g.y_lo ? g.y_lo : (ty >= g.y_hi ? g.y_hi - 1 : ty);
    dy = ty - y;
  }
  return dir_code(dx, dy); /* (0, 0) -> DIR_STAY */
}

/* random_walk_to_attractor (src/agent_actions.jl:263-267): round.(Int, normalize(attractor .- pos)) as the
 * exact integer predicate of SURVEY.md A-6 */
ABM_HD void attractor_dir(const Params& p, const Geo& g, int x, int y, int& dx, int& dy) {
  const int64_t ex = p.ax2 - 2 * (int64_t)(x + 1), ey = p.ay2 - 2 * (int64_t)(y + g.y_off + 1);
  dx = 0; dy = 0;
  if (ex == 0 && ey == 0) return;
  if (3 * ex * ex > ey * ey) dx = ex > 0 ? 1 : -1;
  if (3 * ey * ey > ex * ex) dy = ey > 0 ? 1 : -1;
}

/* custom_agent_step! up to the point where other agents matter (src/agent_actions.jl:36-67): the draws of agent `id` at
 * (x, y) for tick v.p.tick, the walk rule, the death and birth flags -> decision byte.  `nowalk`: random_walkmap found no
 * walkable neighbour (the reference throws on rand(1:0)). */
ABM_HD uint8_t propose_state(const View& v, int x, int y, uint32_t id, double energy, bool& nowalk) {
  const Params& p = v.p;
  AgentStream rs;
  rs.init(p.seed, p.tick, id);
  uint8_t st = 0;
  bool cond = false; /* move gated by isempty(target) */
  int dx = 0, dy = 0;
  bool randomwalk = false;
  nowalk = false;

  switch (p.walk_type) {
    case ABM_RANDOM_WALK: randomwalk = true; break; /* :37-39 */
    case ABM_RANDOM_WALK_ATTRACTOR:                 /* :40-42, execute_walk! :109-117 */
      if (rs.next_f64() < p.prob_random_walk) randomwalk = true;
      else { attractor_dir(p, v.g, x, y, dx, dy); cond = true; /* walk!(...; ifempty=true), :266 */ }
      break;
    case ABM_RANDOM_WALK_GREGARIOUS:                /* :43-45 */
      if (rs.next_f64() < p.prob_random_walk) randomwalk = true;
      else { st = DIR_GREG; rs.n += 1; /* draw 1 is consumed in resolve (wsample or rand(1:8)) */ }
      break;
    case ABM_RANDOM_WALKMAP: {                      /* :49-51 with prob 0, random_walkmap :182-186 */
      rs.n += 1; /* rand(model.rng, Uniform(0,1)) < 0 is never true but is drawn */
      /* nearby_walkable(pos, model, pathfinder, 1): bit d of `ok` = neighbour d (offsets_at_radius order) is walkable */
      uint32_t ok = 0;
      const int lx = x & 31;
      const bool inner = x >= 1 && x + 1 < v.g.nx && y >= 1 && y + 1 < v.g.ny && (v.g.py || (y - 1 >= v.g.y_lo && y + 1 < v.g.y_hi));
      if (inner) { /* away from the rim nothing wraps: per row one word of the tile-major bitmap, and the neighbouring tile's
                    * word for the one cell that may lie across a tile border (predicated loads, no divergent path) */
        const uint32_t tx = (uint32_t)(x >> 5);
        uint32_t b3[3];
#pragma unroll
        for (int r = 0; r < 3; ++r) {
          const int yy = y + r - 1;
          const uint32_t* wp = v.walkbits + ((size_t)(yy >> 5) * v.g.tiles_x + tx) * 32 + (yy & 31);
          const uint32_t wc = wp[0];
          const uint32_t wl = lx == 0 ? wp[-32] : wc, wr = lx == 31 ? wp[32] : wc;
          b3[r] = ((wl >> ((lx - 1) & 31)) & 1u) | (((wc >> lx) & 1u) << 1) | (((wr >> ((lx + 1) & 31)) & 1u) << 2);
        }
        ok = b3[0] | ((b3[1] & 1u) << 3) | ((b3[1] & 4u) << 2) | (b3[2] << 5);
      } else {
        int xs[3] = {x - 1, x, x + 1};
        if (v.g.periodic) { xs[0] = wrap(xs[0], v.g.nx); xs[2] = wrap(xs[2], v.g.nx); }
        else { if (xs[0] < 0) xs[0] = -1; if (xs[2] >= v.g.nx) xs[2] = -1; } /* nearby_positions drops out-of-bounds */
#pragma unroll
        for (int r = 0; r < 3; ++r) {
          int yy = y + r - 1;
          if (v.g.py) yy = wrap(yy, v.g.ny); else if (yy < v.g.y_lo || yy >= v.g.y_hi) continue;
          const uint32_t* row = v.walkbits + ((size_t)(yy >> 5) * v.g.tiles_x) * 32 + (yy & 31); /* word of tile column 0 */
#pragma unroll
          for (int q = 0; q < 3; ++q) {
            if (r == 1 && q == 1) continue;
            const int xx = xs[q];
            if (xx < 0) continue;
            const uint32_t bit = (row[(size_t)(xx >> 5) * 32] >> (xx & 31)) & 1u;
            const int idx = r * 3 + q;
            ok |= bit << (idx < 4 ? idx : idx - 1);
          }
        }
      }
      const int nc = popc(ok);
      if (nc == 0) { nowalk = true; st = DIR_STAY | ST_DECIDED; break; }
      const int d = nth_set_bit(ok, (int)rs.next_below((uint64_t)nc)); /* rand(model.rng, 1:length(nearby_pos)) */
      int e = d; /* away from the rim the step neither wraps nor clamps: the direction is the draw */
      if (!inner) { off8(d, dx, dy); e = effective_dir(v.g, x, y, dx, dy); }
      st = (uint8_t)(e | ST_DECIDED | (e != DIR_STAY ? ST_MOVED : 0)); /* move_agent!, unconditional */
      break;
    }
    default: break; /* ABM_ALTERNATED_WALK is proposed by AlternatedProposeBody */
  }
  if (randomwalk) { /* randomwalk!: rand(rng, offsets_at_radius(model, 1)) then walk!(...; ifempty) */
    off8((int)rs.next_below(8), dx, dy);
    cond = true;
    if (!p.ifempty) {
      const int e = effective_dir(v.g, x, y, dx, dy);
      st = (uint8_t)(e | ST_DECIDED | (e != DIR_STAY ? ST_MOVED : 0));
      cond = false;
    }
  }
  if (cond) {
    const int e = effective_dir(v.g, x, y, dx, dy);
    /* a target equal to the own cell is never empty (the agent stands on it): decided, no move */
    st = (uint8_t)(e == DIR_STAY ? (DIR_STAY | ST_DECIDED) : e);
  }
  if (energy - p.movement_cost < 0) st |= ST_DEAD;
  if (p.reproduce && rs.next_f64() <= p.reproduction_prob) st |= ST_BIRTH;
  return st;
}

struct ProposeBody {
  View v;
  uint32_t* pending;       /* worklist of undecided agents (general path; null on the tile path) */
  uint32_t* pending_count;
  uint32_t* birth_flags;   /* tile path: number of agents that will reproduce (sizes the output SoA); else null */
  uint32_t* parent_list;   /* optional: their order keys, in any order (the offspring ids are ranked while the tick resolves) */
  uint32_t parent_cap;

  ABM_HD void operator()(uint64_t i) const {
    const uint32_t k = (uint32_t)i;
    int x, y;
    dec(v.g, v.cell[k], x, y);
    bool nowalk;
    const uint8_t st = propose_state(v, x, y, v.id[k], v.energy[k], nowalk);
    if (nowalk) atomic_or(v.err, ERR_NOWALKABLE);
    v.state[k] = st;
    if (pending && !(st & ST_DECIDED)) pending[atomic_add(pending_count, 1u)] = k;
    if (birth_flags && (st & ST_BIRTH)) {
      const uint32_t q = atomic_add(birth_flags, 1u);
      if (parent_list && q < parent_cap) parent_list[q] = v.okey[k];
    }
  }
};

/* Schedulers.randomly: the order keys of this tick, one thread per agent (own and ghost) */
struct OrderKeyBody {
  SchedKey sk; const uint32_t* id; uint32_t* okey;
  ABM_HD void operator()(uint64_t i) const { okey[i] = order_key(sk, id[i]); }
};

/* ------------------------------------------------------------------------------------------------ resolve */

enum { RES_PENDING = 0, RES_DONE = 1 };

/* `ifempty` walk of agent k towards T = old + dir.  SURVEY.md Appendix C D1. */
ABM_HD int resolve_cond(const View& v, uint32_t k, uint8_t st) {
  const uint32_t myid = v.okey[k]; /* order key: the ids below are only compared */
  int x, y, dx, dy, tx, ty;
  dec(v.g, v.cell[k], x, y);
  off8(st & ST_DIRMASK, dx, dy);
  shift(v.g, x, y, dx, dy, tx, ty);
  bool blocked = false, unknown = false;
  int px[9], py[9];
  const int nb = block9(v.g, tx, ty, px, py);
  for (int b = 0; b < nb && !blocked; ++b) {
    const uint32_t cb = enc(v.g, px[b], py[b]);
    const uint32_t beg = v.cell_start[cb], end = v.cell_start[cb + 1];
    for (uint32_t j = beg; j < end; ++j) {
      if (j == k) continue;
      const uint32_t idj = v.okey[j];
      if (b == 0) { /* agents that start the tick on T */
        if (idj > myid) { blocked = true; break; } /* still there at k's turn */
        const uint8_t sj = load_state(&v.state[j]);
        if (!(sj & ST_DECIDED)) unknown = true;
        else if (!(sj & ST_MOVED) && present(sj)) { blocked = true; break; }
      } else { /* agents one step away from T */
        if (idj > myid) continue; /* has not moved yet */
        const uint8_t sj = load_state(&v.state[j]);
        const int dj = sj & ST_DIRMASK;
        if (dj == DIR_STAY || dj == DIR_REMOTE) continue; /* route followers are found through remote_head */
        bool hits = true; /* an undecided gregarious agent may step onto any neighbour */
        if (dj != DIR_GREG) {
          int jx, jy, fx, fy;
          off8(dj, jx, jy);
          shift(v.g, px[b], py[b], jx, jy, fx, fy);
          hits = fx == tx && fy == ty;
        }
        if (!hits) continue;
        if (!(sj & ST_DECIDED)) unknown = true;
        else if ((sj & ST_MOVED) && present(sj)) { blocked = true; break; }
      }
    }
  }
  if (!blocked && v.remote_head) { /* lower-id route followers that already arrived on T (decided at propose) */
    for (uint32_t j = v.remote_head[enc(v.g, tx, ty)]; j != REMOTE_NONE; j = v.remote_next[j]) {
      if (j != k && v.okey[j] < myid && present(v.state[j])) { blocked = true; break; }
    }
  }
  if (blocked) { v.state[k] = (uint8_t)(st | ST_DECIDED); return RES_DONE; }
  if (unknown) return RES_PENDING;
  v.state[k] = (uint8_t)(st | ST_DECIDED | ST_MOVED);
  return RES_DONE;
}

struct BoxMask {
  uint64_t w[GREG_WORDS];
  ABM_HD void clear() { for (int i = 0; i < GREG_WORDS; ++i) w[i] = 0; }
  ABM_HD void set(int bit) { w[bit >> 6] |= 1ull << (bit & 63); }
  ABM_HD bool get(int bit) const { return (w[bit >> 6] >> (bit & 63)) & 1ull; }
};

/* random_walk_gregarious (src/agent_actions.jl:297-336) for agent k.  SURVEY.md Appendix C D3, A-5, A-7. */
ABM_HD int resolve_greg(const View& v, uint32_t k, uint8_t st) {
  const uint32_t myid = v.okey[k]; /* order key */
  const int r = v.p.r, side = 2 * r + 1;
  int x, y;
  dec(v.g, v.cell[k], x, y);
  BoxMask def, unk;
  def.clear(); unk.clear();
#define ABM_MARK(mask, bx, by) do { const int bx_ = (bx), by_ = (by); \
    if (bx_ >= -r && bx_ <= r && by_ >= -r && by_ <= r) (mask).set((by_ + r) * side + (bx_ + r)); } while (0)
  for (int sy = -r - 1; sy <= r + 1; ++sy) {
    const int cy = wrap_far(y + sy, v.g.ny);
    for (int sx = -r - 1; sx <= r + 1; ++sx) {
      const int cx = wrap_far(x + sx, v.g.nx);
      const uint32_t cb = enc(v.g, cx, cy);
      const uint32_t beg = v.cell_start[cb], end = v.cell_start[cb + 1];
      for (uint32_t j = beg; j < end; ++j) {
        if (j == k) continue;
        if (v.okey[j] > myid) { ABM_MARK(def, sx, sy); continue; } /* still on its old cell */
        const uint8_t sj = load_state(&v.state[j]);
        const int dj = sj & ST_DIRMASK;
        int jx = 0, jy = 0;
        if (dj < 8) off8(dj, jx, jy);
        if (sj & ST_DECIDED) {
          if (!present(sj)) continue;
          if (sj & ST_MOVED) ABM_MARK(def, sx + jx, sy + jy); else ABM_MARK(def, sx, sy);
        } else if (dj == DIR_GREG) {
          for (int ay = -1; ay <= 1; ++ay) for (int ax = -1; ax <= 1; ++ax) ABM_MARK(unk, sx + ax, sy + ay);
        } else {
          ABM_MARK(unk, sx, sy);
          ABM_MARK(unk, sx + jx, sy + jy);
        }
      }
    }
  }
#undef ABM_MARK
  /* closest neighbour: strict `<` over nearby_ids order == minimum of (d^2, β index) over occupied cells */
  int closest = -1;
  const int cells = side * side;
  for (int q = 0; q < cells; ++q) {
    const int bit = v.greg_order[q];
    if (def.get(bit)) { closest = bit; break; }
    if (unk.get(bit)) return RES_PENDING;
  }
  AgentStream rs;
  rs.init(v.p.seed, v.p.tick, v.id[k]);
  rs.n = 1; /* draw 0 was the execute_walk! uniform */
  int dir;
  if (closest < 0) {
    dir = (int)rs.next_below(8); /* :334 */
  } else {
    const int bx = closest % side - r, by = closest / side - r;
    double w[8], s1, s2;
    for (int i = 0; i < 8; ++i) {
      int ox, oy;
      off8(i, ox, oy);
      const int kx = ox - bx, ky = oy - by;
      w[i] = v.lut[kx * kx + ky * ky];
    }
    s1 = w[0];
    for (int i = 1; i < 8; ++i) s1 += w[i];
    for (int i = 0; i < 8; ++i) w[i] = w[i] / s1;
    s2 = w[0];
    for (int i = 1; i < 8; ++i) s2 += w[i];
    const double t = rs.next_f64() * s2; /* StatsBase.sample(rng, wv) */
    int i = 0;
    double cw = w[0];
    while (cw < t && i < 7) { ++i; cw += w[i]; }
    dir = i;
  }
  v.state[k] = (uint8_t)((st & (ST_DEAD | ST_BIRTH)) | dir | ST_DECIDED | ST_MOVED); /* move_agent!, :331 */
  return RES_DONE;
}

/* wsample weights of random_walk_gregarious (src/agent_actions.jl:319-328) for every possible offset of
 * the closest neighbour: w_i = lut[d_i^2], p = w ./ sum(w) (left to right), cumulative sums of p as StatsBase's
 * sample(rng, wv) forms them.  Entry 7 equals sum(p), the factor of the uniform draw.  One thread per offset. */
struct GregTableBody {
  const double* lut;
  int32_t r;
  double* table; /* (2r+1)^2 x 8 */
  ABM_HD void operator()(uint64_t code) const {
    const int side = 2 * r + 1;
    const int bx = (int)(code % (uint64_t)side) - r, by = (int)(code / (uint64_t)side) - r;
    double w[8], s1;
    for (int i = 0; i < 8; ++i) {
      int ox, oy;
      off8(i, ox, oy);
      const int kx = ox - bx, ky = oy - by;
      w[i] = lut[kx * kx + ky * ky];
    }
    s1 = w[0];
    for (int i = 1; i < 8; ++i) s1 += w[i];
    for (int i = 0; i < 8; ++i) w[i] = w[i] / s1;
    double cw = w[0];
    table[code * 8] = cw;
    for (int i = 1; i < 8; ++i) { cw += w[i]; table[code * 8 + i] = cw; }
  }
};

struct ResolveBody {
  View v;
  const uint32_t* in;
  const uint32_t* in_count;
  uint32_t* out;
  uint32_t* out_count;

  ABM_HD void operator()(uint64_t i) const {
    if (i >= *in_count) return;
    const uint32_t k = in[i];
    const uint8_t st = v.state[k];
    const int res = (st & ST_DIRMASK) == DIR_GREG ? resolve_greg(v, k, st) : resolve_cond(v, k, st);
    if (res == RES_PENDING) out[atomic_add(out_count, 1u)] = k;
  }
};

/* ------------------------------------------------------------------------------------------------- commit */

struct BirthRec { uint32_t parent; uint32_t base; };

ABM_HD void final_xy(const View& v, int ox, int oy, uint8_t s, int& fx, int& fy) {
  fx = ox; fy = oy;
  if (s & ST_MOVED) {
    int dx, dy;
    off8(s & ST_DIRMASK, dx, dy);
    shift(v.g, ox, oy, dx, dy, fx, fy);
  }
}

/* reduce_energy! + eat! + reproduce! (src/agent_actions.jl:54-67) once every final cell is known, plus the
 * bookkeeping of the counting sort into the next tick's cell-sorted order. */
struct CommitBody {
  View v;
  uint32_t* new_count;  /* gt + 1, zeroed */
  uint32_t* birth_bits; /* bitmap over ids */
  BirthRec* births;
  uint32_t* birth_count;
  uint32_t* new_cell;   /* final tcell of each agent */

  ABM_HD void operator()(uint64_t i) const {
    const uint32_t k = (uint32_t)i;
    const uint8_t st = v.state[k];
    const uint32_t myid = v.okey[k]; /* order key: first come, first served */
    int x, y, fx, fy;
    dec(v.g, v.cell[k], x, y);
    if ((st & ST_DIRMASK) == DIR_REMOTE) dec(v.g, v.remote_dest[k], fx, fy);
    else final_xy(v, x, y, st, fx, fy);
    const uint32_t fc = enc(v.g, fx, fy);
    int px[9], py[9];
    const int nb = block9(v.g, fx, fy, px, py);
    bool lower = false;
    uint32_t slot_alive = 0, slot_birth = 0, alive_tot = 0, birth_tot = 0;
    for (int b = 0; b < nb; ++b) {
      const uint32_t cb = enc(v.g, px[b], py[b]);
      const uint32_t beg = v.cell_start[cb], end = v.cell_start[cb + 1];
      for (uint32_t j = beg; j < end; ++j) {
        const uint8_t sj = v.state[j];
        if (b == 0) { if (sj & ST_MOVED) continue; }
        else {
          if (!(sj & ST_MOVED) || (sj & ST_DIRMASK) == DIR_REMOTE) continue;
          int jx, jy;
          final_xy(v, px[b], py[b], sj, jx, jy);
          if (jx != fx || jy != fy) continue;
        }
        const bool alive = !(sj & ST_DEAD), birth = (sj & ST_BIRTH) != 0;
        alive_tot += alive; birth_tot += birth;
        if (j != k && v.okey[j] < myid) { lower = true; slot_alive += alive; slot_birth += birth; }
      }
    }
    if (v.remote_head) { /* route followers landing on fc */
      for (uint32_t j = v.remote_head[fc]; j != REMOTE_NONE; j = v.remote_next[j]) {
        const uint8_t sj = v.state[j];
        const bool alive = !(sj & ST_DEAD), birth = (sj & ST_BIRTH) != 0;
        alive_tot += alive; birth_tot += birth;
        if (j != k && v.okey[j] < myid) { lower = true; slot_alive += alive; slot_birth += birth; }
      }
    }
    /* eat!: the first agent (lowest id) to stand on a grown cell eats it; killed agents still eat */
    double e = v.energy[k] - v.p.movement_cost;
    if (v.p.eat && !lower && (v.grass[fc] & GRASS_GROWN)) {
      e += v.p.delta_energy;
      v.grass[fc] = (uint8_t)(v.grass[fc] & ~GRASS_GROWN); /* single writer: the lowest id on fc */
    }
    if (st & ST_BIRTH) {
      e /= 2; /* reproduce!, :172-178 */
      atomic_or(&birth_bits[myid >> 5], 1u << (myid & 31));
      BirthRec rec;
      rec.parent = k;
      rec.base = alive_tot + slot_birth;
      births[atomic_add(birth_count, 1u)] = rec;
    }
    v.energy[k] = e;
    v.slot[k] = slot_alive;
    new_cell[k] = fc;
    if (!lower) new_count[fc] = alive_tot + birth_tot;
  }
};

/* survivors into the next tick's arrays */
struct ScatterBody {
  View v;
  const uint32_t* new_cell;
  const uint32_t* new_start;
  uint32_t* out_cell;
  uint32_t* out_id;
  double* out_energy;

  ABM_HD void operator()(uint64_t i) const {
    const uint32_t k = (uint32_t)i;
    if (v.state[k] & ST_DEAD) return; /* kill_agent! */
    const uint32_t fc = new_cell[k];
    const uint32_t pos = new_start[fc] + v.slot[k];
    out_cell[pos] = fc;
    out_id[pos] = v.id[k];
    out_energy[pos] = v.energy[k];
  }
};

/* offspring: id = nextid(model) in order of the parents' ids */
struct BirthScatterBody {
  View v;
  const BirthRec* births;
  const uint32_t* birth_bits;
  const uint32_t* birth_prefix; /* exclusive popcount prefix per bitmap word */
  const uint32_t* new_cell;
  const uint32_t* new_start;
  uint32_t next_id;
  uint32_t* out_cell;
  uint32_t* out_id;
  double* out_energy;

  ABM_HD void operator()(uint64_t i) const {
    const BirthRec rec = births[i];
    const uint32_t k = rec.parent;
    const uint32_t pid = v.okey[k]; /* nextid in the order the parents were stepped */
    const uint32_t w = pid >> 5;
    const uint32_t rank = birth_prefix[w] + (uint32_t)popc(birth_bits[w] & ((1u << (pid & 31)) - 1u));
    const uint32_t fc = new_cell[k];
    const uint32_t pos = new_start[fc] + rec.base;
    out_cell[pos] = fc;
    out_id[pos] = next_id + rank;
    out_energy[pos] = v.energy[k]; /* offspring gets the halved energy */
  }
};

struct PopcountBody { /* per-word popcounts as the input of the rank scan */
  const uint32_t* bits;
  uint32_t* out;
  ABM_HD void operator()(uint64_t i) const { out[i] = (uint32_t)popc(bits[i]); }
};

/* the same per block of 8 words (two 16-byte loads): the scan that follows is 8 times shorter; a lookup adds the popcounts
 * of the words before its own inside the block (rank_in_blocks) */
struct PopcountBlockBody {
  const uint32_t* bits; /* a multiple of 8 words, 32-byte aligned */
  uint32_t* out;
  ABM_HD void operator()(uint64_t b) const {
    uint32_t c = 0;
#if defined(__CUDA_ARCH__)
    const uint4 lo = reinterpret_cast<const uint4*>(bits)[2 * b], hi = reinterpret_cast<const uint4*>(bits)[2 * b + 1];
    c = (uint32_t)(__popc(lo.x) + __popc(lo.y) + __popc(lo.z) + __popc(lo.w) + __popc(hi.x) + __popc(hi.y) + __popc(hi.z) + __popc(hi.w));
#else
    for (int j = 0; j < 8; ++j) c += (uint32_t)popc(bits[8 * b + j]);
#endif
    out[b] = c;
  }
};
/* number of set bits below bit `key`, from the exclusive prefix over 8-word blocks */
ABM_HD uint32_t rank_in_blocks(const uint32_t* bits, const uint32_t* prefix8, uint32_t key) {
  const uint32_t wd = key >> 5, blk = wd >> 3;
  uint32_t r = prefix8[blk] + (uint32_t)popc(bits[wd] & ((1u << (key & 31)) - 1u));
  for (uint32_t w = blk << 3; w < wd; ++w) r += (uint32_t)popc(bits[w]);
  return r;
}

/* ------------------------------------------------------------------------------------------------- regrow */

/* custom_model_step! on four cells at once (one grass byte each: bit 7 = fully_grown, bits 0..6 = countdown, 0x7F =
 * padding).  Per byte: if allowed (walkmap bit, when used), not padding and not grown: countdown == 0 -> grown with
 * countdown = regrowth_time, else countdown -= 1.  SWAR: every step below works on the four bytes independently. */
ABM_HD uint32_t regrow4(uint32_t w, uint32_t rt, uint32_t walk4, bool use_walk) {
  const uint32_t H = 0x80808080u;
  const uint32_t low = w & 0x7F7F7F7Fu;
  const uint32_t zero = ((low + 0x7F7F7F7Fu) & H) ^ H;      /* 0x80 where countdown == 0 */
  const uint32_t pad = (low + 0x01010101u) & H;              /* 0x80 where the byte is 0x7F (or 0xFF) */
  uint32_t act = ~w & H & ~pad;                              /* not grown, not padding */
  if (use_walk) act &= (((walk4 & 0xFu) * 0x00204081u) & 0x01010101u) << 7; /* bit b of walk4 -> byte b */
  const uint32_t grow = act & zero, dec = act & ~zero;
  return (w - (dec >> 7)) | grow | ((grow >> 7) * rt);
}

/* custom_model_step! (src/model_actions.jl:18-30): 16 cells per thread */
struct RegrowBody {
  uint32_t* grass4; /* gt / 4 words */
  const uint32_t* walkbits;
  uint32_t rt;
  int32_t use_walk;

  ABM_HD void operator()(uint64_t i) const {
    uint32_t w[4];
#if defined(__CUDA_ARCH__)
    const uint4 q = reinterpret_cast<const uint4*>(grass4)[i];
    w[0] = q.x; w[1] = q.y; w[2] = q.z; w[3] = q.w;
#else
    for (int a = 0; a < 4; ++a) w[a] = grass4[4 * i + a];
#endif
    uint32_t wb = 0;
    if (use_walk) wb = walkbits[i >> 1] >> ((i & 1) * 16);
    for (int a = 0; a < 4; ++a) w[a] = regrow4(w[a], rt, wb >> (4 * a), use_walk != 0);
#if defined(__CUDA_ARCH__)
    reinterpret_cast<uint4*>(grass4)[i] = make_uint4(w[0], w[1], w[2], w[3]);
#else
    for (int a = 0; a < 4; ++a) grass4[4 * i + a] = w[a];
#endif
  }
};

/* ------------------------------------------------------------------------------------------- alternated walk */

/* Schedulers.by_id rank of every agent (its agent-step index inside the tick) and the inverse map */
struct RankBody {
  const uint32_t* id; const uint32_t* alive_bits; const uint32_t* prefix;
  uint32_t* rank; uint32_t* by_rank;
  ABM_HD void operator()(uint64_t i) const {
    const uint32_t a = id[i], w = a >> 5;
    const uint32_t r = prefix[w] + (uint32_t)popc(alive_bits[w] & ((1u << (a & 31)) - 1u));
    rank[i] = r; by_rank[r] = (uint32_t)i;
  }
};

struct RouteView { /* pathfinder.agent_paths, see abm_astar.h */
  uint32_t* slot_of_id; const uint64_t* r_off; const uint32_t* r_len; uint32_t* r_pos; const uint32_t* pool;
};

/* the draws an agent makes when it finds behav_counter == counter (src/agent_actions.jl:212-222): sample(1:3) and,
 * for 2, the rejection loop of random_walkable.  Leaves the stream positioned after them. */
ABM_HD int alt_switch_draws(const View& v, AgentStream& rs, uint32_t* goal) {
  const int b = (int)rs.next_below(3) + 1;
  if (b == 2) {
    while (true) {
      const int gx = (int)rs.next_below((uint64_t)v.g.nx), gy = (int)rs.next_below((uint64_t)v.g.ny);
      const uint32_t c = enc(v.g, gx, gy);
      if (walkable(v, c)) { if (goal) *goal = c; break; }
    }
  }
  return b;
}

/* one thread per behaviour switch of this tick: the agent whose step finds the global counter at `counter` */
struct AltSwitchBody {
  View v;
  const uint32_t* by_rank;
  uint32_t first, counter;
  uint8_t* seg_behav;     /* behaviour of segment i */
  uint32_t* plan_agent;   /* requests: agent index, goal tcell */
  uint32_t* plan_goal;
  uint32_t* plan_count;
  ABM_HD void operator()(uint64_t i) const {
    const uint32_t k = by_rank[first + (uint32_t)i * counter];
    AgentStream rs;
    rs.init(v.p.seed, v.p.tick, v.id[k]);
    (void)rs.next_f64(); /* execute_walk!'s uniform (prob_random_walk = 0) */
    uint32_t goal = 0;
    const int b = alt_switch_draws(v, rs, &goal);
    seg_behav[i] = (uint8_t)b;
    if (b == 2) {
      const uint32_t q = atomic_add(plan_count, 1u);
      plan_agent[q] = k; plan_goal[q] = goal;
    }
  }
};

/* alternated_walk (src/agent_actions.jl:211-242) with the model-global counter in closed form */
struct AltProposeBody {
  View v;
  RouteView rt;
  const uint32_t* rank;
  const uint8_t* seg_behav;
  uint32_t first, counter;
  int32_t prev_behav;   /* model.behav[1] at tick start */
  int32_t inert_rank0;  /* behav_counter was 0 at tick start: the first agent-step only resets the counter */
  uint32_t* pending; uint32_t* pending_count;

  ABM_HD void operator()(uint64_t i) const {
    const uint32_t k = (uint32_t)i;
    const Params& p = v.p;
    const uint32_t id = v.id[k], r = rank[k];
    int x, y;
    dec(v.g, v.cell[k], x, y);
    AgentStream rs;
    rs.init(p.seed, p.tick, id);
    (void)rs.next_f64();
    int b = prev_behav;
    bool acts = true;
    if (inert_rank0 && r == 0) acts = false;
    else if (r >= first) {
      const uint32_t seg = (r - first) / counter;
      b = seg_behav[seg];
      if ((r - first) % counter == 0) (void)alt_switch_draws(v, rs, nullptr); /* this agent made the switch draws */
    }
    uint8_t st = DIR_STAY | ST_DECIDED; /* behav 3: move_agent!(agent, agent.pos, model); behav 0: nothing */
    if (acts && b == 1) { /* randomwalk!(agent, model): ignores the walkmap (Appendix D-6) */
      int dx, dy;
      off8((int)rs.next_below(8), dx, dy);
      const int e = effective_dir(v.g, x, y, dx, dy);
      if (!p.ifempty) st = (uint8_t)(e | ST_DECIDED | (e != DIR_STAY ? ST_MOVED : 0));
      else st = (uint8_t)(e == DIR_STAY ? (DIR_STAY | ST_DECIDED) : e);
    } else if (acts && b == 2) { /* move_along_route!: unconditional, one cell of the stored path */
      const uint32_t s = rt.slot_of_id[id];
      if (s) {
        const uint32_t pos = rt.r_pos[s - 1];
        if (pos < rt.r_len[s - 1]) {
          const uint32_t dest = rt.pool[rt.r_off[s - 1] + pos];
          rt.r_pos[s - 1] = pos + 1;
          if (dest != v.cell[k]) {
            st = DIR_REMOTE | ST_DECIDED | ST_MOVED;
            v.remote_dest[k] = dest;
            /* lock-free push on the destination cell's chain */
            uint32_t old = atomic_exch(&v.remote_head[dest], k);
            v.remote_next[k] = old;
          }
        }
      }
    }
    if (v.energy[k] - p.movement_cost < 0) st |= ST_DEAD;
    if (p.reproduce && rs.next_f64() <= p.reproduction_prob) st |= ST_BIRTH;
    v.state[k] = st;
    if (!(st & ST_DECIDED)) pending[atomic_add(pending_count, 1u)] = k;
  }
};

struct RemoteResetBody { /* empties the chains touched this tick */
  View v;
  ABM_HD void operator()(uint64_t i) const {
    if ((v.state[i] & ST_DIRMASK) == DIR_REMOTE) v.remote_head[v.remote_dest[i]] = REMOTE_NONE;
  }
};

struct IdIndexBody { /* id -> position in the cell-sorted arrays */
  const uint32_t* id; uint32_t* idx_of_id;
  ABM_HD void operator()(uint64_t i) const { idx_of_id[id[i]] = (uint32_t)i; }
};

/* ---------------------------------------------------------------------------------------- boundary helpers */

/* one agent record of the host: wide (abm_set_agents: id, x, y as int64) or packed (abm_set_agents_packed: id u32,
 * x | y << 16); y becomes the stored row (global y minus the strip offset) */
ABM_HD void host_record(const int64_t* hid, const int64_t* hx, const int64_t* hy, const uint32_t* pid, const uint32_t* pxy,
                        const Geo& g, uint64_t i, int64_t& a, int64_t& x, int64_t& y) {
  if (pid) { const uint32_t q = pxy[i]; a = (int64_t)pid[i]; x = (int64_t)(q & 0xFFFFu); y = (int64_t)(q >> 16) - g.y_off; }
  else { a = hid[i]; x = hx[i]; y = hy[i] - g.y_off; }
}

/* abm_set_agents: host records -> unsorted SoA + counting-sort bookkeeping */
struct IngestBody {
  Geo g;
  const int64_t* hid; const int64_t* hx; const int64_t* hy; const double* he;
  const uint32_t* pid; const uint32_t* pxy; /* abm_set_agents_packed: id, x | y << 16 (else null) */
  int64_t next_id;
  uint32_t* cell; uint32_t* id; double* energy; uint8_t* state; uint32_t* slot;
  uint32_t* new_count;
  uint32_t* alive_bits;
  uint32_t* err;

  ABM_HD void operator()(uint64_t i) const {
    int64_t x, y, a;
    host_record(hid, hx, hy, pid, pxy, g, i, a, x, y);
    if (x < 1 || x > g.nx || y < 1 + g.own_lo || y > g.own_hi || a < 1 || a >= next_id) { atomic_or(err, ERR_INTERNAL); cell[i] = 0; id[i] = 0; energy[i] = 0; state[i] = ST_DEAD; return; }
    const uint32_t c = enc(g, (int)(x - 1), (int)(y - 1));
    const uint32_t bit = 1u << ((uint32_t)a & 31);
    if (atomic_or(&alive_bits[(uint32_t)a >> 5], bit) & bit) atomic_or(err, ERR_INTERNAL); /* duplicate id */
    cell[i] = c; id[i] = (uint32_t)a; energy[i] = he[i]; state[i] = 0;
    slot[i] = atomic_add(&new_count[c], 1u);
  }
};

struct SetBitsBody { /* alive-id bitmap of the current population */
  const uint32_t* id;
  uint32_t* bits;
  ABM_HD void operator()(uint64_t i) const { const uint32_t a = id[i]; atomic_or(&bits[a >> 5], 1u << (a & 31)); }
};

/* abm_get_agents: cell-sorted SoA -> id-ordered host records */
struct ExportBody {
  Geo g;
  const uint32_t* cell; const uint32_t* id; const double* energy;
  const uint32_t* alive_bits; const uint32_t* prefix;
  int64_t* oid; int64_t* ox; int64_t* oy; double* oe;
  uint32_t* pid; uint32_t* pxy; /* abm_get_agents_packed: id, x | y << 16 instead of oid/ox/oy (else null) */

  ABM_HD void operator()(uint64_t i) const {
    const uint32_t a = id[i];
    const uint32_t w = a >> 5;
    const uint32_t rank = prefix[w] + (uint32_t)popc(alive_bits[w] & ((1u << (a & 31)) - 1u));
    int x, y;
    dec(g, cell[i], x, y);
    if (pid) { pid[rank] = a; pxy[rank] = (uint32_t)(x + 1) | ((uint32_t)(y + g.y_off + 1) << 16); }
    else { oid[rank] = a; ox[rank] = x + 1; oy[rank] = y + g.y_off + 1; }
    oe[rank] = energy[i];
  }
};

/* abm_set_grid: column-major host arrays -> tile-major packed bytes; one thread per tile-major cell */
struct GridImportBody {
  Geo g;
  const uint8_t* fg; const int64_t* cd; const uint8_t* wm; const int64_t* el;
  uint64_t row0, rows; /* the chunk of grid rows [row0, row0+rows) present in the staging buffers */
  uint8_t* grass; uint32_t* walkbits; int16_t* elev; uint32_t* err;

  ABM_HD void operator()(uint64_t i) const {
    /* i indexes (row in chunk, x) */
    const uint64_t yy = i / (uint64_t)g.nx, x = i - yy * (uint64_t)g.nx;
    const uint64_t y = row0 + yy;
    const uint32_t c = enc(g, (int)x, (int)y);
    const int64_t k = cd ? cd[i] : 0;
    if (k < 0 || k > 126) atomic_or(err, ERR_INTERNAL);
    grass[c] = (uint8_t)(((fg && fg[i]) ? GRASS_GROWN : 0) | (uint8_t)(k & 0x7F));
    if (!wm || wm[i]) atomic_or(&walkbits[c >> 5], 1u << (c & 31));
    if (el) {
      const int64_t e = el[i];
      if (e < -32768 || e > 32767) atomic_or(err, ERR_INTERNAL);
      elev[c] = (int16_t)e;
    }
  }
};

/* abm_init_random: the grass of initialize_model (scripts/v0.1.jl:73-78; walkable cells only, scripts/v0.5.jl:115-122),
 * one Philox stream per cell keyed by its linear index (x fastest) */
struct InitGrassBody {
  Geo g; uint64_t seed; int32_t regrowth_time, open_terrain;
  uint8_t* grass; uint32_t* walkbits;
  ABM_HD void operator()(uint64_t i) const { /* i = cell of the rows this handle owns, x fastest */
    const uint64_t yy = i / (uint64_t)g.nx, x = i - yy * (uint64_t)g.nx;
    const uint64_t y = yy + (uint64_t)g.own_lo;           /* stored row */
    const uint64_t key = (uint64_t)((int64_t)y + g.y_off) * (uint64_t)g.nx + x; /* the cell's index in the whole GridSpace */
    const uint32_t c = enc(g, (int)x, (int)y);
    if (open_terrain) atomic_or(&walkbits[c >> 5], 1u << (c & 31));
    else if (!((walkbits[c >> 5] >> (c & 31)) & 1u)) { grass[c] = 0; return; }
    AgentStream r;
    r.init(seed, 0xFFFFFFFEull, key);
    const bool grown = (r.next_u64() & 1ull) != 0;
    const uint32_t k = grown ? (uint32_t)regrowth_time : (regrowth_time > 0 ? r.next_below((uint64_t)regrowth_time) : 0u);
    grass[c] = (uint8_t)((grown ? GRASS_GROWN : 0) | (k & 0x7F));
  }
};
/* abm_init_random: sheep i + 1 — energy = rand(1:span) - 1, then random_position until the cell is walkable
 * (scripts/v0.1.jl:66-71, random_walkable in scripts/v0.5.jl:103) — written as host-layout records for the ingest */
struct InitSheepBody {
  Geo g; uint64_t seed, span; const uint32_t* walkbits;
  int32_t ny_global;               /* strips: random_position draws over the whole GridSpace ... */
  const uint32_t* global_bits;     /* ... whose walkmap is this (1 bit per cell, x fastest; null = every cell walkable) */
  int64_t* s_id; int64_t* s_x; int64_t* s_y; double* s_e; uint32_t* err;
  /* strips: only the sheep that land on the own rows are kept, as packed records behind a cursor (a first pass with
   * p_e == null counts them) */
  int32_t strip; uint32_t* cursor; double* p_e; uint32_t* p_id; uint32_t* p_xy;
  ABM_HD void operator()(uint64_t i) const {
    AgentStream r;
    r.init(seed, 0xFFFFFFFFull, i + 1);
    const uint32_t e = r.next_below(span);
    uint32_t x = 0, y = 0;
    int tries = 0;
    while (true) {
      x = r.next_below((uint64_t)g.nx);
      y = r.next_below((uint64_t)(strip ? ny_global : g.ny));
      bool ok;
      if (strip) { const uint64_t l = (uint64_t)y * (uint64_t)g.nx + x; ok = !global_bits || ((global_bits[l >> 5] >> (l & 31)) & 1u); }
      else { const uint32_t c = enc(g, (int)x, (int)y); ok = (walkbits[c >> 5] >> (c & 31)) & 1u; }
      if (ok) break;
      if (++tries >= 4096) { atomic_or(err, ERR_NOWALKABLE); break; }
    }
    if (!strip) { s_id[i] = (int64_t)(i + 1); s_x[i] = (int64_t)x + 1; s_y[i] = (int64_t)y + 1; s_e[i] = (double)e; return; }
    const int64_t ys = (int64_t)y - g.y_off; /* stored row */
    if (ys < g.own_lo || ys >= g.own_hi) return; /* another strip's sheep */
    const uint32_t k = atomic_add(cursor, 1u);
    if (p_e) { p_e[k] = (double)e; p_id[k] = (uint32_t)(i + 1); p_xy[k] = (x + 1) | ((y + 1) << 16); }
  }
};

struct GhostWalkableBody { /* abm_init_random on a strip without terrain: the ghost rows are open land as well */
  Geo g; uint32_t* walkbits;
  ABM_HD void operator()(uint64_t i) const {
    const uint64_t y = i / (uint64_t)g.nx, x = i - y * (uint64_t)g.nx;
    if ((int64_t)y >= g.own_lo && (int64_t)y < g.own_hi) return;
    const uint32_t c = enc(g, (int)x, (int)y);
    atomic_or(&walkbits[c >> 5], 1u << (c & 31));
  }
};

struct GridExportBody {
  Geo g;
  const uint8_t* grass;
  uint64_t row0;
  uint8_t* fg; int64_t* cd;
  ABM_HD void operator()(uint64_t i) const {
    const uint64_t yy = i / (uint64_t)g.nx, x = i - yy * (uint64_t)g.nx;
    const uint8_t s = grass[enc(g, (int)x, (int)(row0 + yy))];
    if (fg) fg[i] = (s & GRASS_GROWN) ? 1 : 0;
    if (cd) cd[i] = s & 0x7F;
  }
};

/* abm_set_grass_packed / abm_get_grass_packed: the grass bytes (bit 7 = fully_grown, bits 0..6 = countdown) between the
 * host's column-major array (x fastest) and the tile-major array; four cells of one row per thread */
struct GrassPackedBody {
  Geo g;
  uint8_t* grass;        /* tile-major */
  uint8_t* host;         /* nx * ny, x fastest (staged on the device) */
  int32_t to_device;
  uint32_t* err;
  uint32_t* open_walkbits; /* import without terrain: every cell of the grid becomes walkable (else null) */
  ABM_HD void operator()(uint64_t i) const {
    const uint64_t per_row = (uint64_t)(g.nx + 3) >> 2;
    const uint64_t y = i / per_row, x0 = (i - y * per_row) << 2;
    for (int k = 0; k < 4 && x0 + k < (uint64_t)g.nx; ++k) {
      const uint32_t c = enc(g, (int)(x0 + k), (int)y);
      const uint64_t l = y * (uint64_t)g.nx + x0 + k;
      if (to_device) {
        const uint8_t b = host[l];
        if ((b & 0x7F) == GRASS_PAD) atomic_or(err, ERR_INTERNAL);
        grass[c] = b;
        if (open_walkbits) atomic_or(&open_walkbits[c >> 5], 1u << (c & 31));
      }
      else host[l] = grass[c];
    }
  }
};

/* reductions for adata/mdata (run!): per-thread partials combined with 64-bit atomics; energy sums are done
 * in fixed point on the host side of the ABI only when exactness matters — here a tree of doubles per block
 * would not be order-stable, so sum uses a deterministic two-pass layout (see engine). */
struct CountGrownBody {
  const uint32_t* grass4;
  unsigned long long* out;
  ABM_HD void operator()(uint64_t i) const {
    const uint32_t w = grass4[i];
    /* padding cells are 0x7F: bit 7 clear */
    const uint32_t c = (uint32_t)popc(w & 0x80808080u);
    if (c) atomic_add64(out, c);
  }
};

/* sum / max / min of agent energy: one partial per 1024-agent chunk, chunks combined in order by the host,
 * so the result does not depend on thread scheduling */
enum { ENERGY_CHUNK = 1024 };
struct EnergyPartialBody {
  const double* energy;
  uint64_t n;
  int32_t what;
  double* out;
  ABM_HD void operator()(uint64_t i) const {
    const uint64_t lo = i * ENERGY_CHUNK, hi = lo + ENERGY_CHUNK < n ? lo + ENERGY_CHUNK : n;
    double v = energy[lo];
    for (uint64_t k = lo + 1; k < hi; ++k) {
      const double e = energy[k];
      if (what == ABM_REDUCE_SUM_ENERGY) v += e;
      else if (what == ABM_REDUCE_MAX_ENERGY) v = e > v ? e : v;
      else v = e < v ? e : v;
    }
    out[i] = v;
  }
};

/* abm_reduce_where: count / sum / min / max of energy over the agents inside a box of (energy, x, y); one partial per chunk,
 * combined in order by the host */
struct WherePartialBody {
  Geo g; const uint32_t* cell; const double* energy; uint64_t n; int32_t what;
  double e_lo, e_hi; int32_t x_lo, x_hi, y_lo, y_hi; /* 0-based stored coordinates, inclusive */
  double* out; uint32_t* any;
  ABM_HD void operator()(uint64_t i) const {
    const uint64_t lo = i * ENERGY_CHUNK, hi = lo + ENERGY_CHUNK < n ? lo + ENERGY_CHUNK : n;
    double v = 0; uint32_t cnt = 0;
    for (uint64_t k = lo; k < hi; ++k) {
      const double e = energy[k];
      int x, y;
      dec(g, cell[k], x, y);
      if (!(e >= e_lo && e <= e_hi && x >= x_lo && x <= x_hi && y >= y_lo && y <= y_hi)) continue;
      if (what == ABM_REDUCE_COUNT_AGENTS) v += 1;
      else if (what == ABM_REDUCE_SUM_ENERGY) v += e;
      else if (what == ABM_REDUCE_MAX_ENERGY) v = (cnt == 0 || e > v) ? e : v;
      else v = (cnt == 0 || e < v) ? e : v;
      ++cnt;
    }
    out[i] = v; any[i] = cnt;
  }
};

/* abm_get_heat: countdown / regrowth_time, or the elevation, as floats in the host's layout (x fastest) */
struct HeatExportBody {
  Geo g; const uint8_t* grass; const int16_t* elev; int32_t which; float inv_rt; float* out;
  ABM_HD void operator()(uint64_t i) const {
    const uint64_t y = i / (uint64_t)g.nx, x = i - y * (uint64_t)g.nx;
    const uint32_t c = enc(g, (int)x, (int)y);
    out[i] = which == 0 ? (float)(grass[c] == GRASS_PAD ? 0 : (grass[c] & 0x7F)) * inv_rt : (float)elev[c];
  }
};

}  // namespace abm
#endif /* ABM_KERNELS_H_ */
