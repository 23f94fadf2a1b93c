/*
 * abm_astar.h — batched A* on the grid (kernel K8) and the route store.
 *
 * Replaces Agents.Pathfinding 5.14.0 `find_path` / `plan_route!` / `move_along_route!` as the reference calls them
 * (src/agent_actions.jl:217-221, :231; pathfinder built at scripts/v0.4.jl:67, scripts/v0.6.jl:87-93).  Semantics
 * restated in SURVEY.md A-8: Moore neighbourhood in product order (x fastest), periodic wrap inherited from the
 * space, walkmap filter, integer costs (MaxDistance, or DirectDistance 10/14; PenaltyMap adds |Δelevation| to both
 * the edge cost and the heuristic), f = g + h, open list ordered by the tuple (f, x, y), parent replaced on strict
 * g improvement, path excludes the start cell.
 *
 * One search per thread.  The open list is a binary min-heap of u64 keys f<<32 | x<<16 | y (1-based x, y), whose
 * integer order IS the reference's tuple order, so the sequence of distinct expansions — and with it every parent
 * pointer — equals the reference's whatever the heap's internal layout.  Per-search state lives in a slot of
 * dense per-cell arrays in HBM (g: u32, meta: u16 = generation<<8 | closed<<4 | parent direction); the generation
 * stamp makes a slot reusable without clearing it.
 */
#ifndef ABM_ASTAR_H_
#define ABM_ASTAR_H_

#include "abm_kernels.h"

namespace abm {

enum { AS_CLOSED = 0x10, AS_NOPARENT = 0x0F };
enum { AS_FOUND = 1, AS_UNREACHABLE = 2, AS_HEAP_FULL = 3 };

struct AStarGrid {
  Geo g;
  const uint32_t* walkbits;
  const int16_t* elev;
  int32_t metric, penalty;
};

ABM_HD int64_t astar_cost(const AStarGrid& a, int ax, int ay, uint32_t ca, int bx, int by, uint32_t cb) {
  int64_t dx = ax > bx ? ax - bx : bx - ax, dy = ay > by ? ay - by : by - ay;
  if (a.g.periodic) {
    if (a.g.nx - dx < dx) dx = a.g.nx - dx;
    if (a.g.ny - dy < dy) dy = a.g.ny - dy;
  }
  const int64_t mx = dx > dy ? dx : dy, mn = dx > dy ? dy : dx;
  int64_t c = a.metric == ABM_ASTAR_MAX ? mx : 10 * (mx - mn) + 14 * mn;
  if (a.penalty) {
    const int64_t e = (int64_t)a.elev[ca] - (int64_t)a.elev[cb];
    c += e < 0 ? -e : e;
  }
  return c;
}

ABM_HD uint64_t astar_key(int64_t f, int x, int y) { return ((uint64_t)f << 32) | ((uint64_t)(x + 1) << 16) | (uint64_t)(y + 1); }

ABM_HD bool heap_push(uint64_t* h, uint32_t& n, uint32_t cap, uint64_t key) {
  if (n >= cap) return false;
  uint32_t i = n++;
  while (i > 0) {
    const uint32_t p = (i - 1) >> 1;
    const uint64_t pk = h[p];
    if (pk <= key) break;
    h[i] = pk;
    i = p;
  }
  h[i] = key;
  return true;
}
ABM_HD uint64_t heap_pop(uint64_t* h, uint32_t& n) {
  const uint64_t top = h[0];
  const uint64_t last = h[--n];
  uint32_t i = 0;
  while (true) {
    uint32_t c = 2 * i + 1;
    if (c >= n) break;
    uint64_t ck = h[c];
    if (c + 1 < n) { const uint64_t rk = h[c + 1]; if (rk < ck) { ck = rk; ++c; } }
    if (last <= ck) break;
    h[i] = ck;
    i = c;
  }
  if (n > 0) h[i] = last;
  return top;
}

/* one wave of searches: search i uses slot i */
struct AStarSearchBody {
  AStarGrid a;
  const uint32_t* src; /* tcell */
  const uint32_t* dst;
  uint32_t* gbuf;      /* slots x gt */
  uint16_t* meta;      /* slots x gt */
  uint64_t* heap;      /* slots x heap_cap */
  uint32_t heap_cap;
  uint32_t gen;        /* 1..255 */
  uint32_t* out_status;
  uint32_t* out_len;
  int64_t* out_cost;
  unsigned long long* expansions;

  ABM_HD void operator()(uint64_t i) const {
    const uint64_t gt = a.g.gt;
    uint32_t* g = gbuf + i * gt;
    uint16_t* m = meta + i * gt;
    uint64_t* h = heap + i * (uint64_t)heap_cap;
    const uint32_t s = src[i], t = dst[i];
    int sx, sy, tx, ty;
    dec(a.g, s, sx, sy);
    dec(a.g, t, tx, ty);
    uint32_t hn = 0;
    g[s] = 0;
    m[s] = (uint16_t)(gen << 8 | AS_NOPARENT);
    heap_push(h, hn, heap_cap, astar_key(astar_cost(a, sx, sy, s, tx, ty, t), sx, sy));
    uint32_t status = AS_UNREACHABLE;
    unsigned long long nexp = 0;
    while (hn > 0) {
      const uint64_t top = heap_pop(h, hn);
      const int x = (int)((top >> 16) & 0xFFFF) - 1, y = (int)(top & 0xFFFF) - 1;
      const uint32_t c = enc(a.g, x, y);
      if (c == t) { status = AS_FOUND; break; }
      const uint16_t mc = m[c];
      if (mc & AS_CLOSED) continue; /* stale duplicate of an expanded node: re-expansion could not improve anything */
      m[c] = (uint16_t)(mc | AS_CLOSED);
      ++nexp;
      const uint32_t gc = g[c];
      for (int d = 0; d < 8; ++d) {
        int ddx, ddy, ox, oy;
        off8(d, ddx, ddy);
        if (!shift(a.g, x, y, ddx, ddy, ox, oy)) continue;
        const uint32_t nc = enc(a.g, ox, oy);
        if (!((a.walkbits[nc >> 5] >> (nc & 31)) & 1u)) continue;
        const uint16_t mn = m[nc];
        const bool seen = (uint32_t)(mn >> 8) == gen;
        if (seen && (mn & AS_CLOSED)) continue;
        const int64_t ng = (int64_t)gc + astar_cost(a, x, y, c, ox, oy, nc);
        if (!seen || ng < (int64_t)g[nc]) {
          g[nc] = (uint32_t)ng;
          m[nc] = (uint16_t)(gen << 8 | d);
          if (!heap_push(h, hn, heap_cap, astar_key(ng + astar_cost(a, ox, oy, nc, tx, ty, t), ox, oy))) { status = AS_HEAP_FULL; hn = 0; break; }
        }
      }
    }
    uint32_t len = 0;
    if (status == AS_FOUND) { /* walk the parents back to the start (the start cell is not part of the path) */
      uint32_t c = t;
      int x = tx, y = ty;
      while (c != s) {
        const int d = m[c] & 0x0F;
        int ddx, ddy;
        off8(d, ddx, ddy);
        x = g_wrap(x - ddx, a.g.nx, a.g.periodic);
        y = g_wrap(y - ddy, a.g.ny, a.g.periodic);
        c = enc(a.g, x, y);
        ++len;
      }
      out_cost[i] = (int64_t)g[t];
    } else {
      out_cost[i] = -1;
    }
    out_status[i] = status;
    out_len[i] = len;
    if (nexp) atomic_add64(expansions, nexp);
  }
};

/* second pass of a wave: paths into the pool (first element = next cell to move to) */
struct AStarWriteBody {
  AStarGrid a;
  const uint32_t* src;
  const uint32_t* dst;
  const uint16_t* meta;
  const uint32_t* status;
  const uint32_t* len;
  const uint64_t* pool_off; /* per search */
  uint32_t* pool;

  ABM_HD void operator()(uint64_t i) const {
    if (status[i] != AS_FOUND) return;
    const uint16_t* m = meta + i * (uint64_t)a.g.gt;
    uint32_t* out = pool + pool_off[i];
    uint32_t k = len[i];
    uint32_t c = dst[i];
    const uint32_t s = src[i];
    int x, y;
    dec(a.g, c, x, y);
    while (c != s) {
      out[--k] = c;
      const int d = m[c] & 0x0F;
      int ddx, ddy;
      off8(d, ddx, ddy);
      x = g_wrap(x - ddx, a.g.nx, a.g.periodic);
      y = g_wrap(y - ddy, a.g.ny, a.g.periodic);
      c = enc(a.g, x, y);
    }
  }
};

/* the route store: pathfinder.agent_paths :: Dict{Int, Path} */
struct RouteStore {
  uint32_t* slot_of_id; /* id -> slot + 1, 0 = no route */
  uint64_t* r_off;      /* per slot: offset into pool */
  uint32_t* r_len;
  uint32_t* r_pos;      /* cells already consumed by move_along_route! */
  uint32_t* pool;
};

/* plan_route!: (re)binds the routes of a finished wave.  A new slot is taken only when the agent had none. */
struct RouteBindBody {
  RouteStore rs;
  int64_t* r_cost;          /* per slot: g(goal) at planning time (bookkeeping for abm_get_paths) */
  const uint32_t* agent_id; /* distinct within a batch */
  const uint32_t* status;
  const uint32_t* len;
  const int64_t* cost;
  const uint64_t* pool_off;
  uint32_t* slot_counter;

  ABM_HD void operator()(uint64_t i) const {
    if (status[i] != AS_FOUND) return; /* `isnothing(path) && return`: the previous route (if any) stays */
    const uint32_t id = agent_id[i];
    uint32_t s = rs.slot_of_id[id];
    if (s == 0) { s = atomic_add(slot_counter, 1u) + 1; rs.slot_of_id[id] = s; }
    rs.r_off[s - 1] = pool_off[i];
    rs.r_len[s - 1] = len[i];
    rs.r_pos[s - 1] = 0;
    r_cost[s - 1] = cost[i];
  }
};

struct GatherPlanBody { /* plan requests of the behaviour switches -> (agent id, start cell) */
  const uint32_t* cell; const uint32_t* id; const uint32_t* plan_agent;
  uint32_t* src; uint32_t* aid;
  ABM_HD void operator()(uint64_t i) const { const uint32_t k = plan_agent[i]; src[i] = cell[k]; aid[i] = id[k]; }
};

/* abm_plan_routes: host (id, goal x, goal y) -> (id, start cell, goal cell); flags unknown ids / bad goals */
struct PlanRequestBody {
  Geo g;
  const int64_t* hid; const int64_t* gx; const int64_t* gy;
  const uint32_t* idx_of_id; const uint32_t* id; const uint32_t* cell; uint32_t n; uint64_t next_id;
  uint32_t* aid; uint32_t* src; uint32_t* dst; uint32_t* err;
  ABM_HD void operator()(uint64_t i) const {
    const int64_t a = hid[i], x = gx[i], y = gy[i];
    aid[i] = 0; src[i] = 0; dst[i] = 0;
    if (a < 1 || (uint64_t)a >= next_id || x < 1 || x > g.nx || y < 1 || y > g.ny) { atomic_or(err, ERR_INTERNAL); return; }
    const uint32_t k = idx_of_id[a];
    if (k >= n || id[k] != (uint32_t)a) { atomic_or(err, ERR_INTERNAL); return; } /* not a living agent */
    aid[i] = (uint32_t)a; src[i] = cell[k]; dst[i] = enc(g, (int)(x - 1), (int)(y - 1));
  }
};

/* abm_get_paths, pass 1: remaining length and planning cost per requested id */
struct PathInfoBody {
  RouteStore rs; const int64_t* r_cost; const int64_t* hid; uint64_t ids_cap;
  int64_t* out_len; int64_t* out_cost;
  ABM_HD void operator()(uint64_t i) const {
    const int64_t a = hid[i];
    uint32_t s = 0;
    if (a >= 1 && (uint64_t)a <= ids_cap) s = rs.slot_of_id[a];
    if (!s) { out_len[i] = 0; out_cost[i] = -1; return; }
    out_len[i] = (int64_t)(rs.r_len[s - 1] - rs.r_pos[s - 1]);
    out_cost[i] = r_cost[s - 1];
  }
};
/* pass 2: 1-based coordinates of the remaining cells */
struct PathCopyBody {
  Geo g; RouteStore rs; const int64_t* hid; uint64_t ids_cap; const int64_t* offset;
  int64_t* px; int64_t* py;
  ABM_HD void operator()(uint64_t i) const {
    const int64_t a = hid[i];
    uint32_t s = 0;
    if (a >= 1 && (uint64_t)a <= ids_cap) s = rs.slot_of_id[a];
    if (!s) return;
    const uint32_t pos = rs.r_pos[s - 1], len = rs.r_len[s - 1];
    int64_t o = offset[i];
    for (uint32_t q = pos; q < len; ++q, ++o) {
      int x, y;
      dec(g, rs.pool[rs.r_off[s - 1] + q], x, y);
      px[o] = x + 1; py[o] = y + 1;
    }
  }
};

}  // namespace abm
#endif /* ABM_ASTAR_H_ */
