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
 * One search per WARP (a 32-thread CTA), searches taken from a device-side queue.  The open list is a 32-ary min-heap of
 * u64 keys f<<32 | x<<16 | y (1-based x, y), whose integer order IS the reference's tuple order — and the keys are
 * unique (a cell is pushed again only with a strictly smaller g, hence f) — so the sequence of distinct expansions,
 * and with it every parent pointer, equals the reference's whatever the heap's internal layout.  A node's 32 children
 * are one coalesced 256-byte load and their minimum two warp reductions: a pop is ~log32(n) dependent memory round
 * trips instead of ~2 log2(n); the 8 neighbours are loaded by 8 lanes while lane 0 loads the popped cell.  Per-search
 * state lives in a slot of dense per-cell arrays in HBM (g: u32, meta: u16 = generation<<8 | closed<<4 | parent
 * direction); the per-slot generation stamp makes a slot reusable without clearing it.  The finished search walks its
 * parents and writes the path straight into the route pool (space claimed with one 64-bit atomicAdd).
 */
#ifndef ABM_ASTAR_H_
#define ABM_ASTAR_H_

#include "abm_kernels.h"

namespace abm {

enum { AS_CLOSED = 0x10, AS_NOPARENT = 0x0F };
enum { AS_FOUND = 1, AS_UNREACHABLE = 2, AS_HEAP_FULL = 3, AS_POOL_FULL = 4 };

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

/* the same cost with the two elevations already loaded (the search kernel issues its loads up front) */
ABM_HD int64_t astar_cost_e(const AStarGrid& a, int ax, int ay, int ea, int bx, int by, int eb) {
  int64_t dx = ax > bx ? ax - bx : bx - ax, dy = ay > by ? ay - by : by - ay;
  if (a.g.periodic) {
    if (a.g.nx - dx < dx) dx = a.g.nx - dx;
    if (a.g.ny - dy < dy) dy = a.g.ny - dy;
  }
  const int64_t mx = dx > dy ? dx : dy, mn = dx > dy ? dy : dx;
  int64_t c = a.metric == ABM_ASTAR_MAX ? mx : 10 * (mx - mn) + 14 * mn;
  if (a.penalty) c += ea > eb ? ea - eb : eb - ea;
  return c;
}

/* the same in 32-bit arithmetic for the search kernel (g is a u32 per cell; one step costs below 2^20 + 2^16) */
ABM_HD uint32_t astar_cost32(const AStarGrid& a, int ax, int ay, int ea, int bx, int by, int eb) {
  int dx = ax > bx ? ax - bx : bx - ax, dy = ay > by ? ay - by : by - ay;
  if (a.g.periodic) {
    if (a.g.nx - dx < dx) dx = a.g.nx - dx;
    if (a.g.ny - dy < dy) dy = a.g.ny - dy;
  }
  const int mx = dx > dy ? dx : dy, mn = dx > dy ? dy : dx;
  int c = a.metric == ABM_ASTAR_MAX ? mx : 10 * (mx - mn) + 14 * mn;
  if (a.penalty) c += ea > eb ? ea - eb : eb - ea;
  return (uint32_t)c;
}

ABM_HD uint64_t astar_key(int64_t f, int x, int y) { return ((uint64_t)f << 32) | ((uint64_t)(x + 1) << 16) | (uint64_t)(y + 1); }

/* Connected components of the walkmap (Moore neighbourhood, the space's wrap): union-find with CAS hooking, then full
 * compression.  A goal in another component than the start — or not walkable at all — can never be popped, so the
 * search is known to end as "no path" without running it (the reference explores the whole component first and returns
 * nothing: same result). */
static const uint32_t AS_NOCOMP = 0xFFFFFFFFu; /* label of a cell that is not walkable */
ABM_HD uint32_t comp_find(const uint32_t* comp, uint32_t x) {
  uint32_t p;
  while ((p = comp[x]) != x) x = p;
  return x;
}
struct CompInitBody {
  const uint32_t* walkbits; uint32_t* comp;
  ABM_HD void operator()(uint64_t c) const { comp[c] = ((walkbits[c >> 5] >> (c & 31)) & 1u) ? (uint32_t)c : AS_NOCOMP; }
};
struct CompHookBody {
  Geo g; const uint32_t* walkbits; uint32_t* comp; uint32_t* changed;
  ABM_HD void operator()(uint64_t i) const {
    const uint32_t c = (uint32_t)i;
    if (!((walkbits[c >> 5] >> (c & 31)) & 1u)) return;
    int x, y;
    dec(g, c, x, y);
    for (int d = 0; d < 8; ++d) {
      int ddx, ddy, ox, oy;
      off8(d, ddx, ddy);
      if (!shift(g, x, y, ddx, ddy, ox, oy)) continue;
      const uint32_t nc = enc(g, ox, oy);
      if (!((walkbits[nc >> 5] >> (nc & 31)) & 1u)) continue;
      uint32_t a = comp_find(comp, c), b = comp_find(comp, nc);
      while (a != b) { /* hook the larger root under the smaller one; a root that was hooked meanwhile hands us its parent */
        if (a < b) { const uint32_t q = a; a = b; b = q; }
        const uint32_t old = atomic_cas(&comp[a], a, b);
        if (old == a) { *changed = 1u; break; }
        a = comp_find(comp, old);
        b = comp_find(comp, b);
      }
    }
  }
};
struct CompCompressBody {
  uint32_t* comp;
  ABM_HD void operator()(uint64_t c) const { if (comp[c] != AS_NOCOMP) comp[c] = comp_find(comp, (uint32_t)c); }
};

enum { AS_ARITY_LOG2 = 5, AS_ARITY = 32 };
static const unsigned long long AS_NOKEY = ~0ull;

enum { AS_TOP = 1 + AS_ARITY }; /* the root and its children live in shared memory for the length of a search */
enum { AS_L3 = AS_TOP + AS_ARITY * AS_ARITY }; /* first heap index whose parent lies outside shared memory */

/* sift-up of a 32-ary heap: one lane, ~1 parent visited in the common case (a new f is rarely below its parent's) */
ABM_HD bool heap_push(uint64_t* h, unsigned long long* top, uint32_t& n, uint32_t cap, uint64_t key) {
  if (n >= cap) return false;
  uint32_t i = n++;
  while (i > 0) {
    const uint32_t p = (i - 1) >> AS_ARITY_LOG2;
    const uint64_t pk = p < AS_TOP ? top[p] : h[p];
    if (pk <= key) break;
    if (i < AS_TOP) top[i] = pk; else h[i] = pk;
    i = p;
  }
  if (i < AS_TOP) top[i] = key; else h[i] = key;
  return true;
}

struct AStarShared {
  unsigned long long top[AS_TOP];  /* heap entries 0..32 */
  unsigned long long ck[AS_ARITY]; /* the children just loaded */
  unsigned long long best[2];      /* their minimum, alternating per level */
  unsigned long long nk[8];        /* keys the 8 neighbour lanes want pushed (0 = none) */
  uint32_t n_g[8]; int32_t n_e[8]; uint16_t n_m[8]; uint8_t n_ok[8]; /* what the neighbour lanes loaded */
  unsigned long long last, nexp;
  uint32_t hn, hn2, cur, lvl, sift, fin, skip, status, req, wipe, gen; /* hn: open-list size before the pop, hn2: after */
  uint32_t s, t; int32_t tx, ty, x, y, ec, et; uint32_t c, gc;
};

/* K8: grid = number of search slots; CTA `slot` keeps taking requests until the queue is empty */
struct AStarWarpCta {
  AStarGrid a;
  const uint32_t* src; /* tcell per request */
  const uint32_t* dst;
  const uint32_t* comp; /* nullable: component label per cell (AS_NOCOMP = not walkable) */
  const uint32_t* todo; /* nullable: the requests to run (indices into src/dst/out_*) */
  uint32_t n_todo;
  uint32_t* next_todo;  /* queue head */
  uint32_t* gbuf;      /* slots x gt */
  uint16_t* meta;      /* slots x gt */
  uint64_t* heap;      /* slots x heap_cap */
  uint32_t heap_cap;
  uint32_t* slot_gen;  /* per slot: last generation used (1..255) */
  uint32_t* out_status; uint32_t* out_len; int64_t* out_cost; uint64_t* out_off;
  uint32_t* pool; unsigned long long* pool_cursor; uint64_t pool_cap;
  unsigned long long* expansions;
  static size_t smem_bytes() { return sizeof(AStarShared); }

#ifndef ABM_EMU
  /* minimum of a 64-bit key over the warp, in every lane (two 32-bit reductions) */
  static __device__ __forceinline__ unsigned long long warp_min64(unsigned long long v) {
    const uint32_t hi = (uint32_t)(v >> 32);
    const uint32_t mhi = __reduce_min_sync(0xffffffffu, hi);
    const uint32_t mlo = __reduce_min_sync(0xffffffffu, hi == mhi ? (uint32_t)v : 0xFFFFFFFFu);
    return ((unsigned long long)mhi << 32) | mlo;
  }
  /* loads ptxas may not sink into the branches that use them (it does with plain ones): a pop issues all of its loads up
   * front and pays one memory round trip instead of three dependent ones */
  static __device__ __forceinline__ uint32_t ld_now(const uint32_t* p) { uint32_t v; asm volatile("ld.volatile.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v; }
  static __device__ __forceinline__ uint32_t ld_now(const uint16_t* p) { uint16_t v; asm volatile("ld.volatile.global.u16 %0, [%1];" : "=h"(v) : "l"(p) : "memory"); return v; }
  static __device__ __forceinline__ int ld_now(const int16_t* p) { int16_t v; asm volatile("ld.volatile.global.s16 %0, [%1];" : "=h"(v) : "l"(p) : "memory"); return (int)v; }
  static __device__ __forceinline__ unsigned long long ld_now(const uint64_t* p) { unsigned long long v; asm volatile("ld.volatile.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory"); return v; }
  /* The device form of the search below: warp-synchronous, every lane carries the search's scalars in registers (no staging
   * through shared memory, no CTA barriers), single-address loads are broadcast loads, lane 0 does the stores.  Only the top
   * 33 heap entries live in shared memory.  Same pops, same relaxations, same parents as the phase form, which stays as the
   * emulated build's version (tests/emu) and as the specification; the GPU tests compare this one with the oracle.
   * Per pop (instruction counts from the SASS, `cuobjdump -sass`): ~110 up to the one batch of loads (lanes 0..7 the
   * neighbours, lanes 8.. the popped cell, plus the heap's last entry), ~85 for the relaxations, a sift-down unrolled by where
   * the level lives (shared memory / boundary / heap array: 23 / 43 / 51 per level) that starts from the smallest NEW key when
   * there is one (pop + push in one pass), and a sift-up (~70) only for the second and later new keys, found with a ballot:
   * ~300 warp instructions per pop against ~700 of round 2's first warp form (1.45x the replans/s; profiles/README.md). */
  __device__ __forceinline__ void search_warp(uint32_t slot, uint32_t* smem) const {
    const unsigned FULL = 0xffffffffu;
    const int lane = (int)threadIdx.x;
    unsigned long long* top = ((AStarShared*)smem)->top;
    const uint64_t gt = a.g.gt;
    uint32_t* g = gbuf + (uint64_t)slot * gt;
    uint16_t* m = meta + (uint64_t)slot * gt;
    uint64_t* h = heap + (uint64_t)slot * (uint64_t)heap_cap;
    int ddx, ddy;
    off8(lane & 7, ddx, ddy); /* lanes 0..7: the lane's neighbour direction */
    while (true) {
      uint32_t req = 0xFFFFFFFFu, gen = 0, wipe = 0;
      if (lane == 0) {
        const uint32_t q = atomic_add(next_todo, 1u);
        if (q < n_todo) {
          req = todo ? todo[q] : q;
          gen = slot_gen[slot] + 1;
          if (gen > 255) { gen = 1; wipe = 1; }
          slot_gen[slot] = gen;
        }
      }
      req = __shfl_sync(FULL, req, 0); gen = __shfl_sync(FULL, gen, 0); wipe = __shfl_sync(FULL, wipe, 0);
      if (req == 0xFFFFFFFFu) break;
      if (wipe) { for (uint64_t k = (uint64_t)lane; k < gt; k += 32) m[k] = 0; } /* the 8-bit stamp wrapped: forget this slot's cells */
      __syncwarp();
      const uint32_t s = src[req], tc = dst[req];
      int sx, sy, tx, ty;
      dec(a.g, s, sx, sy);
      dec(a.g, tc, tx, ty);
      const int et = a.penalty ? (int)a.elev[tc] : 0;
      if (lane == 0) {
        g[s] = 0;
        m[s] = (uint16_t)(gen << 8 | AS_NOPARENT);
        top[0] = astar_key(astar_cost(a, sx, sy, s, tx, ty, tc), sx, sy);
      }
      uint32_t hn = 1, status = AS_UNREACHABLE;
      unsigned long long nexp = 0;
      if (comp && s != tc) { /* a goal the search could never pop: leave with an empty open list */
        const uint32_t ls = comp[s], lt = comp[tc];
        if (lt == AS_NOCOMP || (ls != AS_NOCOMP && ls != lt)) hn = 0;
      }
      __syncwarp();
      while (hn > 0) {
        /* pop: lanes 0..7 load everything about the popped cell's 8 neighbours, the other lanes the same four words of the
         * popped cell itself (one address, one transaction; lane 8 hands them round) — ONE memory round trip, together with
         * the heap's last entry */
        const unsigned long long root = top[0];
        const int x = (int)((root >> 16) & 0xFFFF) - 1, y = (int)(root & 0xFFFF) - 1;
        int ox = x, oy = y;
        bool inside = false;
        if (lane < 8) {
          inside = shift(a.g, x, y, ddx, ddy, ox, oy);
          if (!inside) { ox = x; oy = y; }
        }
        const uint32_t nc = enc(a.g, ox, oy);
        const uint32_t wb = ld_now(a.walkbits + (nc >> 5));
        const uint32_t mn = ld_now(m + nc);
        const uint32_t gn = ld_now(g + nc);
        const int en = a.penalty ? ld_now(a.elev + nc) : 0;
        const uint32_t hn2 = hn - 1;
        const unsigned long long last = hn2 < AS_TOP ? top[hn2] : ld_now(h + hn2);
        const uint32_t c = __shfl_sync(FULL, nc, 8);
        if (c == tc) { status = AS_FOUND; break; }
        const uint32_t mc = __shfl_sync(FULL, mn, 8);
        const uint32_t gc = __shfl_sync(FULL, gn, 8);
        const int ec = __shfl_sync(FULL, en, 8);
        unsigned long long key = 0;
        if (!(mc & AS_CLOSED)) { /* else: a stale duplicate of an expanded node */
          if (lane == 0) m[c] = (uint16_t)(mc | AS_CLOSED);
          nexp += 1;
          if (inside && ((wb >> (nc & 31)) & 1u)) {
            const bool seen = (mn >> 8) == gen;
            if (!(seen && (mn & AS_CLOSED))) {
              const uint32_t ng = gc + astar_cost32(a, x, y, ec, ox, oy, en);
              if (!seen || ng < gn) {
                g[nc] = ng;
                m[nc] = (uint16_t)(gen << 8 | (uint32_t)lane);
                key = ((unsigned long long)(ng + astar_cost32(a, ox, oy, en, tx, ty, et)) << 32) | ((uint32_t)(ox + 1) << 16) | (uint32_t)(oy + 1);
              }
            }
          }
        }
        /* What replaces the popped root: the smallest of the new keys when there is one (pop + push in one sift-down that
         * often ends near the top, since a new key is rarely far above the minimum — and one sift-up fewer), else the heap's last
         * entry.  The heap's internal layout differs from the phase form's; the pop sequence cannot (distinct keys). */
        uint32_t kmask = __ballot_sync(FULL, key != 0);
        unsigned long long place = last;
        uint32_t n = hn2;
        if (kmask) {
          place = warp_min64(key ? key : AS_NOKEY);
          kmask &= ~__ballot_sync(FULL, key == place);
          n = hn;
        }
        /* sift-down, one coalesced load of the 32 children and a warp-reduced minimum per level: level 0 reads its children from
         * shared memory, level 1 writes there, the deeper ones live in the slot's heap array */
        if (n > 0) {
          unsigned long long ck = (uint32_t)(1 + lane) < n ? top[1 + lane] : AS_NOKEY;
          unsigned long long b = warp_min64(ck);
          if (b == AS_NOKEY || place <= b) { if (lane == 0) top[0] = place; }
          else {
            if (lane == 0) top[0] = b;
            uint32_t cur = (uint32_t)__ffs((int)__ballot_sync(FULL, ck == b)); /* keys are unique: exactly one lane; 1 + its index */
            uint32_t ci = cur * AS_ARITY + 1 + (uint32_t)lane;
            ck = ci < n ? h[ci] : AS_NOKEY;
            b = warp_min64(ck);
            if (b == AS_NOKEY || place <= b) { if (lane == 0) top[cur] = place; }
            else {
              if (lane == 0) top[cur] = b;
              cur = cur * AS_ARITY + (uint32_t)__ffs((int)__ballot_sync(FULL, ck == b));
              while (true) {
                ci = cur * AS_ARITY + 1 + (uint32_t)lane;
                ck = ci < n ? h[ci] : AS_NOKEY;
                b = warp_min64(ck);
                if (b == AS_NOKEY || place <= b) { if (lane == 0) h[cur] = place; break; }
                if (lane == 0) h[cur] = b;
                cur = cur * AS_ARITY + (uint32_t)__ffs((int)__ballot_sync(FULL, ck == b));
              }
            }
          }
        }
        __syncwarp();
        /* the other new keys, in direction order: sift-up through the heap array (parents at index >= AS_TOP), then the entry
         * below shared memory, then shared memory */
        hn = n;
        bool full = false;
        while (kmask) {
          const int d = __ffs((int)kmask) - 1;
          kmask &= kmask - 1;
          const unsigned long long kd = __shfl_sync(FULL, key, d);
          if (hn >= heap_cap) { full = true; break; }
          uint32_t i = hn++;
          bool up = true;
          while (i >= AS_L3) {
            const uint32_t p = (i - 1) >> AS_ARITY_LOG2;
            const unsigned long long pk = h[p];
            if (pk <= kd) { up = false; break; }
            if (lane == 0) h[i] = pk;
            i = p;
          }
          if (up && i >= AS_TOP) {
            const uint32_t p = (i - 1) >> AS_ARITY_LOG2;
            const unsigned long long pk = top[p];
            if (pk <= kd) up = false;
            else { if (lane == 0) h[i] = pk; i = p; }
          }
          if (up && i >= 1 && i < AS_TOP) {
            const unsigned long long pk = top[0];
            if (pk > kd) { if (lane == 0) top[i] = pk; i = 0; }
          }
          if (lane == 0) { if (i < AS_TOP) top[i] = kd; else h[i] = kd; }
          __syncwarp();
        }
        if (full) { status = AS_HEAP_FULL; break; }
      }
      if (lane == 0) { /* walk the parents back to the start, claim pool space, write the path (first element = next cell to move to) */
        uint32_t len = 0;
        int64_t cost = -1;
        unsigned long long off = 0;
        if (status == AS_FOUND) {
          uint32_t c = tc;
          int x = tx, y = ty;
          while (c != s) {
            int ddx, ddy;
            off8(m[c] & 0x0F, ddx, ddy);
            x = g_wrap(x - ddx, a.g.nx, a.g.periodic);
            y = g_wrap(y - ddy, a.g.ny, a.g.periodic);
            c = enc(a.g, x, y);
            ++len;
          }
          cost = (int64_t)g[tc];
          off = atomic_add64(pool_cursor, (unsigned long long)len);
          if (off + len > pool_cap) status = AS_POOL_FULL;
          else {
            uint32_t k = len;
            c = tc; x = tx; y = ty;
            while (c != s) {
              pool[off + --k] = c;
              int ddx, ddy;
              off8(m[c] & 0x0F, ddx, ddy);
              x = g_wrap(x - ddx, a.g.nx, a.g.periodic);
              y = g_wrap(y - ddy, a.g.ny, a.g.periodic);
              c = enc(a.g, x, y);
            }
          }
        }
        out_status[req] = status; out_len[req] = len; out_cost[req] = cost; out_off[req] = off;
        if (nexp) atomic_add64(expansions, nexp);
      }
      __syncwarp();
    }
  }
#endif

  ABM_HD void operator()(uint32_t slot, uint32_t* smem) const {
#ifndef ABM_EMU
    search_warp(slot, smem);
    return;
#endif
    const int NT = AS_ARITY;
    AStarShared& sh = *(AStarShared*)smem;
    const uint64_t gt = a.g.gt;
    uint32_t* g = gbuf + (uint64_t)slot * gt;
    uint16_t* m = meta + (uint64_t)slot * gt;
    uint64_t* h = heap + (uint64_t)slot * (uint64_t)heap_cap;
    while (true) {
      ABM_CTA_FOR(t, NT) {
        if (t == 0) {
          const uint32_t q = atomic_add(next_todo, 1u);
          sh.req = q < n_todo ? (todo ? todo[q] : q) : 0xFFFFFFFFu;
          sh.wipe = 0;
          if (q < n_todo) {
            uint32_t gen = slot_gen[slot] + 1;
            if (gen > 255) { gen = 1; sh.wipe = 1; }
            slot_gen[slot] = gen; sh.gen = gen;
          }
        }
      }
      ABM_CTA_SYNC();
      if (sh.req == 0xFFFFFFFFu) break;
      if (sh.wipe) { /* the 8-bit stamp wrapped: forget this slot's cells */
        ABM_CTA_FOR(t, NT) { for (uint64_t k = (uint64_t)t; k < gt; k += NT) m[k] = 0; }
        ABM_CTA_SYNC();
      }
      ABM_CTA_FOR(t, NT) {
        if (t == 0) {
          const uint32_t s = src[sh.req], tc = dst[sh.req];
          int sx, sy, tx, ty;
          dec(a.g, s, sx, sy);
          dec(a.g, tc, tx, ty);
          sh.s = s; sh.t = tc; sh.tx = tx; sh.ty = ty; sh.et = a.penalty ? (int32_t)a.elev[tc] : 0;
          g[s] = 0;
          m[s] = (uint16_t)(sh.gen << 8 | AS_NOPARENT);
          sh.top[0] = astar_key(astar_cost(a, sx, sy, s, tx, ty, tc), sx, sy);
          sh.hn = 1; sh.status = AS_UNREACHABLE; sh.nexp = 0; sh.fin = 0;
          if (comp && s != tc) { /* a goal the search could never pop: leave with an empty open list */
            const uint32_t ls = comp[s], lt = comp[tc];
            if (lt == AS_NOCOMP || (ls != AS_NOCOMP && ls != lt)) sh.hn = 0;
          }
        }
      }
      ABM_CTA_SYNC();
      while (true) {
        /* pop: the root leaves, the last key waits to be placed; the popped cell is closed (or recognised as stale) by
         * lane 0 while lanes 8..15 load everything about its 8 neighbours — one memory round trip for both */
        ABM_CTA_FOR(t, NT) {
          if (sh.hn == 0) { if (t == 0) sh.fin = 1; }
          else {
            const uint64_t top = sh.top[0];
            const int x = (int)((top >> 16) & 0xFFFF) - 1, y = (int)(top & 0xFFFF) - 1;
            const uint32_t c = enc(a.g, x, y);
            if (t >= 8 && t < 16) {
              const int d = t - 8;
              int ddx, ddy, ox, oy;
              off8(d, ddx, ddy);
              const bool inside = shift(a.g, x, y, ddx, ddy, ox, oy);
              const uint32_t nc = inside ? enc(a.g, ox, oy) : c;
              const uint32_t wb = a.walkbits[nc >> 5];
              const uint16_t mn = m[nc];
              const uint32_t gn = g[nc];
              const int32_t en = a.penalty ? (int32_t)a.elev[nc] : 0;
              sh.n_ok[d] = (uint8_t)(inside && ((wb >> (nc & 31)) & 1u));
              sh.n_m[d] = mn; sh.n_g[d] = gn; sh.n_e[d] = en;
            }
            if (t == 0) {
              const uint16_t mc = m[c]; /* the three loads go out together */
              const uint32_t gc = g[c];
              const int32_t ec = a.penalty ? (int32_t)a.elev[c] : 0;
              const uint32_t hn = sh.hn - 1;
              sh.last = hn < AS_TOP ? sh.top[hn] : h[hn]; sh.hn2 = hn;
              sh.cur = 0; sh.lvl = 0; sh.sift = hn > 0 ? 1u : 0u; sh.skip = 0;
              sh.best[0] = AS_NOKEY; sh.best[1] = AS_NOKEY;
              if (c == sh.t) { sh.status = AS_FOUND; sh.fin = 1; }
              else if (mc & AS_CLOSED) sh.skip = 1; /* stale duplicate of an expanded node: re-expansion could not improve anything */
              else { m[c] = (uint16_t)(mc | AS_CLOSED); sh.nexp += 1; sh.gc = gc; sh.ec = ec; sh.x = x; sh.y = y; sh.c = c; }
            }
          }
        }
        ABM_CTA_SYNC();
        if (sh.fin) break;
        /* the 8 neighbours, one lane each (they are 8 different cells) */
        ABM_CTA_FOR(t, NT) {
          if (t < 8) {
            unsigned long long key = 0;
            if (!sh.skip && sh.n_ok[t]) {
              const uint16_t mn = sh.n_m[t];
              const bool seen = (uint32_t)(mn >> 8) == sh.gen;
              if (!(seen && (mn & AS_CLOSED))) {
                int ddx, ddy, ox, oy;
                off8(t, ddx, ddy);
                shift(a.g, sh.x, sh.y, ddx, ddy, ox, oy);
                const uint32_t nc = enc(a.g, ox, oy);
                const int32_t en = sh.n_e[t];
                const uint32_t ng = sh.gc + astar_cost32(a, sh.x, sh.y, sh.ec, ox, oy, en);
                if (!seen || ng < sh.n_g[t]) {
                  g[nc] = ng;
                  m[nc] = (uint16_t)(sh.gen << 8 | (uint32_t)t);
                  key = ((unsigned long long)(ng + astar_cost32(a, ox, oy, en, sh.tx, sh.ty, sh.et)) << 32) | ((uint32_t)(ox + 1) << 16) | (uint32_t)(oy + 1);
                }
              }
            }
            sh.nk[t] = key;
          }
        }
        ABM_CTA_SYNC();
        /* what replaces the popped root: the smallest new key when there is one (the open list then keeps its size and the
         * waiting last entry stays where it is), as in the device form */
        ABM_CTA_FOR(t, NT) {
          if (t == 0) {
            int dmin = -1;
            for (int d = 0; d < 8; ++d)
              if (sh.nk[d] && (dmin < 0 || sh.nk[d] < sh.nk[dmin])) dmin = d;
            if (dmin >= 0) {
              const unsigned long long k0 = sh.nk[dmin];
              for (int d = 0; d < 8; ++d) if (sh.nk[d] == k0) sh.nk[d] = 0;
              sh.last = k0; sh.hn2 = sh.hn; sh.sift = 1;
            }
          }
        }
        ABM_CTA_SYNC();
        /* sift-down of the waiting key: per level one coalesced load of the 32 children and a warp-reduced minimum */
        while (sh.sift) {
          const uint32_t lv = sh.lvl & 1u;
          ABM_CTA_FOR(t, NT) {
            const uint64_t ci = (uint64_t)sh.cur * AS_ARITY + 1 + (uint64_t)t;
            const unsigned long long key = ci < sh.hn2 ? (ci < AS_TOP ? sh.top[ci] : h[ci]) : AS_NOKEY;
            sh.ck[t] = key;
            warp_min64_into(&sh.best[lv], key, t);
          }
          ABM_CTA_SYNC();
          ABM_CTA_FOR(t, NT) {
            const unsigned long long b = sh.best[lv];
            const uint32_t cur = sh.cur;
            if (b == AS_NOKEY) { if (t == 0) { if (cur < AS_TOP) sh.top[cur] = sh.last; else h[cur] = sh.last; sh.sift = 0; } } /* a leaf */
            else if (sh.ck[t] == b) {                                               /* keys are unique: exactly one lane */
              const unsigned long long put = sh.last <= b ? sh.last : b;
              if (cur < AS_TOP) sh.top[cur] = put; else h[cur] = put;
              if (sh.last <= b) sh.sift = 0;
              else { sh.cur = cur * AS_ARITY + 1 + (uint32_t)t; sh.lvl += 1; sh.best[lv ^ 1u] = AS_NOKEY; }
            }
          }
          ABM_CTA_SYNC();
        }
        ABM_CTA_SYNC();
        ABM_CTA_FOR(t, NT) {
          if (t == 0) {
            uint32_t hn = sh.hn2;
            for (int d = 0; d < 8; ++d) {
              const unsigned long long key = sh.nk[d];
              if (key && !heap_push(h, sh.top, hn, heap_cap, key)) { sh.status = AS_HEAP_FULL; sh.fin = 1; break; }
            }
            sh.hn = hn;
          }
        }
        ABM_CTA_SYNC();
        if (sh.fin) break;
      }
      /* walk the parents back to the start (the start cell is not part of the path), claim pool space, write the path
       * (first element = next cell to move to) */
      ABM_CTA_FOR(t, NT) {
        if (t == 0) {
          const uint32_t r = sh.req;
          uint32_t len = 0, status = sh.status;
          int64_t cost = -1;
          unsigned long long off = 0;
          if (status == AS_FOUND) {
            uint32_t c = sh.t;
            int x = sh.tx, y = sh.ty;
            while (c != sh.s) {
              int ddx, ddy;
              off8(m[c] & 0x0F, ddx, ddy);
              x = g_wrap(x - ddx, a.g.nx, a.g.periodic);
              y = g_wrap(y - ddy, a.g.ny, a.g.periodic);
              c = enc(a.g, x, y);
              ++len;
            }
            cost = (int64_t)g[sh.t];
            off = atomic_add64(pool_cursor, (unsigned long long)len);
            if (off + len > pool_cap) status = AS_POOL_FULL;
            else {
              uint32_t k = len;
              c = sh.t; x = sh.tx; y = sh.ty;
              while (c != sh.s) {
                pool[off + --k] = c;
                int ddx, ddy;
                off8(m[c] & 0x0F, ddx, ddy);
                x = g_wrap(x - ddx, a.g.nx, a.g.periodic);
                y = g_wrap(y - ddy, a.g.ny, a.g.periodic);
                c = enc(a.g, x, y);
              }
            }
          }
          out_status[r] = status; out_len[r] = len; out_cost[r] = cost; out_off[r] = off;
          if (sh.nexp) atomic_add64(expansions, sh.nexp);
        }
      }
      ABM_CTA_SYNC();
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

/* scheduling hint: the heuristic distance of a request (long searches are started first, see plan_batch) */
struct AStarHintBody {
  AStarGrid a; const uint32_t* src; const uint32_t* dst; uint32_t* hint;
  ABM_HD void operator()(uint64_t i) const {
    int sx, sy, tx, ty;
    dec(a.g, src[i], sx, sy);
    dec(a.g, dst[i], tx, ty);
    const int64_t c = astar_cost_e(a, sx, sy, 0, tx, ty, 0);
    hint[i] = c > 0xFFFFFFFFll ? 0xFFFFFFFFu : (uint32_t)c;
  }
};

/* plan_route!: (re)binds the routes of a finished batch.  A new slot is taken only when the agent had none. */
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
