/*
 * abm_tile.h — the tile ("T") stepping path: the same tick as abm_kernels.h, organised around 32x32-cell cores.
 *
 * Why.  The rank-ordered semantics of the reference only ever ask, about a cell c and an agent
 * k: "is c occupied when k takes its turn?".  With by-id scheduling that is
 *        maxOld[c] > k                      some agent with a higher id started the tick on c (it has not moved yet)
 *     or minNew[c] < k                      some lower-id agent already ended its step on c and is still there
 * and it is still unknown iff neither holds and
 *        minMaybe[c] < k                    some lower-id agent whose step is not decided yet could end on c.
 * Three u32 per cell therefore replace the per-cell agent lists (GridSpace.stored_ids) for `ifempty` walks
 * (Agents.walk!), for the gregarious closest-neighbour search (src/agent_actions.jl:297-336) and, with a fourth
 * (min id of all arrivals), for first-come grass eating (src/agent_actions.jl:139-147).
 *
 * How.  Agents are grouped by 8x8-cell sub-tile (sub_off/sub_cnt give each group's run in the SoA).  One CTA owns
 * one core = one 32x32 grass tile; it stages the summaries of the core plus an 8-cell halo (a 48x48 window = 6x6
 * sub-tile runs) in shared memory with shared-memory atomics, and iterates the fixed point there.  Decision bytes
 * live in global memory: a decision is unique, so whichever CTA derives it first may publish it, and every CTA
 * re-reads them each round.  What a window cannot settle (chains that leave it) goes to a worklist that is
 * resolved by the same code over a small window centred on each agent, one warp-sized CTA per agent.
 * The commit kernel gathers by DESTINATION: the CTA of a core collects every agent whose final cell lies in it,
 * settles eating against the core's grass tile held in shared memory, applies energy and births, writes the
 * survivors grouped by sub-tile into a chunk of the output SoA taken with one atomicAdd, and regrows and writes
 * back the grass tile — one read and one write of every grass byte per tick.
 */
#ifndef ABM_TILE_H_
#define ABM_TILE_H_

#include "abm_kernels.h"

namespace abm {

enum { SUB = 8, CORE = 32, HALO = 8, WIN = CORE + 2 * HALO, MAX_RUNS = 36, TILE_THREADS = 256, COMMIT_THREADS = 128, AGENT_THREADS = 32 };
enum { NONE32 = 0xFFFFFFFFu };
/* shared-memory work lists of a core; larger populations take the streaming code.  The two capacities trade shared memory
 * (CTAs per SM) against how crowded a core may be before the slower code runs: overridable for tuning runs
 * (-DABM_ULIST_CAP=..., -DABM_ARRIVE_CAP=..., compared with tools/variant_probe.py). */
#ifndef ABM_ULIST_CAP
#define ABM_ULIST_CAP 1024
#endif
#ifndef ABM_ARRIVE_CAP
#define ABM_ARRIVE_CAP 384
#endif
enum { ULIST_CAP = ABM_ULIST_CAP, ARRIVE_CAP = ABM_ARRIVE_CAP, RUNOF_CAP = 2048, RUNTABLE_WORDS = 4 * MAX_RUNS + 1 + RUNOF_CAP / 4 };

struct SubIndex {
  const uint32_t* off; /* first agent of the sub-tile's run */
  const uint32_t* cnt;
  int32_t nsx, nsy;    /* sub-tile grid */
};

ABM_HD int floor_div8(int a) { return a >= 0 ? a >> 3 : -((-a + 7) >> 3); }

/* a window of cells [ox, ox+wx) x [oy, oy+wy) in unwrapped coordinates and the sub-tile runs that cover it */
struct Window {
  int ox, oy, wx, wy, ssx0, ssy0, nssx, nssy;
  int tlo, thi; /* summaries are complete on cells [tlo, thi]^2 (window coordinates): every agent that can reach them was taken in */
  ABM_HD void set(int ox_, int oy_, int wx_, int wy_) {
    ox = ox_; oy = oy_; wx = wx_; wy = wy_; tlo = 1; thi = wx_ - 2;
    ssx0 = floor_div8(ox); ssy0 = floor_div8(oy);
    nssx = floor_div8(ox + wx - 1) - ssx0 + 1;
    nssy = floor_div8(oy + wy - 1) - ssy0 + 1;
  }
  ABM_HD int runs() const { return nssx * nssy; }
  /* global sub-tile of run r, or -1 when it lies outside a non-periodic grid */
  ABM_HD int sub_of_run(int r, const Geo& g, const SubIndex& si) const {
    int sx = ssx0 + r % nssx, sy = ssy0 + r / nssx;
    if (g.periodic) sx = wrap_far(sx, si.nsx); else if (sx < 0 || sx >= si.nsx) return -1;
    if (g.py) sy = wrap_far(sy, si.nsy); else if (sy < 0 || sy >= si.nsy) return -1;
    return sy * si.nsx + sx;
  }
  /* window coordinates of an agent of run r standing on tcell c; false when outside the window */
  ABM_HD bool locate(int r, uint32_t c, int& x, int& y) const {
    x = (ssx0 + r % nssx) * 8 + (int)(c & 7) - ox;
    y = (ssy0 + r / nssx) * 8 + (int)((c >> 5) & 7) - oy;
    return x >= 0 && x < wx && y >= 0 && y < wy;
  }
};

/* shared-memory run table: which slice of the SoA each of the window's sub-tiles occupies */
struct RunTable {
  uint32_t* off;    /* MAX_RUNS */
  uint32_t* prefix; /* MAX_RUNS + 1 */
  uint32_t* cnt;    /* MAX_RUNS */
  uint32_t* xy0;    /* MAX_RUNS: window coordinates of the run's sub-tile origin, (x0 + 64) | (y0 + 64) << 16 */
  uint8_t* runof;   /* RUNOF_CAP: run of flat index m (valid when the window holds at most RUNOF_CAP agents) */
  ABM_HD void bind(uint32_t* p) { off = p; prefix = p + MAX_RUNS; cnt = prefix + MAX_RUNS + 1; xy0 = cnt + MAX_RUNS; runof = (uint8_t*)(xy0 + MAX_RUNS); }
  /* window coordinates of an agent of run r standing on tcell c; false when outside the window */
  ABM_HD bool locate(const Window& w, int r, uint32_t c, int& x, int& y) const {
    const uint32_t o = xy0[r];
    x = (int)(o & 0xFFFF) - 64 + (int)(c & 7);
    y = (int)(o >> 16) - 64 + (int)((c >> 5) & 7);
    return x >= 0 && x < w.wx && y >= 0 && y < w.wy;
  }
  ABM_HD int find(uint32_t m, int nruns) const { /* run containing flat index m */
    if (prefix[nruns] <= RUNOF_CAP) return runof[m];
    int lo = 0, hi = nruns;
    while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (prefix[mid] <= m) lo = mid; else hi = mid; }
    return lo;
  }
  /* phase 1 of 2 (then sync): slices of the window's sub-tiles */
  ABM_HD void load(int t, int NT, int nruns, const Window& w, const Geo& g, const SubIndex& si) const {
    for (int r = t; r < nruns; r += NT) {
      const int sub = w.sub_of_run(r, g, si);
      off[r] = sub < 0 ? 0u : si.off[sub];
      cnt[r] = sub < 0 ? 0u : si.cnt[sub];
      xy0[r] = (uint32_t)((w.ssx0 + r % w.nssx) * 8 - w.ox + 64) | ((uint32_t)((w.ssy0 + r / w.nssx) * 8 - w.oy + 64) << 16);
    }
  }
  /* phase 2 of 2 (then sync): every run sums the counts before it */
  ABM_HD void scan(int t, int NT, int nruns) const {
    for (int r = t; r < nruns; r += NT) {
      uint32_t run = 0;
      for (int q = 0; q <= r; ++q) run += cnt[q];
      prefix[r + 1] = run;
      if (r == 0) prefix[0] = 0;
    }
  }
  /* phase 3 (then sync): flat index -> run */
  ABM_HD void expand(int t, int NT, int nruns) const {
    if (prefix[nruns] > RUNOF_CAP) return;
    for (int r = t; r < nruns; r += NT)
      for (uint32_t m = prefix[r]; m < prefix[r + 1]; ++m) runof[m] = (uint8_t)r;
  }
};

/* n words (a multiple of 4, 16-byte aligned) set to `val` by the NT threads of a CTA, four words per store */
ABM_HD void fill_words(uint32_t* p, uint32_t val, int n, int t, int NT) {
#if defined(__CUDA_ARCH__)
  uint4* q = reinterpret_cast<uint4*>(p);
  const uint4 v4 = make_uint4(val, val, val, val);
  for (int c = t; c < (n >> 2); c += NT) q[c] = v4;
#else
  for (int c = t; c < n; c += NT) p[c] = val;
#endif
}

struct Summaries {
  uint32_t* max_old;
  uint32_t* min_new;
  uint32_t* min_maybe;
  ABM_HD bool occupied(int c, uint32_t id) const { const uint32_t a = max_old[c], b = min_new[c]; return (a > id) | (b < id); }
  ABM_HD bool unknown(int c, uint32_t id) const { return min_maybe[c] < id; }
};

enum { EV_PENDING = 0, EV_DECIDED = 1, EV_OUTSIDE = 2 };

/* what an agent's current decision byte contributes to the window's summaries */
ABM_HD void contribute(const Summaries& s, const Window& w, uint32_t id, int x, int y, uint8_t st) {
  const int dir = st & ST_DIRMASK;
  if (st & ST_DECIDED) {
    if (!present(st)) return; /* killed and left no newborn: occupies nothing */
    int fx = x, fy = y;
    if (st & ST_MOVED) { int dx, dy; off8(dir, dx, dy); fx += dx; fy += dy; }
    if (fx >= 0 && fx < w.wx && fy >= 0 && fy < w.wy) smem_min(&s.min_new[fy * w.wx + fx], id);
    return;
  }
  if (dir == DIR_GREG) { /* moves for sure, onto any of its 8 neighbours */
    for (int d = 0; d < 8; ++d) {
      int dx, dy;
      off8(d, dx, dy);
      const int fx = x + dx, fy = y + dy;
      if (fx >= 0 && fx < w.wx && fy >= 0 && fy < w.wy) smem_min(&s.min_maybe[fy * w.wx + fx], id);
    }
    return;
  }
  smem_min(&s.min_maybe[y * w.wx + x], id); /* `ifempty` walk: stays or steps onto its target */
  int dx, dy;
  off8(dir, dx, dy);
  const int fx = x + dx, fy = y + dy;
  if (fx >= 0 && fx < w.wx && fy >= 0 && fy < w.wy) smem_min(&s.min_maybe[fy * w.wx + fx], id);
}

/* the cells an undecided agent may still end on, into one min_maybe buffer */
ABM_HD void contribute_maybe(uint32_t* min_maybe, const Window& w, uint32_t id, int x, int y, uint8_t st) {
  const int dir = st & ST_DIRMASK;
  if (dir == DIR_GREG) {
    for (int d = 0; d < 8; ++d) {
      int dx, dy;
      off8(d, dx, dy);
      const int fx = x + dx, fy = y + dy;
      if (fx >= 0 && fx < w.wx && fy >= 0 && fy < w.wy) smem_min(&min_maybe[fy * w.wx + fx], id);
    }
    return;
  }
  smem_min(&min_maybe[y * w.wx + x], id);
  int dx, dy;
  off8(dir, dx, dy);
  const int fx = x + dx, fy = y + dy;
  if (fx >= 0 && fx < w.wx && fy >= 0 && fy < w.wy) smem_min(&min_maybe[fy * w.wx + fx], id);
}
/* a decided, still present agent: its final cell */
ABM_HD void contribute_final(uint32_t* min_new, const Window& w, uint32_t id, int x, int y, uint8_t st) {
  if (!present(st)) return;
  if (st & ST_MOVED) { int dx, dy; off8(st & ST_DIRMASK, dx, dy); x += dx; y += dy; }
  if (x >= 0 && x < w.wx && y >= 0 && y < w.wy) smem_min(&min_new[y * w.wx + x], id);
}

enum { GREG_NOBODY = 252, GREG_PENDING = 253, GREG_OUTSIDE = 254 };

/* closest neighbour of a gregarious agent: strict `<` over nearby_ids order == the first occupied cell in (d^2, β index)
 * order.  Returns its box index, GREG_NOBODY, or why that cannot be known yet. */
ABM_HD int greg_scan(const View& v, const Summaries& s, const Window& w, uint32_t id, int x, int y) {
  const int r = v.p.r, side = 2 * r + 1;
  if (x - r < w.tlo || x + r > w.thi || y - r < w.tlo || y + r > w.thi) return GREG_OUTSIDE;
  const int cells = side * side;
  for (int q = 0; q < cells; ++q) {
    const int c = (y + v.greg_dxy[2 * q + 1]) * w.wx + (x + v.greg_dxy[2 * q]);
    if (s.occupied(c, id)) return v.greg_order[q];
    if (s.unknown(c, id)) return GREG_PENDING;
  }
  return GREG_NOBODY;
}
/* the draw that follows: wsample over the 8 neighbours (global RNG routed to the agent's stream, draw 1), or
 * rand(model.rng, 1:8) when nobody is in view (:334); move_agent! is unconditional (:331) */
ABM_HD uint8_t greg_finish(const View& v, uint32_t id, uint8_t st, int code) {
  AgentStream rs;
  rs.init(v.p.seed, v.p.tick, id);
  rs.n = 1; /* draw 0 was the execute_walk! uniform */
  int d;
  if (code == GREG_NOBODY) {
    d = (int)rs.next_below(8);
  } else {
    const double* cw = v.greg_table + code * 8;
    double c8[8];
    for (int i = 0; i < 8; ++i) c8[i] = cw[i]; /* independent loads: one memory latency instead of a chain */
    const double t = rs.next_f64() * c8[7]; /* StatsBase.sample(rng, wv): t = rand(rng) * sum(wv) */
    /* i = 1; cw = w[1]; while cw < t && i < n: i += 1; cw += w[i] — the cumulative sums never decrease, so the stopping
     * index is the number of the first seven sums below t */
    d = 0;
    for (int i = 0; i < 7; ++i) d += c8[i] < t ? 1 : 0;
  }
  return (uint8_t)((st & (ST_DEAD | ST_BIRTH)) | d | ST_DECIDED | ST_MOVED);
}

/* one attempt to decide agent (order key id, RNG stream rng_id, st) at window cell (x, y); summaries are complete on [1, w-2]^2 */
ABM_HD int evaluate(const View& v, const Summaries& s, const Window& w, uint32_t id, uint32_t rng_id, int x, int y, uint8_t st, uint8_t& out) {
  const int dir = st & ST_DIRMASK;
  if (dir != DIR_GREG) { /* walk!(...; ifempty = true) towards old + dir */
    int dx, dy;
    off8(dir, dx, dy);
    const int tx = x + dx, ty = y + dy;
    if (tx < w.tlo || tx > w.thi || ty < w.tlo || ty > w.thi) return EV_OUTSIDE;
    const int t = ty * w.wx + tx;
    if (s.occupied(t, id)) { out = (uint8_t)(st | ST_DECIDED); return EV_DECIDED; }
    if (s.unknown(t, id)) return EV_PENDING;
    out = (uint8_t)(st | ST_DECIDED | ST_MOVED);
    return EV_DECIDED;
  }
  /* random_walk_gregarious (src/agent_actions.jl:297-336) */
  const int code = greg_scan(v, s, w, id, x, y);
  if (code == GREG_OUTSIDE) return EV_OUTSIDE;
  if (code == GREG_PENDING) return EV_PENDING;
  out = greg_finish(v, rng_id, st, code); /* the draw comes from the agent's own stream, whatever its place in the schedule */
  return EV_DECIDED;
}

/*
 * Kernel K_A / K_F.  mode 0: CTA b resolves core b (48x48 window, every agent it can settle, up to max_rounds
 * rounds) and lists the core's agents that stay undecided.  mode 1: CTA b settles pending agent in[b] alone, over
 * the smallest window that holds everything its rule looks at, and re-lists it when that is not yet possible.
 */
struct ResolveWindow {
  View v;
  SubIndex si;
  int32_t mode, cores_x, core_y0, max_rounds, nthreads;
  int32_t row_step;         /* mode 0: CTA row q handles core row core_y0 + q * row_step (1 = consecutive rows) */
  int32_t margin;           /* mode 0: agents farther than this from the core are left out of the window (<= HALO) */
  uint32_t grid;            /* mode 1: number of CTAs launched */
  const uint32_t* in;
  const uint32_t* in_count; /* mode 1: device-side length of `in` */
  uint32_t* out;
  uint32_t* out_count;

  static size_t smem_bytes(int wx, int wy) { return (size_t)(3 * wx * wy + RUNTABLE_WORDS + 8) * 4; }
  /* One min_maybe buffer, cleared and rebuilt from the open entries every round: 9 KB less shared memory per CTA than
   * the double buffer (5 CTAs/SM instead of 4) for a third barrier and a rebuild pass per round — measured 19 % faster
   * (resolve_tiles 0.396 vs 0.489 ms, C2, same box).  -DABM_TWO_MAYBE brings the double buffer back. */
#ifdef ABM_TWO_MAYBE
  enum { MAYBE_BUFS = 2 };
#else
  enum { MAYBE_BUFS = 1 };
#endif
  enum { GREG_REL_WORDS = 120 }; /* (2 * GREG_MAX_R + 1)^2 = 225 window offsets of the box cells in visiting order, int16 */
  static size_t smem_bytes_core() { return (size_t)((2 + MAYBE_BUFS) * WIN * WIN + 3 * ULIST_CAP + RUNTABLE_WORDS + 16 + GREG_REL_WORDS) * 4; }

  /* Core window with the undecided agents held in a shared-memory work list: each round touches only what is still
   * open.  min_new only ever gains entries (decisions are final), min_maybe is rebuilt per round from the list.  Agents are read from global memory exactly once.  Returns false (uniformly) when the list overflows. */
  ABM_HD bool core_fast(uint32_t cta, uint32_t* smem) const {
    ABM_PHASE_BEGIN(); /* -DABM_PHASE_CLOCKS builds only (tools/phase_probe.py) */
    Window w;
    const int cx = (int)(cta % (uint32_t)cores_x), cy = (int)(cta / (uint32_t)cores_x) * row_step + core_y0;
    w.set(cx * CORE - HALO, cy * CORE - HALO, WIN, WIN);
    const int in_lo = HALO - margin, in_hi = HALO + CORE + margin; /* agents taken in: [in_lo, in_hi)^2 */
    w.tlo = in_lo + 1; w.thi = in_hi - 2;
    const int cells = WIN * WIN, nruns = w.runs(), NT = TILE_THREADS;
    uint32_t* max_old = smem;
    uint32_t* min_new = smem + cells;
    uint32_t* maybe[2] = {smem + 2 * cells, smem + (1 + MAYBE_BUFS) * cells};
    /* one work list, never compacted: `ifempty` walkers from the front, gregarious agents from the back (each kind is
     * then evaluated by full warps; their code paths differ by an order of magnitude in length).  Entry = agent index,
     * id, and a word  best:9 | done:1 | state:8 | y:6 | x:6  — `best` is the running minimum of the gregarious search
     * (visiting rank << 1 | only-maybe), in the top bits so that atomicMin on the word orders by it. */
    uint32_t* l_a = smem + (2 + MAYBE_BUFS) * cells;
    uint32_t* l_id = l_a + ULIST_CAP;
    uint32_t* l_xs = l_id + ULIST_CAP;
    RunTable rt;
    rt.bind(l_xs + ULIST_CAP);
    uint32_t* cnt = l_xs + ULIST_CAP + RUNTABLE_WORDS; /* [0] walkers, [5] gregarious,
                                                        * [8 + 2p], [9 + 2p] progress / open agents of the core, round parity p */
    int16_t* grel = (int16_t*)(cnt + 16);              /* the box cells in visiting order, as offsets inside the window */
    const uint32_t XS_DONE = 1u << 20, XS_BEST_NONE = 0x1FFu << 21, XS_LOW = (1u << 21) - 1;

    const int side = 2 * v.p.r + 1, box = side * side;
    ABM_CTA_FOR(t, NT) {
      rt.load(t, NT, nruns, w, v.g, si);
      fill_words(max_old, 0u, cells, t, NT);
      fill_words(min_new, NONE32, (1 + MAYBE_BUFS) * cells, t, NT); /* min_new and the maybe buffer(s) are contiguous */
      if (t < 16) cnt[t] = 0;
      if (v.p.walk_type == ABM_RANDOM_WALK_GREGARIOUS)
        for (int q = t; q < box; q += NT) grel[q] = (int16_t)((int)v.greg_dxy[2 * q + 1] * WIN + (int)v.greg_dxy[2 * q]);
    }
    ABM_CTA_SYNC();
    ABM_CTA_FOR(t, NT) { rt.scan(t, NT, nruns); }
    ABM_CTA_SYNC();
    ABM_CTA_FOR(t, NT) { rt.expand(t, NT, nruns); }
    ABM_CTA_SYNC();
    ABM_PHASE(0); /* setup: fills, run table */
    const uint32_t M = rt.prefix[nruns];
    ABM_CTA_FOR(t, NT) {
      for (uint32_t m0 = (uint32_t)t; m0 < M; m0 += 3u * (uint32_t)NT) {
       int rr[3]; uint32_t aa[3], cc[3], ii[3]; uint8_t ss[3];
       for (int j = 0; j < 3; ++j) { /* all global loads of up to three agents in flight before any is used */
        const uint32_t m = m0 + (uint32_t)(j * NT);
        if (m >= M) { rr[j] = -1; continue; }
        rr[j] = rt.find(m, nruns);
        aa[j] = rt.off[rr[j]] + (m - rt.prefix[rr[j]]);
        cc[j] = v.cell[aa[j]]; ii[j] = v.okey[aa[j]]; ss[j] = load_state(&v.state[aa[j]]); /* the list and the summaries hold order keys */
       }
       for (int j = 0; j < 3; ++j) {
        if (rr[j] < 0) continue;
        int x, y;
        if (!rt.locate(w, rr[j], cc[j], x, y)) continue;
        if (x < in_lo || x >= in_hi || y < in_lo || y >= in_hi) continue;
        const uint32_t id = ii[j];
        const uint8_t st = ss[j];
        smem_max(&max_old[y * WIN + x], id);
        if (st & ST_DECIDED) { contribute_final(min_new, w, id, x, y, st); continue; }
        const bool greg = (st & ST_DIRMASK) == DIR_GREG;
        const uint32_t k = smem_add(&cnt[greg ? 5 : 0], 1u);
        if (k >= ULIST_CAP) continue; /* overflow: detected below from the two counters, the list is abandoned */
        const uint32_t slot = greg ? ULIST_CAP - 1 - k : k;
        l_a[slot] = aa[j]; l_id[slot] = id; l_xs[slot] = (uint32_t)x | ((uint32_t)y << 6) | ((uint32_t)st << 12) | XS_BEST_NONE;
        if (!greg) contribute_maybe(maybe[0], w, id, x, y, st);
       }
      }
    }
    ABM_CTA_SYNC();
    ABM_PHASE(1); /* agents loaded and listed */
    if (cnt[0] + cnt[5] > ULIST_CAP) return false; /* the two ends of the list met: take the streaming code */
    const uint32_t nc = cnt[0], ng = cnt[5];
    ABM_CTA_FOR(t, NT) { /* the gregarious agents' eight candidate cells, by full warps */
      for (uint32_t u = (uint32_t)t; u < ng; u += (uint32_t)NT) {
        const uint32_t e = ULIST_CAP - 1 - u, xs = l_xs[e];
        contribute_maybe(maybe[0], w, l_id[e], (int)(xs & 63), (int)((xs >> 6) & 63), (uint8_t)(xs >> 12));
      }
    }
    ABM_CTA_SYNC();
    /* Rounds, three barriers each.  P1: every open agent is evaluated by its own thread (walkers from the front of the
     * thread range, gregarious agents from the back, so that each kind fills whole warps: their code paths differ by an
     * order of magnitude in length) and publishes its decision at once.  P2: min_maybe, which nobody reads in this phase,
     * is cleared.  P3: the entries still open put their candidate cells back into min_maybe.  min_new only gains entries
     * (decisions are final); a reader that misses a fresh entry still finds the same id in min_maybe and waits a round.
     * (With -DABM_TWO_MAYBE the open entries write into a second buffer during P1 instead and P3 disappears.) */
    ABM_PHASE(2); /* candidate cells of the gregarious agents */
    int cur = 0;
    for (int round = 0; round < max_rounds; ++round) {
      uint32_t* flag = cnt + 8 + 2 * (round & 1);
      uint32_t* flag_next = cnt + 8 + 2 * ((round + 1) & 1);
      Summaries s;
      s.max_old = max_old; s.min_new = min_new; s.min_maybe = maybe[cur];
      ABM_CTA_FOR(t, NT) {
        for (uint32_t u = (uint32_t)t; u < nc; u += (uint32_t)NT) { /* `ifempty` walkers: three loads and a compare */
          const uint32_t xs = l_xs[u];
          if (xs & XS_DONE) continue;
          const uint32_t id = l_id[u];
          const int x = (int)(xs & 63), y = (int)((xs >> 6) & 63);
          const uint8_t st = (uint8_t)(xs >> 12);
          uint8_t nst = st;
          if (evaluate(v, s, w, id, id, x, y, st, nst) == EV_DECIDED) { /* a walker draws nothing here */
            v.state[l_a[u]] = nst;
            contribute_final(min_new, w, id, x, y, nst);
            l_xs[u] = xs | XS_DONE;
            flag[0] = 1;
          } else {
            if (MAYBE_BUFS == 2) contribute_maybe(maybe[cur ^ 1], w, id, x, y, st);
            if (x >= HALO && x < HALO + CORE && y >= HALO && y < HALO + CORE) flag[1] = 1;
          }
        }
        /* gregarious closest-neighbour search (src/agent_actions.jl:297-336): GREG_LANES lanes per agent walk the box cells
         * nearest first, GREG_LANES at a time; the first chunk with an occupied cell (the closest neighbour: draw and move) or
         * a still unknown one (wait) ends the search.  The lanes of a warp leave the loop together, so that the tail (the
         * draw, one lane per agent) runs converged. */
        for (uint32_t ubase = 0; ubase < ng; ubase += (uint32_t)(NT / GREG_LANES)) {
          const uint32_t u = ubase + (uint32_t)(t / GREG_LANES);
          const int lane_in = t % GREG_LANES;
          const uint32_t e = ULIST_CAP - 1 - (u < ng ? u : 0u), xs = l_xs[e];
          const bool open = u < ng && !(xs & XS_DONE);
          const uint32_t id = l_id[e];
          const int x = (int)(xs & 63), y = (int)((xs >> 6) & 63);
          const uint8_t st = (uint8_t)(xs >> 12);
          const bool trusted = !(x - v.p.r < w.tlo || x + v.p.r > w.thi || y - v.p.r < w.tlo || y + v.p.r > w.thi);
          const bool search = open && trusted; /* an untrusted box: the agent stays open */
          const int base = y * WIN + x;
          uint32_t found = NONE32; /* visiting rank << 1 | only-maybe of the group's first hit */
          for (int q0 = 0; q0 < box; q0 += GREG_LANES) {
            const int q = q0 + lane_in;
            uint32_t mine = NONE32;
            if (search && q < box) {
              const int c = base + (int)grel[q];
              if (s.occupied(c, id)) mine = (uint32_t)q << 1;
              else if (s.unknown(c, id)) mine = ((uint32_t)q << 1) | 1u;
            }
            const uint32_t gm = group_min(mine); /* every lane of the warp takes part in the shuffles of every chunk ... */
            if (found == NONE32) found = gm;     /* ... but a group keeps its first hit while the others go on */
            if (warp_all(!search || found != NONE32)) break;
          }
          if (search && lane_in == 0 && !(found != NONE32 && (found & 1u))) {
            const uint32_t rng_id = v.okey == v.id ? id : v.id[l_a[e]];
            const uint8_t nst = greg_finish(v, rng_id, st, found == NONE32 ? (int)GREG_NOBODY : (int)v.greg_order[found >> 1]);
            v.state[l_a[e]] = nst;
            contribute_final(min_new, w, id, x, y, nst);
            l_xs[e] = xs | XS_DONE;
            flag[0] = 1;
          } else if (open && lane_in == 0) {
            if (MAYBE_BUFS == 2) contribute_maybe(maybe[cur ^ 1], w, id, x, y, st);
            if (x >= HALO && x < HALO + CORE && y >= HALO && y < HALO + CORE) flag[1] = 1;
          }
        }
      }
      ABM_CTA_SYNC();
      ABM_CTA_FOR(t, NT) { /* what this round read is dead now: make it the clean target of the next round */
        fill_words(maybe[cur], NONE32, cells, t, NT);
        if (t == 0) { flag_next[0] = 0; flag_next[1] = 0; }
      }
      ABM_CTA_SYNC();
      cur ^= (MAYBE_BUFS == 2);
      if (flag[1] == 0 || flag[0] == 0) break;
      if (MAYBE_BUFS == 1) { /* the single buffer was cleared above: rebuild it from the entries still open */
        ABM_CTA_FOR(t, NT) {
          for (uint32_t u = (uint32_t)t; u < nc + ng; u += (uint32_t)NT) {
            const uint32_t e = u >= nc ? ULIST_CAP - 1 - (u - nc) : u;
            const uint32_t xs = l_xs[e];
            if (xs & XS_DONE) continue;
            contribute_maybe(maybe[0], w, l_id[e], (int)(xs & 63), (int)((xs >> 6) & 63), (uint8_t)(xs >> 12));
          }
        }
        ABM_CTA_SYNC();
      }
    }
    ABM_CTA_SYNC();
    ABM_PHASE(3); /* rounds */
    ABM_CTA_FOR(t, NT) { /* the core's agents that stay open go to the per-agent windows */
      for (uint32_t u = (uint32_t)t; u < nc + ng; u += (uint32_t)NT) {
        const uint32_t e = u >= nc ? ULIST_CAP - 1 - (u - nc) : u;
        const uint32_t xs = l_xs[e];
        if (xs & XS_DONE) continue;
        const int x = (int)(xs & 63), y = (int)((xs >> 6) & 63);
        if (x >= HALO && x < HALO + CORE && y >= HALO && y < HALO + CORE) out[atomic_add(out_count, 1u)] = l_a[e];
      }
    }
    return true;
  }

  ABM_HD void operator()(uint32_t cta, uint32_t* smem) const {
    if (mode == 0) {
      if (core_fast(cta, smem)) return;
      ABM_CTA_SYNC();
      window_body(cta, NONE32, smem);
      return;
    }
    /* per-agent windows: the list length lives on the device (the host launches an upper bound), grid-stride over it */
    const uint32_t n_in = *in_count;
    for (uint32_t b = cta; b < n_in; b += grid) {
      window_body(0, in[b], smem);
      ABM_CTA_SYNC();
    }
  }

  ABM_HD void window_body(uint32_t cta, uint32_t target, uint32_t* smem) const {
    Window w;
    int own_lo = 0, own_hi = 0;
    if (mode == 0) {
      const int cx = (int)(cta % (uint32_t)cores_x), cy = (int)(cta / (uint32_t)cores_x) * row_step + core_y0;
      w.set(cx * CORE - HALO, cy * CORE - HALO, WIN, WIN);
      own_lo = HALO; own_hi = HALO + CORE;
    } else {
      int x, y;
      dec(v.g, v.cell[target], x, y);
      const int q = (v.state[target] & ST_DIRMASK) == DIR_GREG ? v.p.r + 1 : 2;
      w.set(x - q, y - q, 2 * q + 1, 2 * q + 1);
    }
    const int cells = w.wx * w.wy, nruns = w.runs(), NT = nthreads;
    Summaries s;
    s.max_old = smem; s.min_new = smem + cells; s.min_maybe = smem + 2 * cells;
    RunTable rt;
    rt.bind(smem + 3 * cells);
    uint32_t* flags = smem + 3 * cells + RUNTABLE_WORDS; /* [0] progress, [1] undecided agents this CTA answers for */

    ABM_CTA_FOR(t, NT) {
      rt.load(t, NT, nruns, w, v.g, si);
      for (int c = t; c < cells; c += NT) s.max_old[c] = 0u;
    }
    ABM_CTA_SYNC();
    ABM_CTA_FOR(t, NT) { rt.scan(t, NT, nruns); }
    ABM_CTA_SYNC();
    ABM_CTA_FOR(t, NT) { rt.expand(t, NT, nruns); }
    ABM_CTA_SYNC();
    const uint32_t M = rt.prefix[nruns];
    ABM_CTA_FOR(t, NT) { /* static part: highest id that starts the tick on each cell */
      for (uint32_t m = (uint32_t)t; m < M; m += (uint32_t)NT) {
        const int r = rt.find(m, nruns);
        const uint32_t a = rt.off[r] + (m - rt.prefix[r]);
        int x, y;
        if (rt.locate(w, r, v.cell[a], x, y)) smem_max(&s.max_old[y * w.wx + x], v.okey[a]);
      }
    }
    ABM_CTA_SYNC();

    for (int round = 0; round < max_rounds; ++round) {
      ABM_CTA_FOR(t, NT) {
        for (int c = t; c < cells; c += NT) { s.min_new[c] = NONE32; s.min_maybe[c] = NONE32; }
        if (t == 0) { flags[0] = 0; flags[1] = 0; }
      }
      ABM_CTA_SYNC();
      ABM_CTA_FOR(t, NT) {
        for (uint32_t m = (uint32_t)t; m < M; m += (uint32_t)NT) {
          const int r = rt.find(m, nruns);
          const uint32_t a = rt.off[r] + (m - rt.prefix[r]);
          int x, y;
          if (rt.locate(w, r, v.cell[a], x, y)) contribute(s, w, v.okey[a], x, y, load_state(&v.state[a]));
        }
      }
      ABM_CTA_SYNC();
      ABM_CTA_FOR(t, NT) {
        for (uint32_t m = (uint32_t)t; m < M; m += (uint32_t)NT) {
          const int r = rt.find(m, nruns);
          const uint32_t a = rt.off[r] + (m - rt.prefix[r]);
          if (mode == 1 && a != target) continue;
          const uint8_t st = load_state(&v.state[a]);
          if (st & ST_DECIDED) continue;
          int x, y;
          if (!rt.locate(w, r, v.cell[a], x, y)) continue;
          uint8_t nst = st;
          const int res = evaluate(v, s, w, v.okey[a], v.id[a], x, y, st, nst);
          if (res == EV_DECIDED) { v.state[a] = nst; flags[0] = 1; }
          else if (mode == 1 || (x >= own_lo && x < own_hi && y >= own_lo && y < own_hi)) smem_add(&flags[1], 1u);
        }
      }
      ABM_CTA_SYNC();
      const uint32_t progress = flags[0], left = flags[1];
      ABM_CTA_SYNC();
      if (left == 0 || progress == 0) break;
    }
    ABM_CTA_FOR(t, NT) { /* hand what is still open to the next pass */
      for (uint32_t m = (uint32_t)t; m < M; m += (uint32_t)NT) {
        const int r = rt.find(m, nruns);
        const uint32_t a = rt.off[r] + (m - rt.prefix[r]);
        if (mode == 1 && a != target) continue;
        if (load_state(&v.state[a]) & ST_DECIDED) continue;
        int x, y;
        if (!rt.locate(w, r, v.cell[a], x, y)) continue;
        if (mode == 1 || (x >= own_lo && x < own_hi && y >= own_lo && y < own_hi)) out[atomic_add(out_count, 1u)] = a;
      }
    }
  }
};

/*
 * Kernel K_BC: commit by destination + grass + re-grouping.  reduce_energy! / eat! / reproduce!
 * (src/agent_actions.jl:54-67, :85-92, :139-147, :172-178) and custom_model_step! (src/model_actions.jl:18-30).
 */
enum { MIG_CAP = 96 }; /* migrants listed per core and tick; a core that receives more falls back to scanning its halo */

/* After every decision is final: agents whose step crosses a core border are listed at their destination core, so that
 * the commit of a core reads its own agents plus this list instead of scanning the 20 halo sub-tiles around it.
 * One thread per agent (own and ghost). */
struct MigrantScanBody {
  Geo g; int32_t cores_x, core_y0, core_rows;
  const uint32_t* cell; const uint8_t* state;
  uint32_t* mig_cnt; uint32_t* mig_a; uint16_t* mig_lc;
  const uint32_t* go;              /* optional: 0 = decisions are not final, do nothing */
  uint32_t n_own, ghost_cap;       /* strips: ghost agents sit at [n_own, n_own + cap) and [n_own + cap, n_own + 2 cap) ... */
  const uint32_t* ghost_total[2];  /* ... of which only the first *ghost_total[d] are real */
  ABM_HD void operator()(uint64_t i) const {
    if (go && !*go) return;
    if (i >= n_own) {
      const uint32_t k = (uint32_t)i - n_own, d = k >= ghost_cap ? 1u : 0u;
      if (k - d * ghost_cap >= *ghost_total[d]) return;
    }
    const uint8_t st = state[i];
    if (!(st & ST_MOVED)) return;
    const uint32_t c = cell[i];
    int dx, dy;
    off8(st & ST_DIRMASK, dx, dy);
    const int lx = (int)(c & 31) + dx, ly = (int)((c >> 5) & 31) + dy;
    if (lx >= 0 && lx < CORE && ly >= 0 && ly < CORE) return; /* stays inside its core */
    const uint32_t t = c >> 10;
    int ty = (int)(t / (uint32_t)g.tiles_x), tx = (int)(t - (uint32_t)ty * (uint32_t)g.tiles_x);
    tx += lx < 0 ? -1 : (lx >= CORE ? 1 : 0);
    ty += ly < 0 ? -1 : (ly >= CORE ? 1 : 0);
    if (g.periodic) tx = wrap(tx, g.tiles_x); /* the move itself was valid: off-grid targets never get ST_MOVED */
    if (g.py) ty = wrap(ty, g.tiles_y);
    const int cy = ty - core_y0;
    if (cy < 0 || cy >= core_rows || tx < 0 || tx >= cores_x) return; /* lands in a neighbouring strip's rows */
    const uint32_t core = (uint32_t)cy * (uint32_t)cores_x + (uint32_t)tx;
    const uint32_t k = atomic_add(&mig_cnt[core], 1u);
    if (k < MIG_CAP) { mig_a[(size_t)core * MIG_CAP + k] = (uint32_t)i; mig_lc[(size_t)core * MIG_CAP + k] = (uint16_t)(((ly & 31) << 5) | (lx & 31)); }
  }
};

struct CommitTile {
  View v;
  SubIndex si;
  int32_t cores_x, core_y0;
  uint32_t* out_cell; uint32_t* out_id; double* out_energy;
  uint32_t* new_off; uint32_t* new_cnt; /* next tick's sub-tile runs */
  uint32_t* cursor;                     /* next free slot of the output SoA */
  uint32_t* birth_pos; uint32_t* birth_parent; uint32_t* birth_count; uint32_t* birth_bits;
  const uint32_t* go; /* optional device flag: 0 = the host enqueued this commit optimistically and the tick is not ready for it */
  const uint32_t* mig_cnt; const uint32_t* mig_a; const uint16_t* mig_lc; /* optional migrant lists (MigrantScanBody) */

  static size_t smem_bytes() { return (size_t)(CORE * CORE + CORE * CORE / 4 + 16 * 3 + RUNTABLE_WORDS + 8 + 3 * ARRIVE_CAP + 16 * (COMMIT_THREADS / 32)) * 4; }

  ABM_HD void operator()(uint32_t cta, uint32_t* smem) const {
    if (go && !*go) return; /* uniform: nothing has been touched, the host takes the careful path */
    const int cx = (int)(cta % (uint32_t)cores_x), cy = (int)(cta / (uint32_t)cores_x) + core_y0;
    /* with a complete migrant list the window is the core itself (16 sub-tile runs); otherwise core + halo (36 runs) */
    const uint32_t n_mig = mig_cnt ? mig_cnt[cta] : NONE32;
    const bool use_mig = n_mig <= MIG_CAP;
    const int halo = use_mig ? 0 : HALO;
    Window w;
    w.set(cx * CORE - halo, cy * CORE - halo, CORE + 2 * halo, CORE + 2 * halo);
    const int nruns = w.runs(), NT = COMMIT_THREADS;
    uint32_t* min_all = smem;                         /* lowest id arriving on each core cell: the one that eats */
    uint8_t* grass = (uint8_t*)(smem + CORE * CORE);  /* the core's grass tile */
    uint32_t* counts = smem + CORE * CORE + CORE * CORE / 4; /* per sub-tile of the core: arrivals, base, cursor */
    uint32_t* base = counts + 16;
    uint32_t* cur = base + 16;
    RunTable rt;
    rt.bind(cur + 16);
    uint32_t* misc = cur + 16 + RUNTABLE_WORDS; /* [0] chunk, [1] arrivals listed, [2] list overflow */
    uint32_t* ar_a = misc + 8;                 /* arrivals: agent index, id, core cell | state << 16 */
    uint32_t* ar_id = ar_a + ARRIVE_CAP;
    uint32_t* ar_cs = ar_id + ARRIVE_CAP;
    uint32_t* wcounts = ar_cs + ARRIVE_CAP;    /* per warp x per sub-tile arrival counts (one hot address would serialise) */
    const uint32_t tile = (uint32_t)(cy * cores_x + cx); /* CORE == TILE: a core is one grass tile */
    uint32_t* gtile = (uint32_t*)(v.grass + (size_t)tile * TILE_CELLS);
    const uint32_t* my_mig_a = mig_a ? mig_a + (size_t)cta * MIG_CAP : nullptr;
    const uint16_t* my_mig_lc = mig_lc ? mig_lc + (size_t)cta * MIG_CAP : nullptr;

    ABM_CTA_FOR(t, NT) {
      rt.load(t, NT, nruns, w, v.g, si);
      for (int c = t; c < CORE * CORE; c += NT) min_all[c] = NONE32;
      for (int c = t; c < CORE * CORE / 4; c += NT) ((uint32_t*)grass)[c] = gtile[c];
      if (t < 16 * (COMMIT_THREADS / 32)) wcounts[t] = 0;
      if (t < 8) misc[t] = 0;
    }
    ABM_CTA_SYNC();
    ABM_CTA_FOR(t, NT) { rt.scan(t, NT, nruns); }
    ABM_CTA_SYNC();
    ABM_CTA_FOR(t, NT) { rt.expand(t, NT, nruns); }
    ABM_CTA_SYNC();
    const uint32_t M = rt.prefix[nruns];
    ABM_CTA_FOR(t, NT) { /* pass 1: who lands in the core */
      for (uint32_t m = (uint32_t)t; m < M + (use_mig ? n_mig : 0u); m += (uint32_t)NT) {
        uint32_t a;
        int x, y;
        uint8_t st;
        if (m < M) {
          const int r = rt.find(m, nruns);
          a = rt.off[r] + (m - rt.prefix[r]);
          if (!rt.locate(w, r, v.cell[a], x, y)) continue;
          st = v.state[a];
          if (!(st & ST_DECIDED)) { atomic_or(v.err, ERR_INTERNAL); continue; }
          if (st & ST_MOVED) { int dx, dy; off8(st & ST_DIRMASK, dx, dy); x += dx; y += dy; }
          x -= halo; y -= halo;
          if (x < 0 || x >= CORE || y < 0 || y >= CORE) continue;
        } else { /* a migrant: its final cell inside this core was computed by the scan */
          a = my_mig_a[m - M];
          const uint32_t lc = my_mig_lc[m - M];
          x = (int)(lc & 31); y = (int)(lc >> 5);
          st = v.state[a];
        }
        const uint32_t id = v.okey[a]; /* order key: the first to arrive eats */
        smem_min(&min_all[y * CORE + x], id);
        const uint32_t add = ((st & ST_DEAD) ? 0u : 1u) + ((st & ST_BIRTH) ? 1u : 0u);
        if (add) smem_add(&wcounts[(t >> 5) * 16 + (y >> 3) * 4 + (x >> 3)], add);
        const uint32_t slot = smem_add(&misc[1], 1u);
        if (slot < ARRIVE_CAP) { ar_a[slot] = a; ar_id[slot] = id; ar_cs[slot] = (uint32_t)(y * CORE + x) | ((uint32_t)st << 16); }
        else misc[2] = 1;
      }
    }
    ABM_CTA_SYNC();
    ABM_CTA_FOR(t, NT) {
      if (t < 16) { uint32_t c = 0; for (int q = 0; q < COMMIT_THREADS / 32; ++q) c += wcounts[q * 16 + t]; counts[t] = c; }
    }
    ABM_CTA_SYNC();
    ABM_CTA_FOR(t, NT) {
      if (t < 16) { uint32_t b = 0; for (int q = 0; q < t; ++q) b += counts[q]; base[t] = b; }
      if (t == 16) { uint32_t total = 0; for (int q = 0; q < 16; ++q) total += counts[q]; misc[0] = total ? atomic_add(cursor, total) : 0u; }
    }
    ABM_CTA_SYNC();
    ABM_CTA_FOR(t, NT) {
      if (t < 16) {
        const uint32_t at = misc[0] + base[t];
        cur[t] = at;
        const int sub = (cy * 4 + (t >> 2)) * si.nsx + cx * 4 + (t & 3);
        new_off[sub] = at;
        new_cnt[sub] = counts[t];
      }
    }
    ABM_CTA_SYNC();
    const bool listed = misc[2] == 0;
    const uint32_t n_pass2 = listed ? misc[1] : M + (use_mig ? n_mig : 0u);
    ABM_CTA_FOR(t, NT) { /* pass 2: energy, eating, births, output (over the arrival list when it fits) */
      for (uint32_t m = (uint32_t)t; m < n_pass2; m += (uint32_t)NT) {
        uint32_t a, id; /* id: the order key here as well */
        int x, y;
        uint8_t st;
        if (listed) {
          a = ar_a[m]; id = ar_id[m];
          const uint32_t cs = ar_cs[m];
          x = (int)(cs & 31); y = (int)((cs >> 5) & 31); st = (uint8_t)(cs >> 16);
        } else if (m < M) {
          const int r = rt.find(m, nruns);
          a = rt.off[r] + (m - rt.prefix[r]);
          if (!rt.locate(w, r, v.cell[a], x, y)) continue;
          st = v.state[a];
          if (st & ST_MOVED) { int dx, dy; off8(st & ST_DIRMASK, dx, dy); x += dx; y += dy; }
          x -= halo; y -= halo;
          if (x < 0 || x >= CORE || y < 0 || y >= CORE) continue;
          id = v.okey[a];
        } else {
          a = my_mig_a[m - M];
          const uint32_t lc = my_mig_lc[m - M];
          x = (int)(lc & 31); y = (int)(lc >> 5);
          st = v.state[a]; id = v.okey[a];
        }
        const int lc = y * CORE + x;
        double e = v.energy[a] - v.p.movement_cost; /* walk_ocurred is always true */
        if (v.p.eat && min_all[lc] == id && (grass[lc] & GRASS_GROWN)) { /* killed agents still eat (Appendix D-1) */
          e += v.p.delta_energy;
          grass[lc] = (uint8_t)(grass[lc] & ~GRASS_GROWN);
        }
        if (st & ST_BIRTH) e /= 2;
        const uint32_t fc = (tile << 10) | (uint32_t)lc;
        const int q = (y >> 3) * 4 + (x >> 3);
        if (!(st & ST_DEAD)) {
          const uint32_t pos = smem_add(&cur[q], 1u);
          out_cell[pos] = fc; out_id[pos] = v.okey == v.id ? id : v.id[a]; out_energy[pos] = e;
        }
        if (st & ST_BIRTH) { /* offspring on the parent's cell with the halved energy; its id is patched in afterwards */
          const uint32_t pos = smem_add(&cur[q], 1u);
          out_cell[pos] = fc; out_id[pos] = 0u; out_energy[pos] = e;
          const uint32_t b = atomic_add(birth_count, 1u);
          birth_pos[b] = pos; birth_parent[b] = id;
          if (birth_bits) atomic_or(&birth_bits[id >> 5], 1u << (id & 31));
        }
      }
    }
    ABM_CTA_SYNC();
    ABM_CTA_FOR(t, NT) { /* custom_model_step! on the tile, then one coalesced write-back */
      for (int q = t; q < CORE * CORE / 4; q += NT) { /* one word = four cells per thread */
        uint32_t wb = 0;
        if (v.p.walkmap_regrowth) wb = v.walkbits[tile * 32 + (uint32_t)(q >> 3)] >> ((q & 7) * 4);
        gtile[q] = regrow4(((const uint32_t*)grass)[q], (uint32_t)v.p.regrowth_time, wb, v.p.walkmap_regrowth != 0);
      }
    }
  }
};

/* 1 when the tick is ready for an optimistic commit: nobody left undecided and the output arrays are large enough */
struct TickGuardBody {
  const uint32_t* pending; const uint32_t* birth_flags; uint32_t birth_cap; uint32_t* go;
  ABM_HD void operator()(uint64_t) const { *go = (*pending == 0 && *birth_flags <= birth_cap) ? 1u : 0u; }
};

/* the same for strips, from the all-reduced figures of StripStatsBody (identical on every rank, so every rank decides alike) */
struct StripGuardBody {
  const unsigned long long* stats; int32_t n_ranks; uint32_t birth_cap; uint32_t* go;
  ABM_HD void operator()(uint64_t) const {
    bool ok = stats[0] == 0;
    for (int r = 0; r < n_ranks; ++r) ok = ok && stats[1 + n_ranks + r] <= birth_cap;
    *go = ok ? 1u : 0u;
  }
};

/* offspring ids: nextid(model) in the order of the parents' ids */
struct BirthIdBody {
  const uint32_t* birth_pos; const uint32_t* birth_parent; const uint32_t* birth_bits; const uint32_t* birth_prefix;
  uint32_t next_id;
  uint32_t* out_id;
  const uint32_t* birth_count; /* optional: device-side number of births (the launch is then an upper bound) */
  ABM_HD void operator()(uint64_t i) const {
    if (birth_count && i >= *birth_count) return;
    const uint32_t pid = birth_parent[i], wd = pid >> 5;
    const uint32_t rank = birth_prefix[wd] + (uint32_t)popc(birth_bits[wd] & ((1u << (pid & 31)) - 1u));
    out_id[birth_pos[i]] = next_id + rank;
  }
};

/* abm_set_agents on the tile path: host records -> SoA grouped by sub-tile */
struct IngestTileBody {
  Geo g; int32_t nsx;
  const int64_t* hid; const int64_t* hx; const int64_t* hy; const double* he;
  const uint32_t* pid; const uint32_t* pxy; /* abm_set_agents_packed (else null) */
  int64_t next_id;
  uint32_t* cell; uint32_t* id; double* energy; uint32_t* slot; uint32_t* sub;
  uint32_t* sub_cnt; uint32_t* alive_bits; uint32_t* err;
  ABM_HD void operator()(uint64_t i) const {
    int64_t x, y, a;
    host_record(hid, hx, hy, pid, pxy, g, i, a, x, y);
    if (x < 1 || x > g.nx || y < 1 + g.own_lo || y > g.own_hi || a < 1 || a >= next_id) { atomic_or(err, ERR_INTERNAL); cell[i] = 0; id[i] = 0; energy[i] = 0; sub[i] = NONE32; return; }
    const uint32_t bit = 1u << ((uint32_t)a & 31);
    if (atomic_or(&alive_bits[(uint32_t)a >> 5], bit) & bit) atomic_or(err, ERR_INTERNAL);
    const uint32_t st = (uint32_t)(((y - 1) >> 3) * nsx + ((x - 1) >> 3));
    cell[i] = enc(g, (int)(x - 1), (int)(y - 1)); id[i] = (uint32_t)a; energy[i] = he[i];
    sub[i] = st; slot[i] = atomic_add(&sub_cnt[st], 1u);
  }
};
struct ScatterTileBody {
  const uint32_t* cell; const uint32_t* id; const double* energy; const uint32_t* slot; const uint32_t* sub;
  const uint32_t* sub_off;
  uint32_t* out_cell; uint32_t* out_id; double* out_energy;
  ABM_HD void operator()(uint64_t i) const {
    if (sub[i] == NONE32) return;
    const uint32_t pos = sub_off[sub[i]] + slot[i];
    out_cell[pos] = cell[i]; out_id[pos] = id[i]; out_energy[pos] = energy[i];
  }
};

/* ----------------------------------------------------------------------------------- row strips (multi-GPU) */

/* Boundary messages have a fixed capacity `cap` (agents), agreed by all ranks, so that one exchange per tick suffices:
 * [header: nsx + 1 counts (the last is unused) padded to 16 bytes][energy cap x 8][cell cap x 4][id cap x 4][state cap].
 * prefix = exclusive scan of the header counts (nsx + 1 entries, the last = total), computed on the device. */
static inline uint64_t strip_hdr_bytes(int nsx) { return (((uint64_t)nsx + 1) * 4 + 15) & ~(uint64_t)15; }
static inline uint64_t strip_msg_bytes(int nsx, uint64_t cap) { return strip_hdr_bytes(nsx) + cap * 17; }

struct GatherCountsBody { /* the counts of one sub-tile row -> message header */
  const uint32_t* sub_cnt; int32_t row, nsx; uint32_t* out;
  ABM_HD void operator()(uint64_t i) const { out[i] = i < (uint64_t)nsx ? sub_cnt[(uint32_t)row * (uint32_t)nsx + (uint32_t)i] : 0u; }
};

/* exclusive prefix of `nsx` counts by one CTA of NT threads, thread t owning the chunk [t * per, (t + 1) * per):
 * phase 1 (chunk sums -> part[t]), phase 2 (sums of 32 parts -> part2), then every thread adds what precedes it */
ABM_HD uint32_t chunk_prefix(const uint32_t* part, const uint32_t* part2, int t) {
  uint32_t p = 0;
  for (int q = 0; q < (t >> 5); ++q) p += part2[q];
  for (int q = t & ~31; q < t; ++q) p += part[q];
  return p;
}

/* E1 in four launches.  (1) StripHeaderCta, CTA d = direction (0: my first own sub-tile row, to the previous strip;
 * 1: my last, to the next): header counts into the message and their prefix (kept in pre[d] for this tick's decision
 * exchanges).  (2) StripPackCta, one warp per sub-tile: the agents, packed in order, cells normalised to tile row 0
 * (the receiver adds its ghost tile row); the CTA that finishes last publishes the message.  (3) StripRxHeaderCta, CTA
 * d: waits for the neighbour's message, prefix of its header, ghost sub-tile runs.  (4) StripUnpackCta: the agents behind
 * the own ones.  With the peer-memory transport msg[] of (1), (2) is the neighbour's mailbox and nothing else moves. */
struct StripHeaderCta {
  SubIndex si; int32_t row[2];
  uint8_t* msg[2]; uint32_t* pre[2]; /* nsx + 1 entries, the last = total */
  uint32_t* done[2];                 /* pack CTAs finished (reset here) */
  static size_t smem_bytes() { return (size_t)(TILE_THREADS + 32 + 4) * 4; }
  ABM_HD void operator()(uint32_t d, uint32_t* smem) const {
    const int NT = TILE_THREADS, nsx = si.nsx, per = (nsx + NT - 1) / NT;
    uint32_t* part = smem; uint32_t* part2 = smem + NT;
    uint32_t* hdr_counts = (uint32_t*)msg[d];
    ABM_CTA_FOR(t, NT) {
      uint32_t sum = 0;
      for (int i = t * per; i < (t + 1) * per && i < nsx; ++i) { const uint32_t c = si.cnt[(uint32_t)row[d] * (uint32_t)nsx + (uint32_t)i]; hdr_counts[i] = c; sum += c; }
      part[t] = sum;
      if (t == 0) { hdr_counts[nsx] = 0; *done[d] = 0; }
    }
    ABM_CTA_SYNC();
    ABM_CTA_FOR(t, NT) { if (t < NT / 32) { uint32_t q = 0; for (int k = 0; k < 32; ++k) q += part[32 * t + k]; part2[t] = q; } }
    ABM_CTA_SYNC();
    ABM_CTA_FOR(t, NT) {
      uint32_t o = chunk_prefix(part, part2, t);
      for (int i = t * per; i < (t + 1) * per && i < nsx; ++i) { pre[d][i] = o; o += si.cnt[(uint32_t)row[d] * (uint32_t)nsx + (uint32_t)i]; }
      if (t == NT - 1) pre[d][nsx] = o; /* total: the last chunk ends the row */
    }
  }
};
/* The proposals of the two boundary sub-tile rows ahead of everybody else's (grid = 2 * chunks, one warp per sub-tile): the
 * halo message carries them, so the exchange can start while propose works through the rest of the strip.  propose is a
 * pure function of (tick, id, cell, energy): the full pass writes the same bytes again. */
struct ProposeRowsCta {
  View v; SubIndex si; int32_t row[2], chunks;
  ABM_HD void operator()(uint32_t cta, uint32_t*) const {
    const int NT = TILE_THREADS;
    const int d = (int)(cta / (uint32_t)chunks), chunk = (int)(cta % (uint32_t)chunks);
    ABM_CTA_FOR(t, NT) {
      const int i = chunk * (NT / 32) + (t >> 5);
      if (i < si.nsx) {
        const uint32_t sub = (uint32_t)row[d] * (uint32_t)si.nsx + (uint32_t)i, off = si.off[sub], cnt = si.cnt[sub];
        for (uint32_t k = (uint32_t)(t & 31); k < cnt; k += 32) {
          const uint32_t a = off + k;
          int x, y;
          dec(v.g, v.cell[a], x, y);
          bool nowalk;
          const uint8_t st = propose_state(v, x, y, v.id[a], v.energy[a], nowalk);
          if (nowalk) atomic_or(v.err, ERR_NOWALKABLE);
          v.state[a] = st;
        }
      }
    }
  }
};
struct StripPackCta { /* grid = 2 * chunks, chunk = 8 sub-tiles (one per warp) */
  SubIndex si; int32_t row[2], tiles_x, chunks; uint32_t cap;
  const uint32_t* cell; const uint32_t* id; const double* energy; const uint8_t* state;
  uint8_t* msg[2]; const uint32_t* pre[2]; uint64_t hdr;
  uint32_t* err; uint32_t* done[2];
  uint32_t* publish[2]; uint32_t seq;
  ABM_HD void operator()(uint32_t cta, uint32_t*) const {
    const int NT = TILE_THREADS;
    const int d = (int)(cta / (uint32_t)chunks), chunk = (int)(cta % (uint32_t)chunks);
    double* p_energy = (double*)(msg[d] + hdr);
    uint32_t* p_cell = (uint32_t*)(msg[d] + hdr + (uint64_t)cap * 8);
    uint32_t* p_id = (uint32_t*)(msg[d] + hdr + (uint64_t)cap * 12);
    uint8_t* p_state = msg[d] + hdr + (uint64_t)cap * 16;
    ABM_CTA_FOR(t, NT) {
      const int i = chunk * (NT / 32) + (t >> 5);
      if (i < si.nsx) {
        const uint32_t sub = (uint32_t)row[d] * (uint32_t)si.nsx + (uint32_t)i, off = si.off[sub], cnt = si.cnt[sub], o = pre[d][i];
        if (o + cnt > cap) { if ((t & 31) == 0) atomic_or(err, ERR_STRIP_CAP); }
        else for (uint32_t k = (uint32_t)(t & 31); k < cnt; k += 32) {
          const uint32_t c = cell[off + k];
          const uint32_t tx = (c >> 10) % (uint32_t)tiles_x;
          p_cell[o + k] = (tx << 10) | (c & 1023u); p_id[o + k] = id[off + k]; p_energy[o + k] = energy[off + k]; p_state[o + k] = state[off + k];
        }
      }
    }
    ABM_CTA_SYNC();
    ABM_CTA_FOR(t, NT) {
      if (t == 0 && publish[d]) { /* the CTA that finishes last publishes the whole message */
        flag_fence();
        if (atomic_add(done[d], 1u) == (uint32_t)chunks - 1) { *done[d] = 0; flag_publish(publish[d], seq); }
      }
    }
  }
};
struct StripRxHeaderCta {
  int32_t row[2], nsx; uint32_t base[2], cap;
  const uint8_t* msg[2]; uint32_t* pre[2];
  uint32_t* sub_off; uint32_t* sub_cnt;
  uint32_t* err;
  const uint32_t* wait[2]; uint32_t seq;
  static size_t smem_bytes() { return (size_t)(TILE_THREADS + 32 + 4) * 4; }
  ABM_HD void operator()(uint32_t d, uint32_t* smem) const {
    const int NT = TILE_THREADS, per = (nsx + NT - 1) / NT;
    uint32_t* part = smem; uint32_t* part2 = smem + NT;
    const uint32_t* counts = (const uint32_t*)msg[d];
    ABM_CTA_FOR(t, NT) { if (t == 0) flag_wait(wait[d], seq, err, ERR_P2P_TIMEOUT); }
    ABM_CTA_SYNC();
    ABM_CTA_FOR(t, NT) {
      uint32_t sum = 0;
      for (int i = t * per; i < (t + 1) * per && i < nsx; ++i) sum += counts[i];
      part[t] = sum;
    }
    ABM_CTA_SYNC();
    ABM_CTA_FOR(t, NT) { if (t < NT / 32) { uint32_t q = 0; for (int k = 0; k < 32; ++k) q += part[32 * t + k]; part2[t] = q; } }
    ABM_CTA_SYNC();
    ABM_CTA_FOR(t, NT) {
      uint32_t o = chunk_prefix(part, part2, t);
      for (int i = t * per; i < (t + 1) * per && i < nsx; ++i) {
        const uint32_t sub = (uint32_t)row[d] * (uint32_t)nsx + (uint32_t)i, cnt = counts[i];
        pre[d][i] = o;
        if (o + cnt > cap) { atomic_or(err, ERR_STRIP_CAP); sub_off[sub] = base[d]; sub_cnt[sub] = 0; }
        else { sub_off[sub] = base[d] + o; sub_cnt[sub] = cnt; }
        o += cnt;
      }
      if (t == NT - 1) pre[d][nsx] = o;
    }
  }
};
struct StripUnpackCta { /* grid = 2 * chunks, one warp per ghost sub-tile */
  int32_t nsx, chunks; uint32_t tile_row_base[2]; /* (ghost tile row * tiles_x) << 10 */
  uint32_t base[2], cap;
  const uint8_t* msg[2]; const uint32_t* pre[2]; uint64_t hdr;
  uint32_t* cell; uint32_t* id; double* energy; uint8_t* state;
  ABM_HD void operator()(uint32_t cta, uint32_t*) const {
    const int NT = TILE_THREADS;
    const int d = (int)(cta / (uint32_t)chunks), chunk = (int)(cta % (uint32_t)chunks);
    const uint32_t* counts = (const uint32_t*)msg[d];
    const double* p_energy = (const double*)(msg[d] + hdr);
    const uint32_t* p_cell = (const uint32_t*)(msg[d] + hdr + (uint64_t)cap * 8);
    const uint32_t* p_id = (const uint32_t*)(msg[d] + hdr + (uint64_t)cap * 12);
    const uint8_t* p_state = msg[d] + hdr + (uint64_t)cap * 16;
    ABM_CTA_FOR(t, NT) {
      const int i = chunk * (NT / 32) + (t >> 5);
      if (i < nsx) {
        const uint32_t cnt = counts[i], o = pre[d][i];
        if (o + cnt <= cap) for (uint32_t k = (uint32_t)(t & 31); k < cnt; k += 32) {
          cell[base[d] + o + k] = p_cell[o + k] + tile_row_base[d]; id[base[d] + o + k] = p_id[o + k];
          energy[base[d] + o + k] = p_energy[o + k]; state[base[d] + o + k] = p_state[o + k];
        }
      }
    }
  }
};

/* E2, sender side, CTA d = direction: the decision bytes of one boundary row in the order of E1 (straight into the
 * neighbour's mailbox and published there when the peer-memory transport is on) */
struct PackStatesCta { /* grid = 2 * chunks, one warp per sub-tile; the CTA that finishes last publishes */
  SubIndex si; int32_t row[2], chunks; uint32_t cap;
  const uint8_t* state; const uint32_t* pre[2]; uint8_t* out[2];
  uint32_t* done[2]; uint32_t* publish[2]; uint32_t seq;
  ABM_HD void operator()(uint32_t cta, uint32_t*) const {
    const int NT = TILE_THREADS;
    const int d = (int)(cta / (uint32_t)chunks), chunk = (int)(cta % (uint32_t)chunks);
    if (!out[d]) return;
    ABM_CTA_FOR(t, NT) {
      const int i = chunk * (NT / 32) + (t >> 5);
      if (i < si.nsx) {
        const uint32_t sub = (uint32_t)row[d] * (uint32_t)si.nsx + (uint32_t)i, off = si.off[sub], cnt = si.cnt[sub], o = pre[d][i];
        if (o + cnt <= cap) /* else: flagged by the agent exchange */
          for (uint32_t k = (uint32_t)(t & 31); k < cnt; k += 32) out[d][o + k] = state[off + k];
      }
    }
    ABM_CTA_SYNC();
    ABM_CTA_FOR(t, NT) {
      if (t == 0 && publish[d]) {
        flag_fence();
        if (atomic_add(done[d], 1u) == (uint32_t)chunks - 1) { *done[d] = 0; flag_publish(publish[d], seq); }
      }
    }
  }
};
/* E2, receiver side, grid = 2 * MERGE_CTAS: decisions are unique and final — a ghost copy keeps what it already settled
 * locally, the owner's decision wins otherwise */
enum { MERGE_CTAS = 8 };
struct MergeStatesCta {
  const uint8_t* incoming[2]; uint8_t* state; uint32_t base[2]; const uint32_t* total[2];
  const uint32_t* wait[2]; uint32_t seq; uint32_t* err;
  ABM_HD void operator()(uint32_t cta, uint32_t*) const {
    const int NT = TILE_THREADS;
    const int d = (int)(cta / MERGE_CTAS), part = (int)(cta % MERGE_CTAS);
    if (!incoming[d]) return;
    ABM_CTA_FOR(t, NT) { if (t == 0) flag_wait(wait[d], seq, err, ERR_P2P_TIMEOUT); }
    ABM_CTA_SYNC();
    const uint32_t n = *total[d];
    ABM_CTA_FOR(t, NT) {
      for (uint32_t i = (uint32_t)(part * NT + t); i < n; i += (uint32_t)(NT * MERGE_CTAS)) {
        const uint8_t in = incoming[d][i];
        if ((in & ST_DECIDED) || !(state[base[d] + i] & ST_DECIDED)) state[base[d] + i] = in;
      }
    }
  }
};

/* per-tick figures every rank needs from every rank, summed by one all-reduce (one-hot slots):
 * [0] open agents (all ranks), [1 + r] open agents of rank r, [1 + P + r] reproducing agents, [1 + 2P + r] largest boundary row */
struct StripStatsBody {
  const uint32_t* pending; const uint32_t* birth_flags; const uint32_t* tot_up; const uint32_t* tot_down;
  int32_t n_ranks, rank; unsigned long long* out;
  ABM_HD void operator()(uint64_t) const {
    const int P = n_ranks;
    for (int q = 0; q < 1 + 3 * P; ++q) out[q] = 0;
    out[0] = *pending; out[1 + rank] = *pending; out[1 + P + rank] = *birth_flags;
    const uint32_t a = *tot_up, b = *tot_down;
    out[1 + 2 * P + rank] = a > b ? a : b;
  }
};
/* offspring ids across strips: every rank contributes [count, parent ids ...]; an offspring's id
 * is next_id + the number of reproducing parents (anywhere) with a lower id.  One thread per local birth. */
struct BirthPackBody {
  const uint32_t* birth_parent; const uint32_t* birth_count; uint32_t cap; uint32_t* msg;
  ABM_HD void operator()(uint64_t i) const {
    const uint32_t n = *birth_count;
    if (i == 0) msg[0] = n;
    if (i < n && i < cap) msg[1 + i] = birth_parent[i];
  }
};
/* 32 work items per local birth: item (i, lane) counts the lower parent ids among every 32nd gathered id */
struct BirthRankBody {
  const uint32_t* all; /* n_ranks x (1 + cap) */
  int32_t n_ranks; uint32_t cap;
  const uint32_t* birth_parent; const uint32_t* birth_count;
  uint32_t* rank_out; /* zeroed, one per local birth */
  ABM_HD void operator()(uint64_t w) const {
    const uint32_t i = (uint32_t)(w >> 5), lane = (uint32_t)(w & 31);
    if (i >= *birth_count) return;
    const uint32_t mine = birth_parent[i];
    uint32_t below = 0;
    for (int r = 0; r < n_ranks; ++r) {
      const uint32_t* p = all + (size_t)r * (1 + cap);
      const uint32_t n = p[0] < cap ? p[0] : cap;
      for (uint32_t k = lane; k < n; k += 32) below += p[1 + k] < mine ? 1u : 0u;
    }
    if (below) atomic_add(&rank_out[i], below);
  }
};
struct BirthAssignBody {
  const uint32_t* all; int32_t n_ranks; uint32_t cap; uint32_t next_id;
  const uint32_t* birth_pos; const uint32_t* birth_parent; const uint32_t* birth_count;
  const uint32_t* rank_in;                                  /* from BirthRankBody, or */
  const uint32_t* birth_bits; const uint32_t* birth_prefix; /* the bitmap of all gathered parents + its popcount prefix */
  uint32_t* out_id; uint32_t* total_out;
  ABM_HD void operator()(uint64_t i) const {
    if (i == 0) { uint32_t t = 0; for (int r = 0; r < n_ranks; ++r) t += all[(size_t)r * (1 + cap)]; *total_out = t; }
    if (i >= *birth_count) return;
    uint32_t rank;
    if (birth_bits) {
      const uint32_t pid = birth_parent[i], wd = pid >> 5;
      rank = birth_prefix[wd] + (uint32_t)popc(birth_bits[wd] & ((1u << (pid & 31)) - 1u));
    } else rank = rank_in[i];
    out_id[birth_pos[i]] = next_id + rank;
  }
};
/* many offspring (C5: ~10^6 per tick): instead of comparing every parent with every other, mark all gathered parents in
 * a bitmap over the id space; a parent's rank is the popcount below its bit (the single-GPU way, BirthIdBody) */
struct BirthMarkBody {
  const uint32_t* all; int32_t n_ranks; uint32_t cap; uint32_t* bits;
  ABM_HD void operator()(uint64_t i) const {
    const uint32_t r = (uint32_t)(i / cap), k = (uint32_t)(i % cap);
    if ((int)r >= n_ranks) return;
    const uint32_t* p = all + (size_t)r * (1 + cap);
    const uint32_t n = p[0] < cap ? p[0] : cap;
    if (k >= n) return;
    const uint32_t pid = p[1 + k];
    atomic_or(&bits[pid >> 5], 1u << (pid & 31));
  }
};

}  // namespace abm
#endif /* ABM_TILE_H_ */
