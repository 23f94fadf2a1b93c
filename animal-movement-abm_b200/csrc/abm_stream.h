/*
 * abm_stream.h — the stream ("S") stepping path: walk rules whose move never waits for another agent.
 *
 * RANDOM_WALKMAP (src/agent_actions.jl:182-186: move_agent! to a random walkable neighbour, unconditional) and
 * RANDOM_WALK without `ifempty` need no conflict resolution: an agent's step of tick t + 1 depends only on where it
 * stands after tick t, its id and its energy.  What couples agents is eating (the lowest id standing on a grown cell
 * eats it, src/agent_actions.jl:139-147) and the offspring ids (nextid in the order of the parents' ids, :174) — both are
 * settled by the core an agent ENDS its step on.  So the state between two ticks is kept as "arrival lists":
 *
 *     list of core D, tick t  =  one 16-byte record per agent that ends tick t on a cell of D:
 *                                 (destination cell inside D, direction, DEAD / BIRTH flags | id | energy before the step)
 *
 * and ONE kernel per tick does everything: the CTA of core D streams its list once (coalesced 16-byte records), settles
 * eating against the core's grass tile in shared memory, applies energy, death and birth, and — since the survivor's
 * next step depends on nothing else — draws that step on the spot (Philox keyed by (seed, t + 1, id)) and appends the
 * record to the tick-t+1 list of the core it will end on (a warp-aggregated atomic per destination core, one 16-byte
 * store per agent).  The grass tile is regrown and written back by the same CTA (src/model_actions.jl:18-30).
 * Per agent-update that is one 16-byte read and one 16-byte write; per cell one byte read and one written.
 *
 * Offspring ids: the parents of tick t + 1 are known when their records are filed, i.e. inside the kernel of tick t;
 * it marks them in a bitmap over the id space, a popcount prefix over that bitmap runs between the two kernels, and the
 * kernel of tick t + 1 gives the child of parent p the id  next_id + #(parents with a lower id)  without any fix-up pass.
 *
 * Row strips: a record filed at a core of the ghost tile rows belongs to the neighbouring strip; the ghost-row lists are
 * shipped to it at the start of the next tick (the only exchange of the tick, together with the parents' ids).
 *
 * The canonical SoA (cell, id, energy) that the getters of the C ABI read is rebuilt from the lists on demand
 * (StreamExportCta): a record's origin cell is its destination minus its direction.
 */
#ifndef ABM_STREAM_H_
#define ABM_STREAM_H_

#include "abm_tile.h"

namespace abm {

#ifndef ABM_S_THREADS
#define ABM_S_THREADS 96
#endif
#ifndef ABM_S_MIN_CTAS
#define ABM_S_MIN_CTAS 21
#endif
enum { S_THREADS = ABM_S_THREADS, S_MIN_CTAS = ABM_S_MIN_CTAS }; /* CTA size of the tick kernel, CTAs per SM its registers must allow */
/* record word 0: destination cell inside its core (ly << 5 | lx) | direction << 10 (0..7 moved, DIR_STAY) | flags */
enum { SR_LC = 0x3FF, SR_DIR_SHIFT = 10, SR_DEAD = 1u << 14, SR_BIRTH = 1u << 15, SR_NOWALK = 1u << 16 };

struct alignas(16) SRec { uint32_t w0, id, e_lo, e_hi; };

ABM_HD SRec load_rec(const SRec* p) {
#if defined(__CUDA_ARCH__)
  const uint4 q = *reinterpret_cast<const uint4*>(p);
  SRec r; r.w0 = q.x; r.id = q.y; r.e_lo = q.z; r.e_hi = q.w;
  return r;
#else
  return *p;
#endif
}
ABM_HD void store_rec(SRec* p, const SRec& r) {
#if defined(__CUDA_ARCH__)
  *reinterpret_cast<uint4*>(p) = make_uint4(r.w0, r.id, r.e_lo, r.e_hi);
#else
  *p = r;
#endif
}
ABM_HD double rec_energy(const SRec& r) {
  const uint64_t b = (uint64_t)r.e_lo | ((uint64_t)r.e_hi << 32);
#if defined(__CUDA_ARCH__)
  return __longlong_as_double((long long)b);
#else
  double d; memcpy(&d, &b, 8); return d;
#endif
}
ABM_HD void set_rec_energy(SRec& r, double e) {
  uint64_t b;
#if defined(__CUDA_ARCH__)
  b = (uint64_t)__double_as_longlong(e);
#else
  memcpy(&b, &e, 8);
#endif
  r.e_lo = (uint32_t)b; r.e_hi = (uint32_t)(b >> 32);
}

/* the arrival lists of one tick: `cap` record slots per tile of the stored grid (ghost tile rows included), the number of
 * records filed per tile (may exceed cap: the excess sits in the overflow list, tagged with its tile) */
struct StreamLists {
  SRec* rec; uint32_t* cnt; uint32_t cap;
  SRec* ovf_rec; uint32_t* ovf_tile; uint32_t* ovf_cnt; uint32_t ovf_cap;
};

/* where the parents of the next tick are recorded: a bitmap over the id space (one GPU) or a list that the strips
 * all-gather (then every rank marks the bitmap from the gathered list) */
struct ParentSink { uint32_t* bits; uint32_t* list; uint32_t* list_count; uint32_t list_cap; SchedKey sk; /* order keys of that tick */ };

/* Agent (id, energy e) stands on (x, y) of the stored grid when tick v.p.tick begins: draw its step and file the record at
 * the core it ends on.  Every lane of the warp calls this (valid = this lane has an agent): the slots of one core are
 * taken with one atomic per warp. */
template <bool SCHED>
ABM_HD void stream_emit(const View& v, const StreamLists& out, const ParentSink& ps, int x, int y, uint32_t id, double e, bool valid) {
  SRec r;
  r.w0 = 0; r.id = id; r.e_lo = 0; r.e_hi = 0;
  uint32_t tile = NONE32;
  if (valid) {
    bool nowalk;
    const uint8_t st = propose_state(v, x, y, id, e, nowalk);
    int fx = x, fy = y;
    if (st & ST_MOVED) {
      int dx, dy;
      off8(st & ST_DIRMASK, dx, dy);
      if (x >= 1 && x + 1 < v.g.nx && y >= 1 && y + 1 < v.g.ny) { fx = x + dx; fy = y + dy; } /* away from the rim: no wrap */
      else shift(v.g, x, y, dx, dy, fx, fy);
    }
    tile = (uint32_t)((fy >> TILE_SHIFT) * v.g.tiles_x + (fx >> TILE_SHIFT));
    r.w0 = (uint32_t)(((fy & 31) << 5) | (fx & 31)) | ((uint32_t)((st & ST_MOVED) ? (st & ST_DIRMASK) : (int)DIR_STAY) << SR_DIR_SHIFT) |
           ((st & ST_DEAD) ? (uint32_t)SR_DEAD : 0u) | ((st & ST_BIRTH) ? (uint32_t)SR_BIRTH : 0u) | (nowalk ? (uint32_t)SR_NOWALK : 0u);
    set_rec_energy(r, e);
    if (st & ST_BIRTH) { /* parents are recorded by order key: nextid goes to the offspring in the order the parents are stepped */
      const uint32_t pk = SCHED ? order_key(ps.sk, id) : id;
      if (ps.list) { const uint32_t k = atomic_add(ps.list_count, 1u); if (k < ps.list_cap) ps.list[k] = pk; }
      else atomic_or(&ps.bits[pk >> 5], 1u << (pk & 31));
    }
  }
  const uint32_t slot = agg_add(out.cnt, tile, valid);
  if (!valid) return;
  if (slot < out.cap) { store_rec(&out.rec[(size_t)tile * out.cap + slot], r); return; }
  const uint32_t o = atomic_add(out.ovf_cnt, 1u);
  if (o < out.ovf_cap) { store_rec(&out.ovf_rec[o], r); out.ovf_tile[o] = tile; }
  else atomic_or(v.err, ERR_LIST_CAP);
}

/* canonical SoA -> the lists of tick v.p.tick (after abm_set_agents, or when the host changed the tick or the terrain).
 * Launched over n rounded up to a multiple of 32 so that whole warps call stream_emit. */
struct StreamIngestBody {
  View v; StreamLists out; ParentSink ps; uint32_t n;
  ABM_HD void operator()(uint64_t i) const {
    const bool valid = i < n;
    int x = 0, y = 0;
    uint32_t id = 0; double e = 0;
    if (valid) { dec(v.g, v.cell[i], x, y); id = v.id[i]; e = v.energy[i]; }
    stream_emit<true>(v, out, ps, x, y, id, e, valid); /* not a hot path: the scheduler is looked up at run time */
  }
};

/* One tick: CTA b = core b of the own rows.  custom_agent_step! after the walk (reduce_energy!, eat!, reproduce!,
 * src/agent_actions.jl:54-67) for every agent that ends tick v.p.tick on this core, the walk of tick + 1 for what remains,
 * custom_model_step! (src/model_actions.jl:18-30) on the core's grass tile. */
/* SCHED = false: Schedulers.by_id, order key = id (the instantiation the benchmarks run: no key arithmetic at all) */
template <bool SCHED>
struct StreamTickCta {
  View v;                /* v.p.tick = the tick being completed */
  StreamLists in, out;
  ParentSink ps;         /* parents of tick + 1 */
  int32_t cores_x, core_y0, row_step; /* CTA row q works on core row core_y0 + q * row_step */
  uint32_t next_id;      /* nextid(model) when the tick began */
  const uint32_t* birth_bits; const uint32_t* birth_prefix; /* parents of this tick, every strip's, + popcount prefix per 8-word block */
  uint32_t* emitted;     /* records filed for tick + 1 */
  const uint32_t* births_total; /* Schedulers.randomly: offspring of this tick (device side), for the key width of tick + 1 */

  enum { BIRTH_LIST = 128 }; /* parents listed per core and tick; a core with more re-scans its records for them */
  static size_t smem_bytes() { return (size_t)(CORE * CORE + CORE * CORE / 4 + 8 + BIRTH_LIST) * 4; }

  /* the offspring of record r (whose energy field already holds the halved energy): on the parent's cell, id = nextid in
   * the order of the parents' ids (:172-178, Appendix C D6); its first step is drawn like everybody's */
  ABM_HD uint32_t key_of(uint32_t id) const { return SCHED ? order_key(v.p.sk, id) : id; }
  ABM_HD uint32_t newborn(const View& vn, const ParentSink& psn, const SRec& r, bool valid, int x0, int y0) const {
    const uint32_t lc = r.w0 & SR_LC;
    uint32_t child = 0;
    if (valid) {
      child = next_id + rank_in_blocks(birth_bits, birth_prefix, key_of(r.id));
    }
    stream_emit<SCHED>(vn, out, psn, x0 + (int)(lc & 31), y0 + (int)(lc >> 5), child, rec_energy(r), valid);
    return valid ? 1u : 0u;
  }

  ABM_HD void operator()(uint32_t cta, uint32_t* smem) const {
    const int NT = S_THREADS;
    const int cx = (int)(cta % (uint32_t)cores_x), cy = (int)(cta / (uint32_t)cores_x) * row_step + core_y0;
    const uint32_t tile = (uint32_t)(cy * v.g.tiles_x + cx);
    uint32_t* min_all = smem;                        /* lowest id arriving on each cell: the one that eats */
    uint8_t* grass = (uint8_t*)(smem + CORE * CORE); /* the core's grass tile */
    uint32_t* misc = smem + CORE * CORE + CORE * CORE / 4; /* [0] records filed, [1] parents, [2] parents in the overflow list */
    uint32_t* blist = misc + 8;                      /* record index of each parent */
    uint32_t* gtile = (uint32_t*)(v.grass + (size_t)tile * TILE_CELLS);
    const uint32_t filed = in.cnt[tile];
    const uint32_t n = filed < in.cap ? filed : in.cap;
    SRec* mine = in.rec + (size_t)tile * in.cap;
    const uint32_t n_ovf = filed > in.cap ? (*in.ovf_cnt < in.ovf_cap ? *in.ovf_cnt : in.ovf_cap) : 0u;
    const int x0 = cx * CORE, y0 = cy * CORE;
    View vn = v;
    vn.p.tick = v.p.tick + 1;
    ParentSink psn = ps; /* the parents of tick + 1 under that tick's order keys: its key width follows from nextid after this tick */
    if (SCHED) psn.sk.hb = sched_half_bits((uint64_t)next_id + (uint64_t)*births_total);

    ABM_CTA_FOR(t, NT) {
      fill_words(min_all, NONE32, CORE * CORE, t, NT);
      for (int c = t; c < CORE * CORE / 4; c += NT) ((uint32_t*)grass)[c] = gtile[c];
      if (t < 8) misc[t] = 0;
    }
    ABM_CTA_SYNC();
    ABM_CTA_FOR(t, NT) { /* pass 1: who eats where */
      for (uint32_t i = (uint32_t)t; i < n; i += (uint32_t)NT) {
        const SRec r = load_rec(&mine[i]);
        smem_min(&min_all[r.w0 & SR_LC], key_of(r.id)); /* first come (lowest order key), first served */
      }
      for (uint32_t o = (uint32_t)t; o < n_ovf; o += (uint32_t)NT)
        if (in.ovf_tile[o] == tile) { const SRec r = load_rec(&in.ovf_rec[o]); smem_min(&min_all[r.w0 & SR_LC], key_of(r.id)); }
    }
    ABM_CTA_SYNC();
    ABM_CTA_FOR(t, NT) { /* pass 2: the rest of the step, and the next one (whole warps: stream_emit aggregates) */
      uint32_t made = 0;
      for (uint32_t base = 0; base < n + n_ovf; base += (uint32_t)NT) {
        const uint32_t i = base + (uint32_t)t;
        SRec* at = i < n ? &mine[i] : &in.ovf_rec[i - n];
        const bool valid = i < n || (i < n + n_ovf && in.ovf_tile[i - n] == tile);
        SRec r;
        r.w0 = 0; r.id = 0; r.e_lo = 0; r.e_hi = 0;
        if (valid) r = load_rec(at);
        const uint32_t lc = r.w0 & SR_LC;
        double e = rec_energy(r) - v.p.movement_cost; /* walk_ocurred is always true */
        if (valid) {
          if (v.p.eat && min_all[lc] == key_of(r.id) && (grass[lc] & GRASS_GROWN)) { /* killed agents still eat (Appendix D-1) */
            e += v.p.delta_energy;
            grass[lc] = (uint8_t)(grass[lc] & ~GRASS_GROWN);
          }
          if (r.w0 & SR_NOWALK) atomic_or(v.err, ERR_NOWALKABLE);
          if (r.w0 & SR_BIRTH) { /* reproduce!: the halved energy goes back into the record, the offspring is filed in pass 3 */
            e /= 2;
            set_rec_energy(r, e);
            store_rec(at, r);
            if (i < n) { const uint32_t k = smem_add(&misc[1], 1u); if (k < BIRTH_LIST) blist[k] = i; }
            else misc[2] = 1;
          }
        }
        const bool alive = valid && !(r.w0 & SR_DEAD);
        stream_emit<SCHED>(vn, out, psn, x0 + (int)(lc & 31), y0 + (int)(lc >> 5), r.id, e, alive);
        made += alive ? 1u : 0u;
      }
      if (made) smem_add(&misc[0], made);
    }
    ABM_CTA_SYNC();
    const uint32_t nb = misc[1];
    const bool rescan = nb > BIRTH_LIST || misc[2] != 0;
    ABM_CTA_FOR(t, NT) { /* pass 3: the offspring (a handful per core), then custom_model_step! on the tile */
      uint32_t made = 0;
      if (!rescan) {
        for (uint32_t base = 0; base < nb; base += (uint32_t)NT) {
          const uint32_t k = base + (uint32_t)t;
          const bool valid = k < nb;
          SRec r;
          r.w0 = 0; r.id = 0; r.e_lo = 0; r.e_hi = 0;
          if (valid) r = load_rec(&mine[blist[k]]);
          made += newborn(vn, psn, r, valid, x0, y0);
        }
      } else {
        for (uint32_t base = 0; base < n + n_ovf; base += (uint32_t)NT) {
          const uint32_t i = base + (uint32_t)t;
          const bool mine_too = i < n || (i < n + n_ovf && in.ovf_tile[i - n] == tile);
          SRec r;
          r.w0 = 0; r.id = 0; r.e_lo = 0; r.e_hi = 0;
          if (mine_too) r = load_rec(i < n ? &mine[i] : &in.ovf_rec[i - n]);
          made += newborn(vn, psn, r, mine_too && (r.w0 & SR_BIRTH), x0, y0);
        }
      }
      if (made) smem_add(&misc[0], made);
      for (int q = t; q < CORE * CORE / 4; q += NT) { /* one word = four cells per thread, one coalesced write-back */
        uint32_t wb = 0;
        if (v.p.walkmap_regrowth) wb = v.walkbits[tile * 32 + (uint32_t)(q >> 3)] >> ((q & 7) * 4);
        gtile[q] = regrow4(((const uint32_t*)grass)[q], (uint32_t)v.p.regrowth_time, wb, v.p.walkmap_regrowth != 0);
      }
    }
    ABM_CTA_SYNC();
    ABM_CTA_FOR(t, NT) { if (t == 0 && misc[0]) atomic_add(emitted, misc[0]); }
  }
};

/* lists -> canonical SoA (any order): CTA b = tile b of the stored grid (ghost tile rows included: their records are own
 * agents about to cross into the neighbouring strip).  Origin cell = destination - direction. */
struct StreamExportCta {
  Geo g; StreamLists in;
  uint32_t* cell; uint32_t* id; double* energy; uint32_t* cursor;
  static size_t smem_bytes() { return 16; }
  ABM_HD void put(uint32_t pos, uint32_t tile, const SRec& r) const {
    const uint32_t ty = tile / (uint32_t)g.tiles_x, tx = tile - ty * (uint32_t)g.tiles_x;
    int x = (int)(tx << TILE_SHIFT) | (int)(r.w0 & 31), y = (int)(ty << TILE_SHIFT) | (int)((r.w0 >> 5) & 31);
    const int dir = (int)((r.w0 >> SR_DIR_SHIFT) & 15);
    if (dir < 8) { int dx, dy, ox, oy; off8(dir, dx, dy); shift(g, x, y, -dx, -dy, ox, oy); x = ox; y = oy; }
    cell[pos] = enc(g, x, y); id[pos] = r.id; energy[pos] = rec_energy(r);
  }
  ABM_HD void operator()(uint32_t tile, uint32_t* smem) const {
    const int NT = S_THREADS;
    const uint32_t filed = in.cnt[tile];
    const uint32_t n = filed < in.cap ? filed : in.cap;
    if (n == 0) return;
    ABM_CTA_FOR(t, NT) { if (t == 0) smem[0] = atomic_add(cursor, n); }
    ABM_CTA_SYNC();
    ABM_CTA_FOR(t, NT) {
      for (uint32_t i = (uint32_t)t; i < n; i += (uint32_t)NT) put(smem[0] + i, tile, load_rec(&in.rec[(size_t)tile * in.cap + i]));
    }
  }
};
struct StreamExportOvfBody {
  StreamExportCta ex; uint32_t base;
  ABM_HD void operator()(uint64_t o) const { ex.put(base + (uint32_t)o, ex.in.ovf_tile[o], load_rec(&ex.in.ovf_rec[o])); }
};

/* ---- row strips: the lists of the two ghost tile rows travel to the neighbours at the start of a tick ------------- */

ABM_HOSTDEV uint64_t stream_hdr_bytes(int tiles_x) { return (((uint64_t)tiles_x) * 4 + 15) & ~(uint64_t)15; }
ABM_HOSTDEV uint64_t stream_msg_bytes(int tiles_x, uint64_t cap) { return stream_hdr_bytes(tiles_x) + cap * 16; }

/* message = [count per core x tiles_x][records, core after core (prefix of the counts)], capacity `cap` records.
 * One CTA per direction d (0: my upper ghost row -> previous strip, 1: lower -> next). */
struct StreamPackCta {
  StreamLists in; int32_t tiles_x, ghost_row[2]; uint32_t cap;
  uint8_t* msg[2]; uint32_t* err;
  uint32_t* publish[2]; uint32_t seq;
  static size_t smem_bytes() { return (size_t)(TILE_THREADS + 40) * 4; }
  ABM_HD void operator()(uint32_t d, uint32_t* smem) const {
    const int NT = TILE_THREADS, per = (tiles_x + NT - 1) / NT;
    if (!msg[d]) return;
    uint32_t* part = smem; uint32_t* part2 = smem + NT;
    uint32_t* counts = (uint32_t*)msg[d];
    SRec* recs = (SRec*)(msg[d] + stream_hdr_bytes(tiles_x));
    const uint32_t tile0 = (uint32_t)(ghost_row[d] * tiles_x);
    ABM_CTA_FOR(t, NT) {
      uint32_t sum = 0;
      for (int i = t * per; i < (t + 1) * per && i < tiles_x; ++i) {
        const uint32_t f = in.cnt[tile0 + (uint32_t)i];
        if (f > in.cap) atomic_or(err, ERR_STRIP_CAP); /* a ghost core overflowed its slots: the row cannot be shipped */
        const uint32_t c = f < in.cap ? f : in.cap;
        counts[i] = c; sum += c;
      }
      part[t] = sum;
    }
    ABM_CTA_SYNC();
    ABM_CTA_FOR(t, NT) { if (t < NT / 32) { uint32_t q = 0; for (int k = 0; k < 32; ++k) q += part[32 * t + k]; part2[t] = q; } }
    ABM_CTA_SYNC();
    ABM_CTA_FOR(t, NT) {
      uint32_t o = chunk_prefix(part, part2, t);
      for (int i = t * per; i < (t + 1) * per && i < tiles_x; ++i) {
        const uint32_t c = counts[i];
        if (o + c > cap) { atomic_or(err, ERR_STRIP_CAP); counts[i] = 0; }
        else for (uint32_t k = 0; k < c; ++k) store_rec(&recs[o + k], load_rec(&in.rec[(size_t)(tile0 + (uint32_t)i) * in.cap + k]));
        o += c;
      }
    }
    ABM_CTA_SYNC();
    ABM_CTA_FOR(t, NT) { if (t == 0 && publish[d]) flag_publish(publish[d], seq); }
  }
};
/* receiver: the records of message d join the lists of my first (d = 0) / last (d = 1) own core row */
struct StreamUnpackCta {
  StreamLists out; int32_t tiles_x, own_row[2]; uint32_t cap;
  const uint8_t* msg[2]; uint32_t* err;
  const uint32_t* wait[2]; uint32_t seq;
  static size_t smem_bytes() { return (size_t)(TILE_THREADS + 40) * 4; }
  ABM_HD void operator()(uint32_t d, uint32_t* smem) const {
    const int NT = TILE_THREADS, per = (tiles_x + NT - 1) / NT;
    if (!msg[d]) return;
    uint32_t* part = smem; uint32_t* part2 = smem + NT;
    ABM_CTA_FOR(t, NT) { if (t == 0) flag_wait(wait[d], seq, err, ERR_P2P_TIMEOUT); }
    ABM_CTA_SYNC();
    const uint32_t* counts = (const uint32_t*)msg[d];
    const SRec* recs = (const SRec*)(msg[d] + stream_hdr_bytes(tiles_x));
    const uint32_t tile0 = (uint32_t)(own_row[d] * tiles_x);
    ABM_CTA_FOR(t, NT) {
      uint32_t sum = 0;
      for (int i = t * per; i < (t + 1) * per && i < tiles_x; ++i) sum += counts[i];
      part[t] = sum;
    }
    ABM_CTA_SYNC();
    ABM_CTA_FOR(t, NT) { if (t < NT / 32) { uint32_t q = 0; for (int k = 0; k < 32; ++k) q += part[32 * t + k]; part2[t] = q; } }
    ABM_CTA_SYNC();
    ABM_CTA_FOR(t, NT) {
      uint32_t o = chunk_prefix(part, part2, t);
      for (int i = t * per; i < (t + 1) * per && i < tiles_x; ++i) {
        const uint32_t c = counts[i], tile = tile0 + (uint32_t)i;
        if (c && o + c <= cap) {
          const uint32_t at = atomic_add(&out.cnt[tile], c);
          for (uint32_t k = 0; k < c; ++k) {
            const SRec r = load_rec(&recs[o + k]);
            if (at + k < out.cap) store_rec(&out.rec[(size_t)tile * out.cap + at + k], r);
            else { const uint32_t q = atomic_add(out.ovf_cnt, 1u); if (q < out.ovf_cap) { store_rec(&out.ovf_rec[q], r); out.ovf_tile[q] = tile; } else atomic_or(err, ERR_LIST_CAP); }
          }
        }
        o += c;
      }
    }
  }
};
/* after the shipment: the ghost-row lists are empty again (their agents now live in the neighbour's lists) */
struct StreamClearGhostBody {
  uint32_t* cnt; int32_t tiles_x, ghost_row[2];
  ABM_HD void operator()(uint64_t i) const { const int d = (int)(i / (uint64_t)tiles_x), k = (int)(i % (uint64_t)tiles_x); cnt[ghost_row[d] * tiles_x + k] = 0; }
};
/* the gathered parents of every strip -> bitmap over the id space */
struct StreamMarkParentsBody {
  const uint32_t* all; int32_t n_ranks; uint32_t cap; uint32_t* bits; uint32_t* counts_out;
  ABM_HD void operator()(uint64_t i) const {
    const uint32_t cc = cap ? cap : 1u;
    const uint32_t r = (uint32_t)(i / cc), k = (uint32_t)(i % cc);
    if ((int)r >= n_ranks) return;
    const uint32_t* p = all + (size_t)r * (1 + cap);
    const uint32_t n = p[0] < cap ? p[0] : cap;
    if (k == 0) counts_out[r] = p[0];
    if (k < n) { const uint32_t pid = p[1 + k]; atomic_or(&bits[pid >> 5], 1u << (pid & 31)); }
  }
};

}  // namespace abm
#endif /* ABM_STREAM_H_ */
