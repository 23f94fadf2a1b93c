/*
 * abm_host_stream.inl — host side of the stream path (abm_stream.h): list allocation, canonical SoA <-> arrival lists,
 * the one-kernel tick, and (row strips) the shipment of the ghost-row lists and of the parents' ids.
 * Part of abm_engine.cu (one translation unit: included there, inside its anonymous namespace); not a stand-alone header.
 */
enum { SC_EMITTED = 0, SC_OVF0 = 1, SC_OVF1 = 2, SC_PARENTS = 3, SC_CURSOR = 4, SC_ERR = 8, SC_ZERO = 12 /* never written */, SC_GCOUNT = 16 /* parents per rank, 16 ranks */, SC_WORDS = 32, SC_HOST_BIRTHS = 40 };
enum { STREAM_MAX_RANKS = 16 };

/* walk rules whose move never waits for another agent */
static bool stream_rule(const abm_config& c) {
  return c.walk_type == ABM_RANDOM_WALKMAP || (c.walk_type == ABM_RANDOM_WALK && !c.randomwalk_ifempty);
}

static StreamLists stream_lists(abm_handle* h, int which) {
  StreamLists l;
  l.rec = (SRec*)h->s_rec[which].p; l.cnt = (uint32_t*)h->s_cnt[which].p; l.cap = h->s_cap;
  l.ovf_rec = (SRec*)h->s_ovf_rec[which].p; l.ovf_tile = (uint32_t*)h->s_ovf_tile[which].p;
  l.ovf_cnt = (uint32_t*)h->s_ctr.p + (which ? SC_OVF1 : SC_OVF0); l.ovf_cap = (uint32_t)std::min<uint64_t>(h->s_ovf_cap, 0xFFFFFFFFull);
  return l;
}

static ParentSink stream_sink(abm_handle* h, int which) {
  ParentSink ps;
  ps.bits = (uint32_t*)h->s_bits[which].p; ps.list = nullptr; ps.list_count = nullptr; ps.list_cap = 0;
  if (h->strip && h->n_ranks > 1) { /* strips rank the parents of all ranks: a list to all-gather */
    ps.list = (uint32_t*)h->s_parents.p + 1; ps.list_count = (uint32_t*)h->s_ctr.p + SC_PARENTS; ps.list_cap = (uint32_t)h->s_parent_cap;
  }
  /* the parents are recorded under the order keys of the tick whose steps are being filed: h->tick for the lists of the
   * current tick (ingest), h->tick + 1 for those the tick kernel files (it derives that tick's key width itself) */
  ps.sk = sched_key(h, which == h->s_cur ? h->tick : h->tick + 1, h->next_id);
  return ps;
}

static int stream_read_ctr(abm_handle* h) {
  ABM_CK(h, copy_d2h(h->h_counters, h->s_ctr.p, SC_WORDS * sizeof(uint32_t), h->stream));
  return 0;
}

static int stream_error_to_code(abm_handle* h) {
  const uint32_t e = h->h_counters[SC_ERR];
  if (!e) return 0;
  dev_memset((uint32_t*)h->s_ctr.p + SC_ERR, 0, 4, h->stream);
  if (e & ERR_NOWALKABLE) ABM_FAIL(h, ABM_E_NOWALKABLE, "random_walkmap: an agent has no walkable neighbour (the reference throws on rand(1:0))");
  if (e & ERR_P2P_TIMEOUT) { h->poisoned = true; ABM_FAIL(h, ABM_E_CUDA, "peer exchange timed out: a neighbouring strip did not deliver its message (the kernels went on without it: this handle is unusable now)"); }
  if (e & ERR_STRIP_CAP) ABM_FAIL(h, ABM_E_CAPACITY, "a ghost-row list outgrew the agreed message capacity (%llu records) within one tick", (unsigned long long)h->s_msg_cap);
  if (e & ERR_LIST_CAP) ABM_FAIL(h, ABM_E_CAPACITY, "arrival lists and their overflow list are full (%u slots per core, %llu overflow slots)", h->s_cap, (unsigned long long)h->s_ovf_cap);
  ABM_FAIL(h, ABM_E_INTERNAL, "device-side error flags 0x%x on the stream path", e);
}

/* popcount prefix of a parents bitmap per block of 8 words; `words` is a multiple of 8 and the last block is a spare of
 * zeros, so its prefix is the total */
static int stream_prefix(abm_handle* h, const uint32_t* bits, uint64_t words) {
  int rc;
  const uint64_t blocks = words / 8;
  if ((rc = ensure(h, h->bits_prefix, (blocks + 4) * 4))) return rc;
  if ((rc = ensure(h, h->scan_tmp, scan_tmp_words(std::max<uint64_t>(blocks, (uint64_t)h->g.gt + 1)) * 4))) return rc;
  ProfScope ps(h->prof, h->stream, "birth_prefix", words, 2);
  PopcountBlockBody pb;
  pb.bits = bits; pb.out = (uint32_t*)h->bits_prefix.p;
  launch_foreach(h->stream, pb, blocks);
  ABM_CK(h, scan_exclusive_u32(h->stream, (uint32_t*)h->bits_prefix.p, blocks, (uint32_t*)h->scan_tmp.p));
  return 0;
}

/* words of a parents bitmap whose keys are below key_bound: whole 8-word blocks plus the spare block */
static uint64_t stream_bitmap_words(uint64_t key_bound) { return ((bitmap_words(key_bound) + 7) & ~(uint64_t)7) + 8; }
static int stream_agree_parent_cap(abm_handle* h, bool first);

/* canonical SoA -> arrival lists of tick h->tick */
static int stream_ingest(abm_handle* h) {
  int rc;
  const uint64_t n = h->n;
  const uint64_t tiles = (uint64_t)h->g.tiles_x * (uint64_t)h->g.tiles_y;
  const uint64_t cores_own = (uint64_t)(h->g.nx / CORE) * (uint64_t)((h->g.own_hi - h->g.own_lo) / CORE);
  if ((rc = ensure(h, h->s_ctr, SC_WORDS * 4))) return rc;
  uint32_t* sc = (uint32_t*)h->s_ctr.p;
  if (h->s_cap == 0 || h->s_relayout) {
    const uint64_t avg = n / std::max<uint64_t>(cores_own, 1);
    uint64_t cap = (3 * avg + 256 + 63) & ~(uint64_t)63;
    if (h->s_relayout && cap < 2ull * h->s_cap) cap = 2ull * h->s_cap;
    if (h->s_cap_forced) cap = h->s_cap_forced;
    h->s_cap = (uint32_t)std::min<uint64_t>(cap, 1u << 20);
    h->s_relayout = false;
  }
  for (int attempt = 0; ; ++attempt) {
    h->s_ovf_cap = n + n / 2 + 65536;
    h->s_n_layout = n;
    for (int q = 0; q < 2; ++q) {
      if ((rc = ensure(h, h->s_rec[q], tiles * h->s_cap * sizeof(SRec)))) return rc;
      if ((rc = ensure(h, h->s_cnt[q], (tiles + 8) * 4))) return rc;
      if ((rc = ensure(h, h->s_ovf_rec[q], h->s_ovf_cap * sizeof(SRec)))) return rc;
      if ((rc = ensure(h, h->s_ovf_tile[q], h->s_ovf_cap * 4))) return rc;
    }
    h->s_ovf_cap = std::min<uint64_t>(h->s_ovf_rec[0].bytes / sizeof(SRec), h->s_ovf_tile[0].bytes / 4); /* the head room of ensure() counts */
    h->s_ovf_cap = std::min<uint64_t>(h->s_ovf_cap, std::min<uint64_t>(h->s_ovf_rec[1].bytes / sizeof(SRec), h->s_ovf_tile[1].bytes / 4));
    const int cur = h->s_cur;
    h->s_words[cur] = stream_bitmap_words(key_space(h, h->next_id + 1));
    if ((rc = ensure(h, h->s_bits[cur], h->s_words[cur] * 4))) return rc;
    if (h->strip && h->n_ranks > 1) {
      /* the same on every rank (it bounds the all-gathered message): from the grid alone */
      h->s_parent_cap = std::max<uint64_t>(65536, (uint64_t)h->cfg.nx * (uint64_t)h->cfg.ny / (uint64_t)h->n_ranks / 32);
      if ((rc = ensure(h, h->s_parents, (h->s_parent_cap + 1) * 4))) return rc;
    }
    ABM_CK(h, dev_memset(h->s_bits[cur].p, 0, h->s_words[cur] * 4, h->stream));
    ABM_CK(h, dev_memset(h->s_cnt[cur].p, 0, tiles * 4, h->stream));
    ABM_CK(h, dev_memset(sc, 0, SC_WORDS * 4, h->stream));
    if (n > 0) {
      StreamIngestBody ib;
      ib.v = make_view(h); ib.v.err = sc + SC_ERR; ib.out = stream_lists(h, cur); ib.ps = stream_sink(h, cur); ib.n = (uint32_t)n;
      { ProfScope ps(h->prof, h->stream, "stream_ingest", n); launch_foreach(h->stream, ib, (n + 31) & ~(uint64_t)31); }
      if ((rc = check_launch(h, "stream ingest"))) return rc;
    }
    if ((rc = stream_read_ctr(h))) return rc;
    if ((rc = stream_error_to_code(h))) return rc;
    h->s_n_ovf = h->h_counters[cur ? SC_OVF1 : SC_OVF0];
    if (h->s_n_ovf <= std::max<uint64_t>(4096, n / 64) || h->s_cap_forced || h->s_cap >= (1u << 20) || attempt >= 12) break;
    h->s_cap *= 2; /* crowded cores: twice the slots, file again */
  }
  h->s_lists = true;
  if (h->s_parent_used == 0 && (rc = stream_agree_parent_cap(h, true))) return rc; /* collective: every rank files its first lists in the same abm_step */
  return 0;
}

/* arrival lists -> canonical SoA (h->cell/id/energy[h->cur], any order), for the getters of the ABI */
static int stream_canonical(abm_handle* h) {
  if (!h->stream_mode || h->s_canon || !h->s_lists) return 0;
  int rc;
  if ((rc = set_device(h))) return rc;
  const uint64_t n = h->n;
  const uint64_t tiles = (uint64_t)h->g.tiles_x * (uint64_t)h->g.tiles_y;
  if ((rc = ensure_soa(h, h->cur, n))) return rc;
  uint32_t* sc = (uint32_t*)h->s_ctr.p;
  ABM_CK(h, dev_memset(sc + SC_CURSOR, 0, 4, h->stream));
  StreamExportCta ex;
  ex.g = h->g; ex.in = stream_lists(h, h->s_cur);
  ex.cell = (uint32_t*)h->cell[h->cur].p; ex.id = (uint32_t*)h->id[h->cur].p; ex.energy = (double*)h->energy[h->cur].p; ex.cursor = sc + SC_CURSOR;
  launch_cta(h->stream, ex, tiles, S_THREADS, StreamExportCta::smem_bytes());
  if (h->s_n_ovf > 0) {
    StreamExportOvfBody ob;
    ob.ex = ex; ob.base = (uint32_t)(n - h->s_n_ovf);
    launch_foreach(h->stream, ob, h->s_n_ovf);
  }
  if ((rc = check_launch(h, "stream export"))) return rc;
  ABM_CK(h, stream_sync(h->stream));
  h->s_canon = true;
  return 0;
}

/* the host is about to change something the filed steps depend on (tick, terrain, population): back to the canonical form */
static int stream_invalidate(abm_handle* h) {
  if (!h->stream_mode) return 0;
  int rc = stream_canonical(h);
  if (rc) return rc;
  h->s_lists = false;
  return 0;
}

#include "abm_host_stream_strips.inl"

/* one tick on the stream path: [ghost lists + parents exchange] -> birth prefix -> ONE kernel */
static int tick_stream(abm_handle* h) {
  int rc;
  if (!h->s_lists && (rc = stream_ingest(h))) return rc;
  const uint64_t n = h->n;
  const int cur = h->s_cur, nxt = cur ^ 1;
  const uint64_t tiles = (uint64_t)h->g.tiles_x * (uint64_t)h->g.tiles_y;
  const int cores_x = h->g.nx / CORE, core_rows = (h->g.own_hi - h->g.own_lo) / CORE;
  uint32_t* sc = (uint32_t*)h->s_ctr.p;
  /* Strips: arrivals from the neighbours, and every strip's parents.  Over peer memory the list exchange runs on the side
   * stream and only the two boundary core rows wait for it: the interior rows' kernel hides it.  (Not in a tick whose lists
   * overflowed: an interior core would then scan the overflow list while the unpack appends to it.) */
  bool lists_beside = false;
#ifndef ABM_EMU
  lists_beside = h->strip && h->p2p_on && h->n_ranks > 1 && h->overlap_lists && core_rows >= 3 && h->s_n_ovf == 0;
#endif
  if (h->strip) {
    if (lists_beside) {
      event_record(h->ev_rows, h->stream);
      stream_wait_event(h->side, h->ev_rows);
      std::swap(h->stream, h->side);
      rc = stream_exchange_lists(h);
      event_record(h->ev_halo, h->stream);
      std::swap(h->stream, h->side);
      if (rc) return rc;
    } else if ((rc = stream_exchange_lists(h))) return rc;
    if ((rc = stream_exchange_parents(h))) return rc;
  }
  if (h->cfg.reproduce && (rc = stream_prefix(h, (const uint32_t*)h->s_bits[cur].p, h->s_words[cur]))) return rc;
  /* the lists and the parents bitmap of the next tick start empty */
  /* parents of tick + 1 have ids below next_id + births(tick); births = this tick's parents: at most the agents alive on a
   * single handle, at most what the ranks' parents messages can carry on strips (a fuller message is an error) */
  h->s_words[nxt] = stream_bitmap_words(key_space(h, h->next_id + ((h->strip && h->n_ranks > 1) ? (uint64_t)h->n_ranks * h->s_parent_used : n) + 1));
  if ((rc = ensure(h, h->s_bits[nxt], h->s_words[nxt] * 4))) return rc;
  if (h->cfg.reproduce) ABM_CK(h, dev_memset(h->s_bits[nxt].p, 0, h->s_words[nxt] * 4, h->stream));
  ABM_CK(h, dev_memset(h->s_cnt[nxt].p, 0, tiles * 4, h->stream));
  ABM_CK(h, dev_memset(sc + SC_EMITTED, 0, 4, h->stream));
  ABM_CK(h, dev_memset(sc + (nxt ? SC_OVF1 : SC_OVF0), 0, 4, h->stream));
  ABM_CK(h, dev_memset(sc + SC_PARENTS, 0, 4, h->stream));
  auto launch_tick = [&](auto tk, int row0, int nrows, int row_step) { /* one instantiation per scheduler: by_id carries no key arithmetic */
    tk.v = make_view(h); tk.v.err = sc + SC_ERR;
    tk.in = stream_lists(h, cur); tk.out = stream_lists(h, nxt); tk.ps = stream_sink(h, nxt);
    tk.cores_x = cores_x; tk.core_y0 = h->g.own_lo / CORE + row0; tk.row_step = row_step;
    tk.next_id = (uint32_t)h->next_id;
    tk.birth_bits = (const uint32_t*)h->s_bits[cur].p; tk.birth_prefix = (const uint32_t*)h->bits_prefix.p;
    tk.emitted = sc + SC_EMITTED;
    tk.births_total = h->cfg.reproduce ? (const uint32_t*)h->bits_prefix.p + (h->s_words[cur] / 8 - 1) : sc + SC_ZERO;
    ProfScope ps(h->prof, h->stream, "stream_tick", row0 == 0 ? n : 0, row0 == 0 ? 1 : 0);
    launch_cta_lb<S_THREADS, S_MIN_CTAS>(h->stream, tk, (uint64_t)cores_x * (uint64_t)nrows, tk.smem_bytes());
  };
  auto launch_rows = [&](int row0, int nrows, int row_step) {
    if (h->cfg.scheduler == ABM_SCHED_RANDOMLY) launch_tick(StreamTickCta<true>(), row0, nrows, row_step); else launch_tick(StreamTickCta<false>(), row0, nrows, row_step);
  };
  if (lists_beside) { /* interior rows, then (arrivals in place) the first and the last core row in one launch */
    launch_rows(1, core_rows - 2, 1);
    stream_wait_event(h->stream, h->ev_halo);
    launch_rows(0, 2, core_rows - 1);
  } else launch_rows(0, core_rows, 1);
  if ((rc = check_launch(h, "stream tick"))) return rc;
  /* the tick's figures: records filed, overflow, births (= total of the parents bitmap) */
  ABM_CK(h, copy_d2h_async(h->h_counters, sc, SC_WORDS * sizeof(uint32_t), h->stream));
  h->h_counters[SC_HOST_BIRTHS] = 0;
  if (h->cfg.reproduce) ABM_CK(h, copy_d2h_async(h->h_counters + SC_HOST_BIRTHS, (const uint32_t*)h->bits_prefix.p + (h->s_words[cur] / 8 - 1), 4, h->stream));
  ABM_CK(h, stream_sync(h->stream));
  if ((rc = stream_error_to_code(h))) return rc;
  const uint64_t births = h->h_counters[SC_HOST_BIRTHS];
  if (h->next_id + births > 0xFFFFFFFFull) ABM_FAIL(h, ABM_E_CAPACITY, "agent id space (2^32-1) exhausted");
  h->updates += (double)n;
  h->n = h->h_counters[SC_EMITTED];
  h->next_id += births;
  h->tick += 1;
  h->s_cur = nxt;
  h->s_n_ovf = h->h_counters[nxt ? SC_OVF1 : SC_OVF0];
  h->s_canon = false;
  if ((rc = stream_agree_parent_cap(h, false))) return rc;
  /* crowded cores (or a population that outgrew the layout): file everything again with more slots per core */
  if (!h->s_cap_forced && (h->s_n_ovf > std::max<uint64_t>(4096, h->n / 64) || h->n + h->n / 4 > h->s_ovf_cap)) {
    if ((rc = stream_canonical(h))) return rc;
    h->s_lists = false;
    h->s_relayout = true;
  }
  return 0;
}
