/*
 * abm_host_stream_strips.inl — row strips on the stream path: at the start of a tick the lists of the two ghost tile rows
 * travel to the neighbouring strips and every strip's parents are all-gathered.  Over peer memory the two exchanges run
 * side by side: the lists on the side stream, hidden behind the tick kernel of the interior core rows (tick_stream).
 * Part of abm_engine.cu (included by abm_host_stream.inl); not a stand-alone header.
 */

/* records one ghost-row message can carry: the same on every rank (derived from the grid alone) */
static uint64_t stream_msg_cap(abm_handle* h) {
  uint64_t cap = (uint64_t)h->g.tiles_x * 256;
  if (const char* e = getenv("ABM_STREAM_MSG_PER_CORE")) { const long q = atol(e); if (q >= 1 && q <= 65536) cap = (uint64_t)h->g.tiles_x * (uint64_t)q; }
#ifndef ABM_EMU
  if (h->p2p_on) {
    const uint64_t hdr = stream_hdr_bytes(h->g.tiles_x);
    const uint64_t fit = h->p2p_lay.agents_bytes > hdr ? (h->p2p_lay.agents_bytes - hdr) / sizeof(SRec) : 0;
    cap = std::min(cap, fit);
  }
#endif
  return cap;
}

struct StreamParentCountBody { /* [count][ids ...]: the count in front of the list the kernels appended to */
  const uint32_t* count; uint32_t cap; uint32_t* msg; uint32_t* err;
  ABM_HD void operator()(uint64_t) const { const uint32_t c = *count; if (c > cap) atomic_or(err, ERR_LIST_CAP); msg[0] = c < cap ? c : cap; }
};

/* the arrivals from the neighbouring strips: my two ghost-row lists travel, theirs join the lists of my boundary core rows */
static int stream_exchange_lists(abm_handle* h) {
  int rc;
  const int cur = h->s_cur, P = h->n_ranks, tiles_x = h->g.tiles_x;
  uint32_t* sc = (uint32_t*)h->s_ctr.p;
  h->s_msg_cap = stream_msg_cap(h);
  const uint64_t bytes = stream_msg_bytes(tiles_x, h->s_msg_cap);
  StreamPackCta pk;
  pk.in = stream_lists(h, cur); pk.tiles_x = tiles_x; pk.ghost_row[0] = h->g.own_lo / CORE - 1; pk.ghost_row[1] = h->g.own_hi / CORE;
  pk.cap = (uint32_t)h->s_msg_cap; pk.err = sc + SC_ERR; pk.seq = 0;
  StreamUnpackCta up;
  up.out = stream_lists(h, cur); up.tiles_x = tiles_x; up.own_row[0] = h->g.own_lo / CORE; up.own_row[1] = h->g.own_hi / CORE - 1;
  up.cap = (uint32_t)h->s_msg_cap; up.err = sc + SC_ERR; up.seq = 0;
  bool direct = false;
  for (int d = 0; d < 2; ++d) { pk.msg[d] = nullptr; pk.publish[d] = nullptr; up.msg[d] = nullptr; up.wait[d] = nullptr; }
#ifndef ABM_EMU
  if (h->p2p_on) {
    const uint32_t seq = ++h->p2p_seq[P2P_CH_AGENTS];
    const int parity = (int)(seq & 1u);
    const int prev = (h->my_rank + P - 1) % P, next = (h->my_rank + 1) % P;
    const P2PLayout& L = h->p2p_lay;
    pk.seq = up.seq = seq;
    if (h->has_prev) { pk.msg[0] = h->p2p_peer[prev] + L.agents_off(1, parity); pk.publish[0] = (uint32_t*)(h->p2p_peer[prev] + L.flag_off(P2P_CH_AGENTS, 1, parity)); }
    if (h->has_next) { pk.msg[1] = h->p2p_peer[next] + L.agents_off(0, parity); pk.publish[1] = (uint32_t*)(h->p2p_peer[next] + L.flag_off(P2P_CH_AGENTS, 0, parity)); }
    if (h->has_prev) { up.msg[0] = h->p2p_mine + L.agents_off(0, parity); up.wait[0] = (const uint32_t*)(h->p2p_mine + L.flag_off(P2P_CH_AGENTS, 0, parity)); }
    if (h->has_next) { up.msg[1] = h->p2p_mine + L.agents_off(1, parity); up.wait[1] = (const uint32_t*)(h->p2p_mine + L.flag_off(P2P_CH_AGENTS, 1, parity)); }
    direct = true;
  }
#endif
  if (!direct) {
    for (int d = 0; d < 2; ++d) {
      if ((rc = ensure(h, h->s_tx[d], bytes))) return rc;
      if ((rc = ensure(h, h->s_rx[d], bytes))) return rc;
    }
    if (h->has_prev) { pk.msg[0] = (uint8_t*)h->s_tx[0].p; up.msg[0] = (const uint8_t*)h->s_rx[0].p; }
    if (h->has_next) { pk.msg[1] = (uint8_t*)h->s_tx[1].p; up.msg[1] = (const uint8_t*)h->s_rx[1].p; }
  }
  { ProfScope ps(h->prof, h->stream, "list_exchange", 2 * h->s_msg_cap);
    launch_cta(h->stream, pk, 2, TILE_THREADS, StreamPackCta::smem_bytes());
    if (!direct && (rc = comm_sendrecv(h, h->s_tx[0].p, bytes, h->s_tx[1].p, bytes, h->s_rx[0].p, bytes, h->s_rx[1].p, bytes))) return rc;
    launch_cta(h->stream, up, 2, TILE_THREADS, StreamUnpackCta::smem_bytes());
    StreamClearGhostBody cg;
    cg.cnt = (uint32_t*)h->s_cnt[cur].p; cg.tiles_x = tiles_x; cg.ghost_row[0] = pk.ghost_row[0]; cg.ghost_row[1] = pk.ghost_row[1];
    launch_foreach(h->stream, cg, 2 * (uint64_t)tiles_x); }
  return check_launch(h, "stream list exchange");
}

/* every strip's parents of this tick -> bitmap over the key space (the kernels appended the own ones to s_parents) */
static int stream_exchange_parents(abm_handle* h) {
  int rc;
  const int cur = h->s_cur, P = h->n_ranks;
  if (P <= 1 || !h->cfg.reproduce) return 0;
  uint32_t* sc = (uint32_t*)h->s_ctr.p;
  uint32_t* ctr = (uint32_t*)h->counters.p;
  const uint64_t cap = h->s_parent_used, words = 1 + cap;
  if ((rc = ensure(h, h->s_gather, words * 4 * (size_t)(P + 1) + 64))) return rc;
  uint32_t* msg = (uint32_t*)h->s_parents.p;
  uint32_t* all = (uint32_t*)h->s_gather.p;
  ProfScope ps(h->prof, h->stream, "parents_exchange", cap);
  StreamParentCountBody pc;
  pc.count = sc + SC_PARENTS; pc.cap = (uint32_t)cap; pc.msg = msg; pc.err = sc + SC_ERR;
  launch_foreach(h->stream, pc, 1);
  bool gathered = false;
#ifndef ABM_EMU
  if (h->p2p_on) {
    if ((1 + cap) * 4 > h->p2p_lay.births_bytes) ABM_FAIL(h, ABM_E_CAPACITY, "parents message (%llu ids) exceeds the mailbox allocated at abm_p2p_export", (unsigned long long)cap);
    const uint32_t seq = ++h->p2p_seq[P2P_CH_BIRTHS];
    const int parity = (int)(seq & 1u);
    P2PBirthSendCta sd;
    sd.birth_parent = msg + 1; sd.birth_count = msg; sd.cap = (uint32_t)cap; sd.seq = seq; sd.done = ctr + CTR_BIRTH_DONE;
    P2PBirthRecvCta rv;
    rv.seq = seq; rv.cap = (uint32_t)cap; rv.all = all; rv.err = sc + SC_ERR;
    for (int r = 0; r < P; ++r) {
      sd.dst[r] = (uint32_t*)(h->p2p_peer[r] + h->p2p_lay.births_off(h->my_rank, parity));
      sd.flag[r] = (uint32_t*)(h->p2p_peer[r] + h->p2p_lay.flag_off(P2P_CH_BIRTHS, h->my_rank, parity));
      rv.src[r] = (const uint32_t*)(h->p2p_mine + h->p2p_lay.births_off(r, parity));
      rv.flag[r] = (const uint32_t*)(h->p2p_mine + h->p2p_lay.flag_off(P2P_CH_BIRTHS, r, parity));
    }
    launch_cta(h->stream, sd, (uint64_t)P * BIRTH_CTAS, 256, 16);
    launch_cta(h->stream, rv, (uint64_t)P * BIRTH_CTAS, 256, 16);
    gathered = true;
  } else if (h->nccl_comm) {
    ABM_NCCL(h, g_nccl.AllGather(msg, all, (size_t)words, ncclUint32, (ncclComm_t)h->nccl_comm, h->stream));
    gathered = true;
  }
#endif
  if (!gathered) {
    std::vector<uint32_t> hs((size_t)words), hr((size_t)words * P);
    ABM_CK(h, copy_d2h(hs.data(), msg, words * 4, h->stream));
    if ((rc = comm_allgather_u32(h, hs.data(), hr.data(), (uint32_t)words))) return rc;
    ABM_CK(h, copy_h2d(all, hr.data(), words * 4 * (size_t)P, h->stream));
  }
  StreamMarkParentsBody mk;
  mk.all = all; mk.n_ranks = P; mk.cap = (uint32_t)cap; mk.bits = (uint32_t*)h->s_bits[cur].p; mk.counts_out = sc + SC_GCOUNT;
  launch_foreach(h->stream, mk, (uint64_t)P * std::max<uint64_t>(cap, 1));
  return check_launch(h, "stream parents exchange");
}

/* the capacity (ids) of the next parents message, the same on every rank: from the gathered counts of this tick */
static int stream_agree_parent_cap(abm_handle* h, bool first) {
  const int P = h->n_ranks;
  if (!h->strip || P <= 1 || !h->cfg.reproduce) return 0;
  int rc;
  uint64_t bound = 0;
  if (first) { /* after filing: every rank's local count */
    const uint64_t v = h->h_counters[SC_PARENTS];
    std::vector<uint64_t> all((size_t)P);
    if ((rc = comm_allgather_small(h, &v, 1, all.data()))) return rc;
    for (int r = 0; r < P; ++r) bound = std::max(bound, all[r]);
  } else { /* the counts every rank gathered this tick (StreamMarkParentsBody left them beside the tick's counters) */
    for (int r = 0; r < P; ++r) bound = std::max<uint64_t>(bound, h->h_counters[SC_GCOUNT + r]);
  }
  h->s_parent_used = std::min<uint64_t>(2 * bound + 4096, h->s_parent_cap);
  return 0;
}
