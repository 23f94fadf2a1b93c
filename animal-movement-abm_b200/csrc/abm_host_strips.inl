/*
 * abm_host_strips.inl — host side of the row strips: ring transports (NCCL, host callbacks, peer memory), halo / decision / counter / offspring exchanges
 * Part of abm_engine.cu (one translation unit: included there, inside its anonymous namespace where it says so);
 * not a stand-alone header.
 */
/* ------------------------------------------------------------------------------------------ ring transport */
#ifndef ABM_EMU
struct NcclApi {
  void* lib = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  bool load() {
    if (lib) return true;
    lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL); /* the copy the host process already loaded (torch's), if any */
    if (!lib) lib = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!lib) return false;
#define ABM_NCCL_SYM(name) *(void**)(&name) = dlsym(lib, "nccl" #name); if (!name) return false;
    ABM_NCCL_SYM(GetUniqueId) ABM_NCCL_SYM(CommInitRank) ABM_NCCL_SYM(CommDestroy) ABM_NCCL_SYM(Send) ABM_NCCL_SYM(Recv)
    ABM_NCCL_SYM(GroupStart) ABM_NCCL_SYM(GroupEnd) ABM_NCCL_SYM(AllReduce) ABM_NCCL_SYM(AllGather) ABM_NCCL_SYM(GetErrorString)
#undef ABM_NCCL_SYM
    return true;
  }
};
static NcclApi g_nccl;
#define ABM_NCCL(h, call) do { ncclResult_t r_ = (call); if (r_ != ncclSuccess) ABM_FAIL(h, ABM_E_CUDA, "%s failed: %s", #call, g_nccl.GetErrorString(r_)); } while (0)
#endif

/* boundary exchange with the ring neighbours; all pointers are device memory of this handle */
static int comm_sendrecv(abm_handle* h, const void* to_prev, uint64_t b_tp, const void* to_next, uint64_t b_tn, void* from_prev, uint64_t b_fp, void* from_next, uint64_t b_fn,
                         unsigned long long* d_allreduce = nullptr, int n_allreduce = 0) {
  if (!h->has_prev) { b_tp = 0; b_fp = 0; }
  if (!h->has_next) { b_tn = 0; b_fn = 0; }
  if (h->n_ranks == 1) { /* a ring of one: my first rows are the ghost rows below my last rows and vice versa */
    if (b_tp != b_fn || b_tn != b_fp) ABM_FAIL(h, ABM_E_INTERNAL, "self exchange size mismatch");
    if (b_tp) ABM_CK(h, copy_d2d(from_next, to_prev, b_tp, h->stream));
    if (b_tn) ABM_CK(h, copy_d2d(from_prev, to_next, b_tn, h->stream));
    return 0;
  }
  const int prev = (h->my_rank + h->n_ranks - 1) % h->n_ranks, next = (h->my_rank + 1) % h->n_ranks;
#ifndef ABM_EMU
  if (h->nccl_comm) {
    ncclComm_t c = (ncclComm_t)h->nccl_comm;
    ABM_NCCL(h, g_nccl.GroupStart()); /* one launch: the ring exchange and, when asked, the all-reduce of the device counters */
    if (d_allreduce) ABM_NCCL(h, g_nccl.AllReduce(d_allreduce, d_allreduce, (size_t)n_allreduce, ncclUint64, ncclSum, c, h->stream));
    if (b_tp) ABM_NCCL(h, g_nccl.Send(to_prev, b_tp, ncclUint8, prev, c, h->stream));
    if (b_tn) ABM_NCCL(h, g_nccl.Send(to_next, b_tn, ncclUint8, next, c, h->stream));
    if (b_fn) ABM_NCCL(h, g_nccl.Recv(from_next, b_fn, ncclUint8, next, c, h->stream));
    if (b_fp) ABM_NCCL(h, g_nccl.Recv(from_prev, b_fp, ncclUint8, prev, c, h->stream));
    ABM_NCCL(h, g_nccl.GroupEnd());
    return 0;
  }
#endif
  (void)prev; (void)next;
  if (!h->has_tr) ABM_FAIL(h, ABM_E_STATE, "n_ranks > 1 needs abm_comm_init_nccl or abm_set_transport before abm_step");
  const void* src[2] = {to_prev, to_next};
  const uint64_t sb[2] = {b_tp, b_tn}, rb[2] = {b_fp, b_fn};
  for (int q = 0; q < 2; ++q) {
    h->hstage[q].resize(sb[q] + 8);
    h->hstage[2 + q].resize(rb[q] + 8);
    if (sb[q]) ABM_CK(h, copy_d2h(h->hstage[q].data(), src[q], sb[q], h->stream));
  }
  ABM_CK(h, stream_sync(h->stream));
  if (h->tr.sendrecv(h->tr.user, h->hstage[0].data(), b_tp, h->hstage[1].data(), b_tn, h->hstage[2].data(), b_fp, h->hstage[3].data(), b_fn) != 0)
    ABM_FAIL(h, ABM_E_INTERNAL, "transport sendrecv failed");
  if (b_fp) ABM_CK(h, copy_h2d(from_prev, h->hstage[2].data(), b_fp, h->stream));
  if (b_fn) ABM_CK(h, copy_h2d(from_next, h->hstage[3].data(), b_fn, h->stream));
  return 0;
}

static int comm_allreduce_u64(abm_handle* h, uint64_t* vals, int n) {
  if (h->n_ranks == 1) return 0;
#ifndef ABM_EMU
  if (h->nccl_comm) {
    int rc = ensure(h, h->red_buf, (size_t)n * 8);
    if (rc) return rc;
    ABM_CK(h, copy_h2d(h->red_buf.p, vals, (size_t)n * 8, h->stream));
    ABM_NCCL(h, g_nccl.AllReduce(h->red_buf.p, h->red_buf.p, (size_t)n, ncclUint64, ncclSum, (ncclComm_t)h->nccl_comm, h->stream));
    ABM_CK(h, copy_d2h(vals, h->red_buf.p, (size_t)n * 8, h->stream));
    return 0;
  }
#endif
  if (!h->has_tr) ABM_FAIL(h, ABM_E_STATE, "n_ranks > 1 needs abm_comm_init_nccl or abm_set_transport before abm_step");
  if (h->tr.allreduce_sum_u64(h->tr.user, vals, n) != 0) ABM_FAIL(h, ABM_E_INTERNAL, "transport allreduce failed");
  return 0;
}

/* every rank contributes n u32; recv holds n_ranks * n */
static int comm_allgather_u32(abm_handle* h, const uint32_t* send, uint32_t* recv, uint32_t n) {
  if (h->n_ranks == 1) { memcpy(recv, send, (size_t)n * 4); return 0; }
#ifndef ABM_EMU
  if (h->nccl_comm) {
    int rc = ensure(h, h->red_buf, (size_t)n * 4 * (size_t)(h->n_ranks + 1));
    if (rc) return rc;
    uint32_t* d_send = (uint32_t*)h->red_buf.p;
    uint32_t* d_recv = d_send + n;
    ABM_CK(h, copy_h2d(d_send, send, (size_t)n * 4, h->stream));
    ABM_NCCL(h, g_nccl.AllGather(d_send, d_recv, (size_t)n, ncclUint32, (ncclComm_t)h->nccl_comm, h->stream));
    ABM_CK(h, copy_d2h(recv, d_recv, (size_t)n * 4 * (size_t)h->n_ranks, h->stream));
    return 0;
  }
#endif
  if (!h->has_tr) ABM_FAIL(h, ABM_E_STATE, "n_ranks > 1 needs abm_comm_init_nccl or abm_set_transport before abm_step");
  if (h->tr.allgather_u32(h->tr.user, send, recv, n) != 0) ABM_FAIL(h, ABM_E_INTERNAL, "transport allgather failed");
  return 0;
}


#ifndef ABM_EMU
/* ------------------------------------------------------------------------------- peer-memory transport */

/* the counters of this round from every rank, summed into red_buf: two launches (send + publish, wait + sum [+ guard]) */
static void p2p_stats_send(abm_handle* h, const uint32_t* d_pending) {
  const uint32_t seq = ++h->p2p_seq[P2P_CH_STATS];
  const int parity = (int)(seq & 1u), P = h->n_ranks;
  uint32_t* ctr = (uint32_t*)h->counters.p;
  P2PStatsSendCta sc;
  sc.pending = d_pending; sc.birth_flags = ctr + CTR_BIRTH_FLAGS;
  sc.tot_up = (const uint32_t*)h->sx_pre[0].p + h->nsx; sc.tot_down = (const uint32_t*)h->sx_pre[1].p + h->nsx;
  sc.n_ranks = P; sc.rank = h->my_rank; sc.seq = seq;
  for (int r = 0; r < P; ++r) {
    sc.dst[r] = (unsigned long long*)(h->p2p_peer[r] + h->p2p_lay.stats_off(h->my_rank, parity));
    sc.flag[r] = (uint32_t*)(h->p2p_peer[r] + h->p2p_lay.flag_off(P2P_CH_STATS, h->my_rank, parity));
  }
  launch_cta(h->stream, sc, 1, 64, P2P_STATS_WORDS * 8);
}
static int p2p_stats_recv(abm_handle* h, uint32_t* go, uint32_t birth_cap) {
  const uint32_t seq = h->p2p_seq[P2P_CH_STATS];
  const int parity = (int)(seq & 1u), P = h->n_ranks;
  uint32_t* ctr = (uint32_t*)h->counters.p;
  P2PStatsRecvCta rc_;
  rc_.n_ranks = P; rc_.seq = seq; rc_.out = (unsigned long long*)h->red_buf.p; rc_.err = ctr + CTR_ERR; rc_.go = go; rc_.birth_cap = birth_cap;
  for (int r = 0; r < P; ++r) {
    rc_.src[r] = (const unsigned long long*)(h->p2p_mine + h->p2p_lay.stats_off(r, parity));
    rc_.flag[r] = (const uint32_t*)(h->p2p_mine + h->p2p_lay.flag_off(P2P_CH_STATS, r, parity));
  }
  launch_cta(h->stream, rc_, 1, 64, 16);
  return check_launch(h, "peer counters");
}

/* every rank's offspring message [count, parent ids ...] (at most cap ids) into all[rank * (1 + cap) ...] */
static int p2p_allgather_births(abm_handle* h, uint32_t* all, uint64_t cap) {
  if ((1 + cap) * 4 > h->p2p_lay.births_bytes) ABM_FAIL(h, ABM_E_CAPACITY, "offspring message (%llu ids) exceeds the mailbox allocated at abm_p2p_export", (unsigned long long)cap);
  const uint32_t seq = ++h->p2p_seq[P2P_CH_BIRTHS];
  const int parity = (int)(seq & 1u), P = h->n_ranks;
  uint32_t* ctr = (uint32_t*)h->counters.p;
  P2PBirthSendCta sc;
  sc.birth_parent = (const uint32_t*)h->birth_parent.p; sc.birth_count = ctr + CTR_BIRTHS; sc.cap = (uint32_t)cap; sc.seq = seq; sc.done = ctr + CTR_BIRTH_DONE;
  for (int r = 0; r < P; ++r) {
    sc.dst[r] = (uint32_t*)(h->p2p_peer[r] + h->p2p_lay.births_off(h->my_rank, parity));
    sc.flag[r] = (uint32_t*)(h->p2p_peer[r] + h->p2p_lay.flag_off(P2P_CH_BIRTHS, h->my_rank, parity));
  }
  ProfScope ps(h->prof, h->stream, "offspring_exchange", cap);
  launch_cta(h->stream, sc, (uint64_t)P * BIRTH_CTAS, 256, 16);
  P2PBirthRecvCta rv;
  rv.seq = seq; rv.cap = (uint32_t)cap; rv.all = all; rv.err = ctr + CTR_ERR;
  for (int r = 0; r < P; ++r) {
    rv.src[r] = (const uint32_t*)(h->p2p_mine + h->p2p_lay.births_off(r, parity));
    rv.flag[r] = (const uint32_t*)(h->p2p_mine + h->p2p_lay.flag_off(P2P_CH_BIRTHS, r, parity));
  }
  launch_cta(h->stream, rv, (uint64_t)P * BIRTH_CTAS, 256, 16);
  return check_launch(h, "peer all-gather");
}
#endif

/* every rank's n (<= 4) values, rank-major, on the host: the collective behind abm_reduce on strips */
static int comm_allgather_small(abm_handle* h, const uint64_t* mine, int n, uint64_t* all) {
  const int P = h->n_ranks;
  if (P == 1) { memcpy(all, mine, (size_t)n * 8); return 0; }
#ifndef ABM_EMU
  if (h->p2p_on) {
    int rc = ensure(h, h->red_buf, 4096);
    if (rc) return rc;
    unsigned long long* d_src = (unsigned long long*)h->red_buf.p;
    unsigned long long* d_out = d_src + 8;
    ABM_CK(h, copy_h2d(d_src, mine, (size_t)n * 8, h->stream));
    const uint32_t seq = ++h->p2p_seq[P2P_CH_STATS];
    const int parity = (int)(seq & 1u);
    P2PSmallSendCta sd;
    sd.src = d_src; sd.n = n; sd.n_ranks = P; sd.seq = seq;
    P2PSmallRecvCta rv;
    rv.n = n; rv.n_ranks = P; rv.seq = seq; rv.out = d_out; rv.err = (uint32_t*)h->counters.p + CTR_ERR;
    for (int r = 0; r < P; ++r) {
      sd.dst[r] = (unsigned long long*)(h->p2p_peer[r] + h->p2p_lay.stats_off(h->my_rank, parity));
      sd.flag[r] = (uint32_t*)(h->p2p_peer[r] + h->p2p_lay.flag_off(P2P_CH_STATS, h->my_rank, parity));
      rv.src[r] = (const unsigned long long*)(h->p2p_mine + h->p2p_lay.stats_off(r, parity));
      rv.flag[r] = (const uint32_t*)(h->p2p_mine + h->p2p_lay.flag_off(P2P_CH_STATS, r, parity));
    }
    launch_cta(h->stream, sd, 1, 64, 16);
    launch_cta(h->stream, rv, 1, 64, 16);
    if ((rc = check_launch(h, "peer all-gather"))) return rc;
    ABM_CK(h, copy_d2h(all, d_out, (size_t)n * 8 * (size_t)P, h->stream));
    return 0;
  }
#endif
  std::vector<uint32_t> snd((size_t)2 * n), rcv((size_t)2 * n * P);
  memcpy(snd.data(), mine, (size_t)n * 8);
  int rc = comm_allgather_u32(h, snd.data(), rcv.data(), (uint32_t)(2 * n));
  if (rc) return rc;
  memcpy(all, rcv.data(), (size_t)n * 8 * (size_t)P);
  return 0;
}

/* capacity of a boundary message, agreed by all ranks from the same all-reduced figures */
static uint64_t strip_cap_for(uint64_t max_row) { return 2 * max_row + 4096; }

static int strip_alloc(abm_handle* h) {
  int rc;
  const uint64_t bytes = strip_msg_bytes(h->nsx, h->bcap);
  for (int d = 0; d < 2; ++d) {
    if ((rc = ensure(h, h->sx_buf[d], bytes))) return rc;
    if ((rc = ensure(h, h->rx_buf[d], bytes))) return rc;
    if ((rc = ensure(h, h->sx_pre[d], ((size_t)h->nsx + 8) * 4))) return rc;
    if ((rc = ensure(h, h->rx_pre[d], ((size_t)h->nsx + 8) * 4))) return rc;
    if ((rc = ensure(h, h->sx_st[d], h->bcap + 64))) return rc;
    if ((rc = ensure(h, h->rx_st[d], h->bcap + 64))) return rc;
  }
  return ensure(h, h->red_buf, 4096);
}

/* E1: this strip's first and last sub-tile rows (cell, id, energy, proposal) become the neighbours' ghost rows.
 * One fixed-size message per neighbour; header, prefix and packing by one CTA per direction: no host round trip. */
/* room for this tick's ghost agents: they live at fixed places behind the own agents, [n, n + cap) from the previous
 * strip and [n + cap, n + 2 cap) from the next.  Called before any kernel of the tick takes pointers. */
static int strip_prepare(abm_handle* h) {
  int rc;
  const int cur = h->cur;
  if ((rc = strip_alloc(h))) return rc;
  const uint64_t tot = h->n + 2 * h->bcap + 1;
  if ((rc = grow_keep(h, h->cell[cur], tot * 4))) return rc;
  if ((rc = grow_keep(h, h->id[cur], tot * 4))) return rc;
  if ((rc = grow_keep(h, h->energy[cur], tot * 8))) return rc;
  if ((rc = grow_keep(h, h->state, tot))) return rc;
  if ((rc = ensure(h, h->pend[0], tot * 4))) return rc;
  if ((rc = ensure(h, h->pend[1], tot * 4))) return rc;
  return 0;
}

static int strip_exchange_agents(abm_handle* h) {
  int rc;
  const int nsx = h->nsx, cur = h->cur;
  const uint64_t cap = h->bcap, hdr = strip_hdr_bytes(nsx), bytes = strip_msg_bytes(nsx, cap);
  const int chunks = (nsx + TILE_THREADS / 32 - 1) / (TILE_THREADS / 32);
  uint32_t* ctr = (uint32_t*)h->counters.p;
  uint32_t* err = ctr + CTR_ERR;
  uint8_t* tx_msg[2] = {(uint8_t*)h->sx_buf[0].p, (uint8_t*)h->sx_buf[1].p};
  const uint8_t* rx_msg[2] = {(const uint8_t*)h->rx_buf[0].p, (const uint8_t*)h->rx_buf[1].p};
  uint32_t* publish[2] = {nullptr, nullptr};
  const uint32_t* wait[2] = {nullptr, nullptr};
  uint32_t seq = 0;
  bool direct = false;
#ifndef ABM_EMU
  if (h->p2p_on) { /* the pack CTAs write straight into the neighbours' mailboxes and publish; the receiving CTAs wait for theirs */
    if (bytes > h->p2p_lay.agents_bytes) ABM_FAIL(h, ABM_E_CAPACITY, "halo message (%llu bytes) exceeds the mailbox allocated at abm_p2p_export", (unsigned long long)bytes);
    seq = ++h->p2p_seq[P2P_CH_AGENTS];
    const int parity = (int)(seq & 1u);
    const int prev = (h->my_rank + h->n_ranks - 1) % h->n_ranks, next = (h->my_rank + 1) % h->n_ranks;
    const P2PLayout& L = h->p2p_lay;
    if (h->has_prev) { tx_msg[0] = h->p2p_peer[prev] + L.agents_off(1, parity); publish[0] = (uint32_t*)(h->p2p_peer[prev] + L.flag_off(P2P_CH_AGENTS, 1, parity)); }
    if (h->has_next) { tx_msg[1] = h->p2p_peer[next] + L.agents_off(0, parity); publish[1] = (uint32_t*)(h->p2p_peer[next] + L.flag_off(P2P_CH_AGENTS, 0, parity)); }
    for (int d = 0; d < 2; ++d) {
      rx_msg[d] = h->p2p_mine + L.agents_off(d, parity); /* a side without a neighbour keeps its zeroed header: an empty ghost row */
      const bool has = d == 0 ? h->has_prev : h->has_next;
      wait[d] = has ? (const uint32_t*)(h->p2p_mine + L.flag_off(P2P_CH_AGENTS, d, parity)) : nullptr;
    }
    direct = true;
  }
#endif
  StripHeaderCta hd;
  hd.si = sub_index(h, cur); hd.row[0] = h->g.own_lo / SUB; hd.row[1] = h->g.own_hi / SUB - 1;
  StripPackCta pk;
  pk.si = hd.si; pk.row[0] = hd.row[0]; pk.row[1] = hd.row[1]; pk.tiles_x = h->g.tiles_x; pk.chunks = chunks; pk.cap = (uint32_t)cap;
  pk.cell = (const uint32_t*)h->cell[cur].p; pk.id = (const uint32_t*)h->id[cur].p; pk.energy = (const double*)h->energy[cur].p; pk.state = (const uint8_t*)h->state.p;
  pk.hdr = hdr; pk.err = err; pk.seq = seq;
  StripRxHeaderCta rh;
  rh.row[0] = h->g.own_lo / SUB - 1; rh.row[1] = h->g.own_hi / SUB; rh.nsx = nsx;
  rh.base[0] = (uint32_t)h->n; rh.base[1] = (uint32_t)(h->n + cap); rh.cap = (uint32_t)cap;
  rh.sub_off = (uint32_t*)h->sub_off[cur].p; rh.sub_cnt = (uint32_t*)h->sub_cnt[cur].p; rh.err = err; rh.seq = seq;
  StripUnpackCta up;
  up.nsx = nsx; up.chunks = chunks; up.cap = (uint32_t)cap; up.hdr = hdr;
  up.tile_row_base[0] = ((uint32_t)(h->g.own_lo / CORE - 1) * (uint32_t)h->g.tiles_x) << 10;
  up.tile_row_base[1] = ((uint32_t)(h->g.own_hi / CORE) * (uint32_t)h->g.tiles_x) << 10;
  up.base[0] = rh.base[0]; up.base[1] = rh.base[1];
  up.cell = (uint32_t*)h->cell[cur].p; up.id = (uint32_t*)h->id[cur].p; up.energy = (double*)h->energy[cur].p; up.state = (uint8_t*)h->state.p;
  for (int d = 0; d < 2; ++d) {
    hd.msg[d] = tx_msg[d]; hd.pre[d] = (uint32_t*)h->sx_pre[d].p; hd.done[d] = ctr + CTR_PACK_DONE + d;
    pk.msg[d] = tx_msg[d]; pk.pre[d] = hd.pre[d]; pk.done[d] = hd.done[d]; pk.publish[d] = publish[d];
    rh.msg[d] = rx_msg[d]; rh.pre[d] = (uint32_t*)h->rx_pre[d].p; rh.wait[d] = wait[d];
    up.msg[d] = rx_msg[d]; up.pre[d] = rh.pre[d];
  }
  { ProfScope ps(h->prof, h->stream, "halo_exchange", 2 * cap);
    launch_cta(h->stream, hd, 2, TILE_THREADS, StripHeaderCta::smem_bytes());
    launch_cta(h->stream, pk, 2 * (uint64_t)chunks, TILE_THREADS, 16);
    if (!direct) {
      for (int d = 0; d < 2; ++d) ABM_CK(h, dev_memset(h->rx_buf[d].p, 0, hdr, h->stream)); /* a missing neighbour sends nothing */
      if ((rc = comm_sendrecv(h, h->sx_buf[0].p, bytes, h->sx_buf[1].p, bytes, h->rx_buf[0].p, bytes, h->rx_buf[1].p, bytes))) return rc;
    }
    launch_cta(h->stream, rh, 2, TILE_THREADS, StripRxHeaderCta::smem_bytes());
    launch_cta(h->stream, up, 2 * (uint64_t)chunks, TILE_THREADS, 16); }
  h->n_ghost = 2 * cap; /* upper bound; the sub-tile index knows the real counts */
  return check_launch(h, "strip halo exchange");
}

/* E2: the decision bytes of the same boundary rows (same order), merged into the ghost copies */
static int strip_exchange_states(abm_handle* h, unsigned long long* d_allreduce = nullptr, int n_allreduce = 0) {
  int rc;
  const int cur = h->cur, nsx = h->nsx;
  const uint64_t cap = h->bcap;
  uint32_t* err = (uint32_t*)h->counters.p + CTR_ERR;
  PackStatesCta pk;
  pk.si = sub_index(h, cur); pk.row[0] = h->g.own_lo / SUB; pk.row[1] = h->g.own_hi / SUB - 1; pk.cap = (uint32_t)cap;
  pk.state = (const uint8_t*)h->state.p; pk.seq = 0;
  const int chunks = (nsx + TILE_THREADS / 32 - 1) / (TILE_THREADS / 32);
  pk.chunks = chunks;
  for (int d = 0; d < 2; ++d) {
    pk.pre[d] = (const uint32_t*)h->sx_pre[d].p; pk.out[d] = (uint8_t*)h->sx_st[d].p; pk.publish[d] = nullptr;
    pk.done[d] = (uint32_t*)h->counters.p + CTR_PACK_DONE + 2 + d;
  }
  MergeStatesCta mg;
  mg.state = (uint8_t*)h->state.p; mg.err = err; mg.seq = 0;
  for (int d = 0; d < 2; ++d) {
    const bool has = d == 0 ? h->has_prev : h->has_next;
    mg.incoming[d] = has ? (const uint8_t*)h->rx_st[d].p : nullptr; mg.base[d] = (uint32_t)(h->n + (uint64_t)d * cap);
    mg.total[d] = (const uint32_t*)h->rx_pre[d].p + nsx; mg.wait[d] = nullptr;
  }
  bool direct = false;
#ifndef ABM_EMU
  if (h->p2p_on) {
    if (cap > h->p2p_lay.state_bytes) ABM_FAIL(h, ABM_E_CAPACITY, "decision message exceeds the mailbox allocated at abm_p2p_export");
    const uint32_t seq = ++h->p2p_seq[P2P_CH_STATE];
    const int parity = (int)(seq & 1u);
    const int prev = (h->my_rank + h->n_ranks - 1) % h->n_ranks, next = (h->my_rank + 1) % h->n_ranks;
    const P2PLayout& L = h->p2p_lay;
    pk.seq = mg.seq = seq;
    pk.out[0] = h->has_prev ? h->p2p_peer[prev] + L.state_off(1, parity) : nullptr;
    pk.out[1] = h->has_next ? h->p2p_peer[next] + L.state_off(0, parity) : nullptr;
    pk.publish[0] = h->has_prev ? (uint32_t*)(h->p2p_peer[prev] + L.flag_off(P2P_CH_STATE, 1, parity)) : nullptr;
    pk.publish[1] = h->has_next ? (uint32_t*)(h->p2p_peer[next] + L.flag_off(P2P_CH_STATE, 0, parity)) : nullptr;
    for (int d = 0; d < 2; ++d) {
      const bool has = d == 0 ? h->has_prev : h->has_next;
      mg.incoming[d] = has ? h->p2p_mine + L.state_off(d, parity) : nullptr;
      mg.wait[d] = has ? (const uint32_t*)(h->p2p_mine + L.flag_off(P2P_CH_STATE, d, parity)) : nullptr;
    }
    direct = true;
  }
#endif
  { ProfScope ps(h->prof, h->stream, "decision_exchange", 2 * cap);
    launch_cta(h->stream, pk, 2 * (uint64_t)chunks, TILE_THREADS, 16);
    if (!direct && (rc = comm_sendrecv(h, h->sx_st[0].p, cap, h->sx_st[1].p, cap, h->rx_st[0].p, cap, h->rx_st[1].p, cap, d_allreduce, n_allreduce))) return rc;
    launch_cta(h->stream, mg, 2 * MERGE_CTAS, TILE_THREADS, 16); }
  return check_launch(h, "strip decision exchange");
}

/* true when the all-reduce of the counters can stay on the device (NCCL, or a ring of one) */
static bool strip_stats_on_device(abm_handle* h) {
#ifndef ABM_EMU
  if (h->p2p_on) return true;
#endif
  return h->nccl_comm != nullptr || h->n_ranks == 1;
}

static void strip_stats_kernel(abm_handle* h, const uint32_t* d_pending) {
  uint32_t* ctr = (uint32_t*)h->counters.p;
  StripStatsBody sb;
  sb.pending = d_pending; sb.birth_flags = ctr + CTR_BIRTH_FLAGS;
  sb.tot_up = (const uint32_t*)h->sx_pre[0].p + h->nsx; sb.tot_down = (const uint32_t*)h->sx_pre[1].p + h->nsx;
  sb.n_ranks = h->n_ranks; sb.rank = h->my_rank; sb.out = (unsigned long long*)h->red_buf.p;
  launch_foreach(h->stream, sb, 1);
}

/* counters all-reduced into red_buf and (optionally) the decision exchange, without any read-back */
static int strip_stats_enqueue(abm_handle* h, const uint32_t* d_pending, bool with_state_exchange, uint32_t* go = nullptr, uint32_t guard_cap = 0) {
  int rc;
  const int nv = 1 + 3 * h->n_ranks;
  unsigned long long* d = (unsigned long long*)h->red_buf.p;
#ifndef ABM_EMU
  if (h->p2p_on) { /* the commit guard is folded into the kernel that sums the counters */
    { ProfScope ps(h->prof, h->stream, "counters_exchange", (uint64_t)nv); p2p_stats_send(h, d_pending); rc = p2p_stats_recv(h, go, guard_cap); }
    if (rc) return rc;
    return with_state_exchange ? strip_exchange_states(h) : 0;
  }
#endif
  strip_stats_kernel(h, d_pending);
#ifndef ABM_EMU
  if (h->nccl_comm) {
    if (with_state_exchange) return strip_exchange_states(h, d, nv);
    ABM_NCCL(h, g_nccl.AllReduce(d, d, (size_t)nv, ncclUint64, ncclSum, (ncclComm_t)h->nccl_comm, h->stream));
    return 0;
  }
#endif
  (void)nv; (void)d;
  if (with_state_exchange && (rc = strip_exchange_states(h))) return rc; /* ring of one: the local figures are the global ones */
  return 0;
}

static int strip_stats(abm_handle* h, const uint32_t* d_pending, std::vector<uint64_t>& out, bool with_state_exchange = false) {
  int rc;
  const int P = h->n_ranks, nv = 1 + 3 * P;
  unsigned long long* d = (unsigned long long*)h->red_buf.p;
  out.assign((size_t)nv, 0);
#ifndef ABM_EMU
  if (h->p2p_on) {
    p2p_stats_send(h, d_pending);
    if ((rc = p2p_stats_recv(h, nullptr, 0))) return rc;
    if (with_state_exchange && (rc = strip_exchange_states(h))) return rc;
    ABM_CK(h, copy_d2h(out.data(), d, (size_t)nv * 8, h->stream));
    return 0;
  }
#endif
  strip_stats_kernel(h, d_pending);
#ifndef ABM_EMU
  if (h->nccl_comm) {
    if (with_state_exchange) { if ((rc = strip_exchange_states(h, d, nv))) return rc; } /* same NCCL group as the decision bytes */
    else ABM_NCCL(h, g_nccl.AllReduce(d, d, (size_t)nv, ncclUint64, ncclSum, (ncclComm_t)h->nccl_comm, h->stream));
    ABM_CK(h, copy_d2h(out.data(), d, (size_t)nv * 8, h->stream));
    return 0;
  }
#endif
  if (with_state_exchange && (rc = strip_exchange_states(h))) return rc;
  ABM_CK(h, copy_d2h(out.data(), d, (size_t)nv * 8, h->stream));
  if ((rc = comm_allreduce_u64(h, out.data(), nv))) return rc;
  return 0;
}

/* offspring ids across strips: all-gather of [count, parent ids ...] and a rank kernel; returns the total through a
 * device counter read with the commit counters */
static int strip_birth_ids(abm_handle* h, uint64_t bbcap, int nxt, uint32_t* d_total) {
  int rc;
  const int P = h->n_ranks;
  const uint64_t words = 1 + bbcap;
  if ((rc = ensure(h, h->birth_ids, words * 4 * (size_t)(P + 2) + 64))) return rc;
  uint32_t* msg = (uint32_t*)h->birth_ids.p;
  uint32_t* all = msg + words;
  uint32_t* ctr = (uint32_t*)h->counters.p;
  bool packed_remotely = false;
#ifndef ABM_EMU
  if (h->p2p_on && P > 1) {
    if ((rc = p2p_allgather_births(h, all, bbcap))) return rc;
    packed_remotely = true;
  }
#endif
  if (!packed_remotely) {
    BirthPackBody bp;
    bp.birth_parent = (const uint32_t*)h->birth_parent.p; bp.birth_count = ctr + CTR_BIRTHS; bp.cap = (uint32_t)bbcap; bp.msg = msg;
    launch_foreach(h->stream, bp, bbcap ? bbcap : 1);
  }
  if (packed_remotely) {
  } else if (P == 1) {
    ABM_CK(h, copy_d2d(all, msg, words * 4, h->stream));
  } else {
    bool done = false;
#ifndef ABM_EMU
    if (h->nccl_comm) {
      ABM_NCCL(h, g_nccl.AllGather(msg, all, (size_t)words, ncclUint32, (ncclComm_t)h->nccl_comm, h->stream));
      done = true;
    }
#endif
    if (!done) {
      std::vector<uint32_t> hs((size_t)words), hr((size_t)words * P);
      ABM_CK(h, copy_d2h(hs.data(), msg, words * 4, h->stream));
      if ((rc = comm_allgather_u32(h, hs.data(), hr.data(), (uint32_t)words))) return rc;
      ABM_CK(h, copy_h2d(all, hr.data(), words * 4 * (size_t)P, h->stream));
    }
  }
  uint32_t* ranks = all + words * (size_t)P; /* zeroed scratch, one counter per local birth */
  BirthAssignBody ba;
  ba.all = all; ba.n_ranks = P; ba.cap = (uint32_t)bbcap; ba.next_id = (uint32_t)h->next_id;
  ba.birth_pos = (const uint32_t*)h->birth_pos.p; ba.birth_parent = (const uint32_t*)h->birth_parent.p; ba.birth_count = ctr + CTR_BIRTHS;
  ba.rank_in = ranks; ba.birth_bits = nullptr; ba.birth_prefix = nullptr;
  ba.out_id = (uint32_t*)h->id[nxt].p; ba.total_out = d_total;
  if ((uint64_t)P * bbcap > h->birth_bitmap_from) { /* many offspring: bitmap over the id space + popcount prefix */
    const uint64_t idw = bitmap_words(key_space(h, h->next_id)); /* birth_parent holds order keys */
    if ((rc = ensure(h, h->bits, (idw + 4) * 4))) return rc;
    ProfScope ps(h->prof, h->stream, "birth_ids", bbcap, 2);
    ABM_CK(h, dev_memset(h->bits.p, 0, idw * 4, h->stream));
    BirthMarkBody bm;
    bm.all = all; bm.n_ranks = P; bm.cap = (uint32_t)bbcap; bm.bits = (uint32_t*)h->bits.p;
    launch_foreach(h->stream, bm, (uint64_t)P * bbcap);
    if ((rc = rank_prefix(h, idw))) return rc;
    ba.birth_bits = (const uint32_t*)h->bits.p; ba.birth_prefix = (const uint32_t*)h->bits_prefix.p;
    launch_foreach(h->stream, ba, bbcap ? bbcap : 1);
  } else {
    ABM_CK(h, dev_memset(ranks, 0, (bbcap + 1) * 4, h->stream));
    BirthRankBody br;
    br.all = all; br.n_ranks = P; br.cap = (uint32_t)bbcap;
    br.birth_parent = (const uint32_t*)h->birth_parent.p; br.birth_count = ctr + CTR_BIRTHS; br.rank_out = ranks;
    ProfScope ps(h->prof, h->stream, "birth_ids", bbcap, 2);
    launch_foreach(h->stream, br, (bbcap ? bbcap : 1) * 32);
    launch_foreach(h->stream, ba, bbcap ? bbcap : 1);
  }
  return check_launch(h, "birth ids");
}

