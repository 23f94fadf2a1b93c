/*
 * abm_host_routes.inl — host side of the batched A*: walkmap components, plan_batch, ALTERNATED_WALK proposals
 * Part of abm_engine.cu (one translation unit: included there, inside its anonymous namespace where it says so);
 * not a stand-alone header.
 */
/* component labels of the walkmap, computed once per abm_set_grid, on the first plan */
static int ensure_components(abm_handle* h) {
  int rc;
  if (h->comp_valid) return 0;
  const uint64_t gt = h->g.gt;
  if ((rc = ensure(h, h->as_comp, gt * 4))) return rc;
  uint32_t* ctr = (uint32_t*)h->counters.p;
  ProfScope ps(h->prof, h->stream, "walkmap_components", gt);
  CompInitBody ci;
  ci.walkbits = (const uint32_t*)h->walkbits.p; ci.comp = (uint32_t*)h->as_comp.p;
  launch_foreach(h->stream, ci, gt);
  for (int it = 0; it < 64; ++it) {
    ABM_CK(h, dev_memset(ctr + CTR_NEXT_REQ, 0, 4, h->stream));
    CompHookBody hk;
    hk.g = h->g; hk.walkbits = ci.walkbits; hk.comp = ci.comp; hk.changed = ctr + CTR_NEXT_REQ;
    launch_foreach(h->stream, hk, gt);
    CompCompressBody cc;
    cc.comp = ci.comp;
    launch_foreach(h->stream, cc, gt);
    if ((rc = check_launch(h, "walkmap components"))) return rc;
    if ((rc = read_counters(h))) return rc;
    if (!h->h_counters[CTR_NEXT_REQ]) { h->comp_valid = true; return 0; }
  }
  ABM_FAIL(h, ABM_E_INTERNAL, "component labelling did not converge");
}

/* plan_route! for B requests held in device arrays (agent id, start tcell, goal tcell): ONE launch of `slots` warps that
 * take requests from a device-side queue, search, and write their paths into the route pool themselves.  The host only
 * reads the status words back; requests that found the pool full are re-run after growing it. */
static int plan_batch(abm_handle* h, uint64_t B, const uint32_t* d_id, const uint32_t* d_src, const uint32_t* d_dst) {
  int rc;
  if (B == 0) return 0;
  if (h->g.nx > 65534 || h->g.ny > 65534) ABM_FAIL(h, ABM_E_BADARG, "A* needs grid dims below 65535");
  if (h->cfg.periodic && (h->cfg.nx < 3 || h->cfg.ny < 3)) ABM_FAIL(h, ABM_E_BADARG, "A* on a periodic grid needs dims >= 3 (two neighbour offsets wrap onto one cell)");
  if (B > 0xFFFFFFF0ull) ABM_FAIL(h, ABM_E_BADARG, "too many route requests in one batch");
  const uint64_t gt = h->g.gt;
  /* Open-list capacity per slot: a search that pushes every neighbour of every cell holds at most 8 * cells entries, and never
   * more than 2^20 are allowed; searches on real terrain hold a few thousand, so slots start with 2^18 entries (2 MB instead of
   * 8: at 2048 x 2048 that is 27 instead of 33 MB per slot, a quarter more searches in flight) and the rare search that
   * overflows is run again after the heaps grew fourfold.  ABM_ASTAR_HEAP0 (tests) sets the first capacity. */
  const uint64_t heap_max = std::min<uint64_t>(8ull * (uint64_t)h->g.nx * h->g.ny + 16, 1ull << 20);
  if (h->as_heap_want == 0 || h->as_heap_want > heap_max) {
    h->as_heap_want = std::min<uint64_t>(heap_max, 1ull << 18);
    if (const char* hs = getenv("ABM_ASTAR_HEAP0")) h->as_heap_want = std::min<uint64_t>(heap_max, (uint64_t)std::max(atoll(hs), 2ll));
  }
  uint64_t max_resident = 148ull * 32; /* one warp-sized CTA per slot, at most 32 CTAs per SM */
  if (const char* ms = getenv("ABM_ASTAR_SLOTS")) max_resident = (uint64_t)std::max(atoll(ms), 1ll);
  /* (re)allocates the slots for the wanted heap capacity; a larger heap may mean fewer slots */
  auto size_slots = [&]() -> int {
    const uint64_t heap_cap = h->as_heap_want;
    if (h->as_slots != 0 && h->as_heap_cap == heap_cap && h->as_slots >= std::min<uint64_t>(B, max_resident)) return 0;
    const uint64_t per_slot = gt * 6 + heap_cap * 8;
    /* searches are latency-bound: slots in flight are throughput.  The arrays already held count as free. */
    const uint64_t held = h->as_g.bytes + h->as_meta.bytes + h->as_heap.bytes;
    const uint64_t budget = std::min<uint64_t>((device_free_bytes() + held) / 4 * 3, h->astar_budget);
    const uint64_t slots = std::max<uint64_t>(1, std::min<uint64_t>(std::min<uint64_t>(B, max_resident), budget / per_slot));
    if (slots > h->as_slots || h->as_heap_cap != heap_cap) {
      /* exact sizes (ensure() adds a quarter of head room, which is tens of GB here), the old array released first */
      auto exact = [&](DevBuf& b, uint64_t bytes) -> int {
        bytes = (bytes + 255) & ~(uint64_t)255;
        if (b.p && b.bytes >= bytes) return 0;
        dev_free(b.p);
        b.p = nullptr; b.bytes = 0;
        if (dev_malloc(&b.p, bytes) != 0) { b.p = nullptr; ABM_FAIL(h, ABM_E_NOMEM, "device allocation of %llu bytes failed", (unsigned long long)bytes); }
        b.bytes = bytes;
        return 0;
      };
      int rc2;
      if ((rc2 = exact(h->as_g, slots * gt * 4))) return rc2;
      if ((rc2 = exact(h->as_meta, slots * gt * 2))) return rc2;
      if ((rc2 = exact(h->as_heap, slots * heap_cap * 8))) return rc2;
      if ((rc2 = ensure(h, h->as_slot_gen, slots * 4))) return rc2;
      ABM_CK(h, dev_memset(h->as_meta.p, 0, slots * gt * 2, h->stream));
      ABM_CK(h, dev_memset(h->as_slot_gen.p, 0, slots * 4, h->stream));
      h->as_slots = slots; h->as_heap_cap = heap_cap;
    }
    return 0;
  };
  if ((rc = size_slots())) return rc;
  uint64_t S = std::min<uint64_t>(std::min<uint64_t>(h->as_slots, B), max_resident);
  if ((rc = ensure_components(h))) return rc;
  if ((rc = ensure(h, h->as_status, B * 4))) return rc;
  if ((rc = ensure(h, h->as_len, B * 4))) return rc;
  if ((rc = ensure(h, h->as_cost, B * 8))) return rc;
  if ((rc = ensure(h, h->as_off, B * 8))) return rc;
  uint32_t* ctr = (uint32_t*)h->counters.p;
  unsigned long long* d_exp = (unsigned long long*)(ctr + CTR_EXP_LO);
  unsigned long long* d_cursor = (unsigned long long*)(ctr + CTR_POOL_LO);
  std::vector<uint32_t> st(B), todo;
  /* room for the new paths: a guess from the grid size; a request that does not fit reports AS_POOL_FULL and runs again */
  uint64_t pool_cap = std::max<uint64_t>(h->rt_pool.bytes / 4, 1);
  {
    uint64_t room = std::min<uint64_t>(B * (uint64_t)std::max(h->g.nx, h->g.ny), 1ull << 28);
    if (const char* pg = getenv("ABM_ROUTE_POOL_GUESS")) room = (uint64_t)std::max(atoll(pg), 0ll); /* tests: force the overflow path */
    const uint64_t guess = h->rt_pool_used + room + 1;
    if (guess > pool_cap) { if ((rc = grow_keep(h, h->rt_pool, guess * 4))) return rc; pool_cap = h->rt_pool.bytes / 4; }
  }
  uint64_t n_todo = B;
  if (B > S) { /* more requests than slots: far goals first, so that the longest searches do not start last */
    AStarHintBody hb;
    hb.a = astar_grid(h); hb.src = d_src; hb.dst = d_dst; hb.hint = (uint32_t*)h->as_status.p;
    launch_foreach(h->stream, hb, B);
    if ((rc = check_launch(h, "astar hints"))) return rc;
    ABM_CK(h, copy_d2h(st.data(), h->as_status.p, B * 4, h->stream));
    ABM_CK(h, stream_sync(h->stream));
    todo.resize(B);
    for (uint64_t i = 0; i < B; ++i) todo[i] = (uint32_t)i;
    std::stable_sort(todo.begin(), todo.end(), [&](uint32_t x, uint32_t y) { return st[x] > st[y]; });
  }
  for (int attempt = 0; n_todo > 0; ++attempt) {
    if (attempt > 40) ABM_FAIL(h, ABM_E_INTERNAL, "route pool kept overflowing");
    unsigned long long cursor = h->rt_pool_used;
    ABM_CK(h, copy_h2d(d_cursor, &cursor, 8, h->stream));
    ABM_CK(h, dev_memset(ctr + CTR_NEXT_REQ, 0, 4, h->stream));
    if (!todo.empty()) {
      if ((rc = ensure(h, h->as_todo, todo.size() * 4))) return rc;
      ABM_CK(h, copy_h2d(h->as_todo.p, todo.data(), todo.size() * 4, h->stream));
    }
    AStarWarpCta sb;
    sb.a = astar_grid(h); sb.src = d_src; sb.dst = d_dst; sb.comp = (const uint32_t*)h->as_comp.p;
    sb.todo = todo.empty() ? nullptr : (const uint32_t*)h->as_todo.p; sb.n_todo = (uint32_t)n_todo; sb.next_todo = ctr + CTR_NEXT_REQ;
    sb.gbuf = (uint32_t*)h->as_g.p; sb.meta = (uint16_t*)h->as_meta.p; sb.heap = (uint64_t*)h->as_heap.p;
    sb.heap_cap = (uint32_t)h->as_heap_cap; sb.slot_gen = (uint32_t*)h->as_slot_gen.p;
    sb.out_status = (uint32_t*)h->as_status.p; sb.out_len = (uint32_t*)h->as_len.p; sb.out_cost = (int64_t*)h->as_cost.p;
    sb.out_off = (uint64_t*)h->as_off.p;
    sb.pool = (uint32_t*)h->rt_pool.p; sb.pool_cursor = d_cursor; sb.pool_cap = pool_cap;
    sb.expansions = d_exp;
    const uint64_t grid = std::min<uint64_t>(S, n_todo);
    { ProfScope ps(h->prof, h->stream, "astar_search", n_todo); launch_cta(h->stream, sb, grid, AS_ARITY, AStarWarpCta::smem_bytes()); }
    if ((rc = check_launch(h, "astar search"))) return rc;
    ABM_CK(h, copy_d2h(st.data(), h->as_status.p, B * 4, h->stream));
    ABM_CK(h, copy_d2h(&cursor, d_cursor, 8, h->stream));
    if ((rc = read_counters(h))) return rc; /* synchronises the stream */
    std::vector<uint32_t> again;
    bool heap_full = false, pool_full = false;
    auto look = [&](uint32_t i) {
      if (st[i] == AS_POOL_FULL) { pool_full = true; again.push_back(i); }
      else if (st[i] == AS_HEAP_FULL) { heap_full = true; again.push_back(i); }
    };
    if (todo.empty()) { for (uint64_t i = 0; i < B; ++i) look((uint32_t)i); }
    else for (uint32_t i : todo) look(i);
    if (again.empty()) { h->rt_pool_used = cursor; break; }
    /* everything that fitted stays (and stays counted); the rest runs again with more room */
    h->rt_pool_used = std::min<uint64_t>(cursor, pool_cap);
    if (heap_full) {
      if (h->as_heap_cap >= heap_max) ABM_FAIL(h, ABM_E_CAPACITY, "A* open list exceeded %llu entries", (unsigned long long)heap_max);
      h->as_heap_want = std::min<uint64_t>(heap_max, h->as_heap_cap * 4);
      if ((rc = size_slots())) return rc;
      S = std::min<uint64_t>(std::min<uint64_t>(h->as_slots, B), max_resident);
    }
    if (pool_full) {
      if ((rc = grow_keep(h, h->rt_pool, (2 * pool_cap + (cursor - std::min<uint64_t>(cursor, pool_cap)) + 1) * 4))) return rc;
      pool_cap = h->rt_pool.bytes / 4;
    }
    todo.swap(again);
    n_todo = todo.size();
  }
  const uint64_t used = h->h_counters[CTR_ROUTE_SLOTS];
  if (used + B > h->rt_slots_cap) { /* at most one new slot per request */
    const uint64_t cap = (used + B) * 2 + 64;
    if ((rc = grow_keep(h, h->rt_off, cap * 8))) return rc;
    if ((rc = grow_keep(h, h->rt_len, cap * 4))) return rc;
    if ((rc = grow_keep(h, h->rt_pos, cap * 4))) return rc;
    if ((rc = grow_keep(h, h->rt_cost, cap * 8))) return rc;
    h->rt_slots_cap = cap;
  }
  RouteBindBody bb;
  bb.rs = route_store(h); bb.r_cost = (int64_t*)h->rt_cost.p; bb.agent_id = d_id; bb.status = (const uint32_t*)h->as_status.p;
  bb.len = (const uint32_t*)h->as_len.p; bb.cost = (const int64_t*)h->as_cost.p; bb.pool_off = (const uint64_t*)h->as_off.p;
  bb.slot_counter = ctr + CTR_ROUTE_SLOTS;
  { ProfScope ps(h->prof, h->stream, "route_bind", B); launch_foreach(h->stream, bb, B); }
  if ((rc = check_launch(h, "route bind"))) return rc;
  unsigned long long e = 0;
  ABM_CK(h, copy_d2h(&e, d_exp, 8, h->stream));
  h->expansions = (double)e;
  return 0;
}

/* ALTERNATED_WALK: ranks, behaviour segments, route planning of the switching agents, proposals */
static int alternated_propose(abm_handle* h, View& v, uint32_t* ctr) {
  int rc;
  const uint64_t n = h->n;
  const uint32_t C = (uint32_t)h->cfg.counter;
  const uint64_t words = bitmap_words(key_space(h, h->next_id)); /* ranks in schedule order: a bitmap over the order keys */
  if ((rc = ensure(h, h->bits, (words + 4) * 4))) return rc;
  if ((rc = ensure(h, h->rank, n * 4))) return rc;
  if ((rc = ensure(h, h->by_rank, n * 4))) return rc;
  if ((rc = ensure(h, h->remote_next, n * 4))) return rc;
  if ((rc = ensure(h, h->remote_dest, n * 4))) return rc;
  if ((rc = ensure_route_ids(h, h->next_id))) return rc;
  v = make_view(h);
  ABM_CK(h, dev_memset(h->bits.p, 0, words * 4, h->stream));
  SetBitsBody sbb;
  sbb.id = v.okey; sbb.bits = (uint32_t*)h->bits.p;
  { ProfScope ps(h->prof, h->stream, "rank_by_id", n, 4); launch_foreach(h->stream, sbb, n);
    if ((rc = rank_prefix(h, words))) return rc;
    RankBody rb;
    rb.id = v.okey; rb.alive_bits = (const uint32_t*)h->bits.p; rb.prefix = (const uint32_t*)h->bits_prefix.p;
    rb.rank = (uint32_t*)h->rank.p; rb.by_rank = (uint32_t*)h->by_rank.p;
    launch_foreach(h->stream, rb, n); }
  /* counter value seen by rank r: ((c - 1 - r) mod C) + 1 for c in 1..C; a switch happens where it equals C */
  const int64_t c = h->behav_counter;
  const int inert = c == 0 ? 1 : 0;
  const uint64_t first = inert ? 1 : (uint64_t)(c % C);
  const uint64_t nsw = first < n ? (n - first + C - 1) / C : 0;
  if ((rc = ensure(h, h->seg_behav, nsw + 1))) return rc;
  if ((rc = ensure(h, h->plan_agent, (nsw + 1) * 4))) return rc;
  if ((rc = ensure(h, h->plan_goal, (nsw + 1) * 4))) return rc;
  ABM_CK(h, dev_memset(ctr + CTR_PLANS, 0, 4, h->stream));
  uint64_t plans = 0;
  if (nsw > 0) {
    AltSwitchBody sw;
    sw.v = v; sw.by_rank = (const uint32_t*)h->by_rank.p; sw.first = (uint32_t)first; sw.counter = C;
    sw.seg_behav = (uint8_t*)h->seg_behav.p; sw.plan_agent = (uint32_t*)h->plan_agent.p; sw.plan_goal = (uint32_t*)h->plan_goal.p;
    sw.plan_count = ctr + CTR_PLANS;
    { ProfScope ps(h->prof, h->stream, "alt_switch", nsw); launch_foreach(h->stream, sw, nsw); }
    if ((rc = check_launch(h, "behaviour switch"))) return rc;
    if ((rc = read_counters(h))) return rc;
    plans = h->h_counters[CTR_PLANS];
  }
  if (plans > 0) { /* plan_route!(agent, random_walkable(...), pathfinder) from the agents' tick-start cells */
    if ((rc = ensure(h, h->as_src, plans * 4))) return rc;
    if ((rc = ensure(h, h->as_id, plans * 4))) return rc;
    GatherPlanBody gp;
    gp.cell = v.cell; gp.id = v.id; gp.plan_agent = (const uint32_t*)h->plan_agent.p;
    gp.src = (uint32_t*)h->as_src.p; gp.aid = (uint32_t*)h->as_id.p;
    launch_foreach(h->stream, gp, plans);
    if ((rc = plan_batch(h, plans, (const uint32_t*)h->as_id.p, (const uint32_t*)h->as_src.p, (const uint32_t*)h->plan_goal.p))) return rc;
  }
  AltProposeBody ap;
  ap.v = v;
  ap.rt.slot_of_id = (uint32_t*)h->rt_slot_of_id.p; ap.rt.r_off = (const uint64_t*)h->rt_off.p; ap.rt.r_len = (const uint32_t*)h->rt_len.p;
  ap.rt.r_pos = (uint32_t*)h->rt_pos.p; ap.rt.pool = (const uint32_t*)h->rt_pool.p;
  ap.rank = (const uint32_t*)h->rank.p; ap.seg_behav = (const uint8_t*)h->seg_behav.p;
  ap.first = (uint32_t)first; ap.counter = C; ap.prev_behav = (int32_t)h->behav; ap.inert_rank0 = inert;
  ap.pending = (uint32_t*)h->pend[0].p; ap.pending_count = ctr + CTR_PEND_A;
  { ProfScope ps(h->prof, h->stream, "propose", n); launch_foreach(h->stream, ap, n); }
  if ((rc = check_launch(h, "alternated propose"))) return rc;
  /* model.behav[1] / model.behav_counter[1] after this tick's n agent-steps */
  if (nsw > 0) {
    uint8_t last = 0;
    ABM_CK(h, copy_d2h(&last, (uint8_t*)h->seg_behav.p + (nsw - 1), 1, h->stream));
    h->behav = last;
  }
  if (n > 0) {
    const int64_t c1 = inert ? (int64_t)C : c;
    const int64_t steps = inert ? (int64_t)n - 1 : (int64_t)n;
    int64_t m = (c1 - 1 - steps) % (int64_t)C;
    if (m < 0) m += C;
    h->behav_counter = m + 1;
  }
  return 0;
}

