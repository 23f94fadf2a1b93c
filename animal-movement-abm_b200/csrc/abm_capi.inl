/*
 * abm_capi.inl — the C ABI of include/abm.h: argument checks, host<->device conversion, calls into the tick functions
 * Part of abm_engine.cu (one translation unit: included there, inside its anonymous namespace where it says so);
 * not a stand-alone header.
 */
extern "C" {

void ABM_API(config_default)(abm_config* c) {
  memset(c, 0, sizeof(*c));
  c->struct_size = (int32_t)sizeof(abm_config);
  c->nx = 80; c->ny = 80; c->periodic = 0; c->walk_type = ABM_RANDOM_WALK; c->eat = 1; c->reproduce = 1;
  c->walkmap_regrowth = 0; c->regrowth_time = 30; c->visual_distance = 5; c->counter = 50;
  c->randomwalk_ifempty = 1; c->astar_metric = ABM_ASTAR_DIRECT; c->astar_penalty = 0; c->device = 0;
  c->prob_random_walk = 0.9; c->delta_energy = 4; c->movement_cost = 1; c->reproduction_prob = 0.001;
  c->attractor_x = 40; c->attractor_y = 40; c->seed = 321;
}

int ABM_API(config_size)(void) { return (int)sizeof(abm_config); }

const char* ABM_API(last_error)(abm_handle* h) { return h ? h->err.c_str() : g_create_error.c_str(); }

void ABM_API(destroy)(abm_handle* h) {
  if (!h) return;
#ifndef ABM_EMU
  cudaSetDevice(h->device);
  if (h->stream) { cudaStreamSynchronize(h->stream); }
  for (int r = 0; r < P2P_MAXP; ++r) if (h->p2p_peer[r] && h->p2p_peer[r] != h->p2p_mine) cudaIpcCloseMemHandle(h->p2p_peer[r]);
  if (h->p2p_mine) cudaFree(h->p2p_mine);
  if (h->nccl_comm && g_nccl.CommDestroy) g_nccl.CommDestroy((ncclComm_t)h->nccl_comm);
  side_stream_destroy(h->side);
  event_destroy(h->ev_proposed); event_destroy(h->ev_ranked); event_destroy(h->ev_rows); event_destroy(h->ev_halo);
  if (h->stream) { cudaStreamDestroy(h->stream); }
  if (h->h_counters) cudaFreeHost(h->h_counters);
#else
  free(h->h_counters);
#endif
  h->tm_total.destroy();
  h->prof.destroy();
  delete h;
}

int ABM_API(create)(const abm_config* cfg, abm_handle** out) {
  if (!cfg || !out) { g_create_error = "null argument"; return ABM_E_BADARG; }
  *out = nullptr;
  int rc = validate_config(cfg, g_create_error);
  if (rc) return rc;
  abm_handle* h = new (std::nothrow) abm_handle();
  if (!h) { g_create_error = "out of host memory"; return ABM_E_NOMEM; }
  h->cfg = *cfg;
  h->device = cfg->device;
#ifndef ABM_EMU
  {
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev <= cfg->device || cfg->device < 0) {
      g_create_error = std::string("no usable CUDA device (libabm has no CPU fallback): ") + cudaGetErrorString(e);
      delete h; return ABM_E_CUDA;
    }
    cudaDeviceProp prop;
    e = cudaSetDevice(cfg->device);
    if (e == cudaSuccess) e = cudaGetDeviceProperties(&prop, cfg->device);
    if (e != cudaSuccess || prop.major < 10) {
      g_create_error = "libabm kernels are built for sm_100a only; device is not a Blackwell-class GPU";
      delete h; return ABM_E_CUDA;
    }
    if (cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking) != cudaSuccess || side_stream_create(&h->side) != cudaSuccess ||
        event_create(&h->ev_proposed) != cudaSuccess || event_create(&h->ev_ranked) != cudaSuccess ||
        event_create(&h->ev_rows) != cudaSuccess || event_create(&h->ev_halo) != cudaSuccess ||
        cudaMallocHost((void**)&h->h_counters, 2 * CTR_WORDS * sizeof(uint32_t)) != cudaSuccess) {
      g_create_error = "stream / pinned memory creation failed";
      delete h; return ABM_E_CUDA;
    }
  }
#else
  h->h_counters = (uint32_t*)calloc(2 * CTR_WORDS, sizeof(uint32_t));
#endif
  h->tm_total.init();
  h->strip = cfg->n_ranks > 1 || cfg->strip_ny > 0;
  h->n_ranks = cfg->n_ranks > 1 ? cfg->n_ranks : 1;
  h->my_rank = h->strip ? cfg->rank : 0;
  Geo& g = h->g;
  g.nx = cfg->nx; g.ny = h->strip ? cfg->strip_ny + 2 * CORE : cfg->ny;
  g.tiles_x = (g.nx + 31) / 32; g.tiles_y = (g.ny + 31) / 32;
  g.gt = (uint32_t)((uint64_t)g.tiles_x * g.tiles_y * 1024);
  g.periodic = cfg->periodic ? 1 : 0;
  g.py = g.periodic; g.y_lo = 0; g.y_hi = g.ny; g.y_off = 0; g.own_lo = 0; g.own_hi = g.ny;
  if (h->strip) { /* stored rows: [ghost tile row][own rows][ghost tile row]; the ring exchange carries the y wrap */
    g.py = 0; g.own_lo = CORE; g.own_hi = CORE + cfg->strip_ny; g.y_off = cfg->strip_y0 - CORE;
    const bool first = cfg->strip_y0 == 0, last = cfg->strip_y0 + cfg->strip_ny == cfg->ny;
    g.y_lo = (!g.periodic && first) ? g.own_lo : 0;
    g.y_hi = (!g.periodic && last) ? g.own_hi : g.ny;
    h->has_prev = g.periodic || !first;
    h->has_next = g.periodic || !last;
  }
  Params& p = h->params;
  memset(&p, 0, sizeof p);
  p.walk_type = cfg->walk_type; p.eat = cfg->eat; p.reproduce = cfg->reproduce; p.ifempty = cfg->randomwalk_ifempty;
  p.walkmap_regrowth = cfg->walkmap_regrowth; p.regrowth_time = cfg->regrowth_time; p.r = cfg->visual_distance; p.counter = cfg->counter;
  p.prob_random_walk = cfg->prob_random_walk; p.delta_energy = cfg->delta_energy; p.movement_cost = cfg->movement_cost;
  p.reproduction_prob = cfg->reproduction_prob;
  p.ax2 = (int64_t)llround(2 * cfg->attractor_x); p.ay2 = (int64_t)llround(2 * cfg->attractor_y);
  p.seed = cfg->seed;

  /* tile path on grids made of whole 32x32 cores that are wider than one window; everything else (and the model-global
   * counter of ALTERNATED_WALK) takes the general path.  cfg->reserved0 = 1 forces the general path (tests). */
  h->tile_mode = h->strip || (cfg->reserved0 != 1 && cfg->walk_type != ABM_ALTERNATED_WALK && cfg->nx % CORE == 0 && cfg->ny % CORE == 0 &&
                             cfg->nx >= 2 * CORE && cfg->ny >= 2 * CORE);
  h->nsx = g.nx / SUB; h->nsy = g.ny / SUB;
  h->stream_mode = h->tile_mode && stream_rule(*cfg) && h->n_ranks <= STREAM_MAX_RANKS;
  if (const char* sm = getenv("ABM_STREAM")) { if (atoi(sm) == 0) h->stream_mode = false; }
  if (const char* sq = getenv("ABM_STREAM_CAP")) { const int q = atoi(sq); if (q >= 1 && q <= (1 << 20)) h->s_cap_forced = (uint32_t)q; }
  if (const char* mg = getenv("ABM_MIGRANTS")) h->use_migrants = atoi(mg) != 0;
  if (const char* ag = getenv("ABM_ASTAR_GB")) { const int q = atoi(ag); if (q >= 1 && q <= 160) h->astar_budget = (uint64_t)q << 30; }
  if (const char* bb = getenv("ABM_BIRTH_BITMAP")) h->birth_bitmap_from = (uint64_t)std::max(atoll(bb), 0ll);
  if (const char* op = getenv("ABM_OPTIMISTIC")) h->optimistic = atoi(op) != 0;
  if (const char* ar = getenv("ABM_AGENT_ROUNDS")) { const int q = atoi(ar); if (q >= 0 && q <= 16) h->fixed_agent_rounds = h->min_agent_rounds = q; }
  if (const char* tm = getenv("ABM_TILE_MARGIN")) { const int q = atoi(tm); if (q >= 2 && q <= HALO) h->tile_margin = q; }
  if (const char* eb = getenv("ABM_EARLY_BIRTHS")) h->early_births = atoi(eb) != 0;
  if (const char* oh = getenv("ABM_OVERLAP_HALO")) h->overlap_halo = atoi(oh) != 0;
  if (const char* ol = getenv("ABM_OVERLAP_LISTS")) h->overlap_lists = atoi(ol) != 0;
  if (const char* tr = getenv("ABM_TILE_ROUNDS")) { const int q = atoi(tr); if (q >= 0 && q <= 64) h->tile_rounds = q; }
  rc = 0;
  do {
    if ((rc = ensure(h, h->counters, CTR_WORDS * 4))) break;
    if ((rc = ensure(h, h->grass, g.gt))) break;
    if ((rc = ensure(h, h->walkbits, (size_t)g.gt / 8 + 16))) break;
    if (h->tile_mode) {
      const size_t nsub = (size_t)h->nsx * h->nsy;
      for (int q = 0; q < 2; ++q) {
        if ((rc = ensure(h, h->sub_off[q], (nsub + 8) * 4))) break; /* dev_malloc zero-fills: empty ghost rows stay empty */
        if ((rc = ensure(h, h->sub_cnt[q], (nsub + 8) * 4))) break;
      }
      if (rc) break;
    } else {
      if ((rc = ensure(h, h->cell_start[0], ((size_t)g.gt + 8) * 4))) break;
      if ((rc = ensure(h, h->cell_start[1], ((size_t)g.gt + 8) * 4))) break;
    }
    if ((rc = ensure(h, h->scan_tmp, scan_tmp_words((uint64_t)g.gt + 1) * 4))) break;
    if ((rc = ensure_soa(h, 0, 1))) break;
    if ((rc = ensure_soa(h, 1, 1))) break;
    if (cfg->walk_type == ABM_ALTERNATED_WALK) {
      if ((rc = ensure(h, h->remote_head, (size_t)g.gt * 4))) break;
      if (dev_memset(h->remote_head.p, 0xFF, (size_t)g.gt * 4, h->stream) != 0) { rc = ABM_E_CUDA; break; }
    }
    if ((rc = ensure(h, h->elev, (size_t)g.gt * 2))) break;
    /* default tables: exp(-sqrt(k)/3) from the C library; closest-neighbour visiting order */
    const int r = cfg->walk_type == ABM_RANDOM_WALK_GREGARIOUS ? cfg->visual_distance : 1;
    const int nl = 2 * (r + 1) * (r + 1) + 1;
    std::vector<double> lut(nl);
    for (int k = 0; k < nl; ++k) lut[k] = exp(-sqrt((double)k) / 3.0);
    if ((rc = ensure(h, h->lut, nl * 8))) break;
    if (copy_h2d(h->lut.p, lut.data(), nl * 8, h->stream) != 0) { rc = ABM_E_CUDA; break; }
    const int side = 2 * r + 1;
    std::vector<std::pair<int, int>> order;
    for (int by = -r; by <= r; ++by) for (int bx = -r; bx <= r; ++bx) order.push_back(std::make_pair(bx * bx + by * by, (by + r) * side + (bx + r)));
    std::sort(order.begin(), order.end());
    std::vector<uint8_t> ord(order.size());
    for (size_t i = 0; i < order.size(); ++i) ord[i] = (uint8_t)order[i].second;
    if ((rc = ensure(h, h->greg_order, ord.size()))) break;
    if (copy_h2d(h->greg_order.p, ord.data(), ord.size(), h->stream) != 0) { rc = ABM_E_CUDA; break; }
    std::vector<int8_t> dxy(2 * ord.size());
    for (size_t i = 0; i < ord.size(); ++i) { dxy[2 * i] = (int8_t)(ord[i] % side - r); dxy[2 * i + 1] = (int8_t)(ord[i] / side - r); }
    if ((rc = ensure(h, h->greg_dxy, dxy.size()))) break;
    if (copy_h2d(h->greg_dxy.p, dxy.data(), dxy.size(), h->stream) != 0) { rc = ABM_E_CUDA; break; }
    if ((rc = build_greg_table(h))) break;
  } while (0);
  if (rc) { g_create_error = h->err.empty() ? "device allocation failed" : h->err; ABM_API(destroy)(h); return rc; }
  *out = h;
  return ABM_OK;
}

int ABM_API(set_weight_lut)(abm_handle* h, const double* w, int32_t n) {
  if (!h || !w) return ABM_E_BADARG;
  int rc = set_device(h);
  if (rc) return rc;
  const int r = h->cfg.walk_type == ABM_RANDOM_WALK_GREGARIOUS ? h->cfg.visual_distance : 1;
  const int nl = 2 * (r + 1) * (r + 1) + 1;
  if (n < nl) ABM_FAIL(h, ABM_E_BADARG, "weight table needs %d entries (k = 0..2(r+1)^2), got %d", nl, n);
  ABM_CK(h, copy_h2d(h->lut.p, w, (size_t)nl * 8, h->stream));
  return build_greg_table(h);
}

int ABM_API(set_grid)(abm_handle* h, const uint8_t* fg, const int64_t* cd, const uint8_t* wm, const int64_t* el) {
  if (!h) return ABM_E_BADARG;
  if ((fg == nullptr) != (cd == nullptr)) ABM_FAIL(h, ABM_E_BADARG, "fully_grown and countdown come together (both NULL: terrain only, the grass starts bare with countdown 0)");
  int rc = set_device(h);
  if (rc) return rc;
  const Geo& g = h->g;
  if ((rc = stream_invalidate(h))) return rc; /* the filed steps were drawn on the old walkmap */
  ABM_CK(h, dev_memset(h->grass.p, GRASS_PAD, g.gt, h->stream));
  ABM_CK(h, dev_memset(h->walkbits.p, 0, (size_t)g.gt / 8, h->stream));
  h->comp_valid = false;
  if ((rc = ensure(h, h->elev, (size_t)g.gt * 2))) return rc;
  ABM_CK(h, dev_memset(h->elev.p, 0, (size_t)g.gt * 2, h->stream));
  ABM_CK(h, dev_memset((uint32_t*)h->counters.p + CTR_ERR, 0, 4, h->stream));
  const uint64_t rows_per = std::max<uint64_t>(1, (16ull << 20) / (uint64_t)g.nx);
  for (uint64_t row0 = 0; row0 < (uint64_t)g.ny; row0 += rows_per) {
    const uint64_t rows = std::min<uint64_t>(rows_per, (uint64_t)g.ny - row0);
    const uint64_t cells = rows * (uint64_t)g.nx, off = row0 * (uint64_t)g.nx;
    const size_t need = cells * 18 + 64;
    if ((rc = ensure(h, h->stage, need))) return rc;
    uint8_t* s = (uint8_t*)h->stage.p;
    int64_t* s_cd = (int64_t*)s;
    int64_t* s_el = (int64_t*)(s + cells * 8);
    uint8_t* s_fg = s + cells * 16;
    uint8_t* s_wm = s + cells * 17;
    if (cd) ABM_CK(h, copy_h2d(s_cd, cd + off, cells * 8, h->stream));
    if (fg) ABM_CK(h, copy_h2d(s_fg, fg + off, cells, h->stream));
    if (wm) ABM_CK(h, copy_h2d(s_wm, wm + off, cells, h->stream));
    if (el) ABM_CK(h, copy_h2d(s_el, el + off, cells * 8, h->stream));
    GridImportBody ib;
    ib.g = g; ib.fg = fg ? s_fg : nullptr; ib.cd = cd ? s_cd : nullptr; ib.wm = wm ? s_wm : nullptr; ib.el = el ? s_el : nullptr;
    ib.row0 = row0; ib.rows = rows;
    ib.grass = (uint8_t*)h->grass.p; ib.walkbits = (uint32_t*)h->walkbits.p; ib.elev = (int16_t*)h->elev.p;
    ib.err = (uint32_t*)h->counters.p + CTR_ERR;
    launch_foreach(h->stream, ib, cells);
    if ((rc = check_launch(h, "grid import"))) return rc;
    ABM_CK(h, stream_sync(h->stream));
  }
  if ((rc = read_counters(h))) return rc;
  if (h->h_counters[CTR_ERR]) {
    ABM_CK(h, dev_memset((uint32_t*)h->counters.p + CTR_ERR, 0, 4, h->stream));
    ABM_FAIL(h, ABM_E_BADARG, "countdown must be in 0..126 and elevation in -32768..32767");
  }
  h->grid_set = true;
  h->has_walkmap = wm != nullptr;
  return ABM_OK;
}

int ABM_API(get_grid)(abm_handle* h, uint8_t* fg, int64_t* cd) {
  if (!h) return ABM_E_BADARG;
  if (!h->grid_set) ABM_FAIL(h, ABM_E_STATE, "abm_set_grid has not been called");
  int rc = set_device(h);
  if (rc) return rc;
  const Geo& g = h->g;
  const uint64_t rows_per = std::max<uint64_t>(1, (16ull << 20) / (uint64_t)g.nx);
  for (uint64_t row0 = 0; row0 < (uint64_t)g.ny; row0 += rows_per) {
    const uint64_t rows = std::min<uint64_t>(rows_per, (uint64_t)g.ny - row0);
    const uint64_t cells = rows * (uint64_t)g.nx, off = row0 * (uint64_t)g.nx;
    if ((rc = ensure(h, h->stage, cells * 9 + 64))) return rc;
    uint8_t* s = (uint8_t*)h->stage.p;
    GridExportBody eb;
    eb.g = g; eb.grass = (const uint8_t*)h->grass.p; eb.row0 = row0;
    eb.cd = cd ? (int64_t*)s : nullptr; eb.fg = fg ? s + cells * 8 : nullptr;
    launch_foreach(h->stream, eb, cells);
    if ((rc = check_launch(h, "grid export"))) return rc;
    if (cd) ABM_CK(h, copy_d2h(cd + off, s, cells * 8, h->stream));
    if (fg) ABM_CK(h, copy_d2h(fg + off, s + cells * 8, cells, h->stream));
    ABM_CK(h, stream_sync(h->stream));
  }
  return ABM_OK;
}

/* room for n host-layout agent records (id, x, y: i64; energy: f64) in the staging buffer */
static int stage_agents(abm_handle* h, int64_t n) { return ensure(h, h->stage, (size_t)std::max<int64_t>(n, 1) * 32 + 64); }

/* the agent records waiting in the staging buffer become the population: validated, converted to the SoA and grouped by
 * cell (general path) or by 8x8 sub-tile (tile path).  Shared by abm_set_agents (records copied from the host) and
 * abm_init_random (records generated on the device). */
static int ingest_staged(abm_handle* h, int64_t n, int64_t next_id, bool packed = false) {
  int rc;
  const Geo& g = h->g;
  int64_t* s_id = (int64_t*)h->stage.p;
  int64_t* s_x = s_id + n;
  int64_t* s_y = s_x + n;
  double* s_e = (double*)(s_y + n);
  const uint32_t* p_id = nullptr; const uint32_t* p_xy = nullptr;
  if (packed) { /* abm_set_agents_packed staged [energy f64 x n][id u32 x n][x | y << 16 u32 x n] */
    s_e = (double*)h->stage.p;
    p_id = (const uint32_t*)(s_e + n); p_xy = p_id + n;
    s_id = s_x = s_y = nullptr;
  }
  const int tmp = h->cur ^ 1, dst = h->cur;
  if ((rc = ensure_soa(h, tmp, n))) return rc;
  if ((rc = ensure_soa(h, dst, n))) return rc;
  if ((rc = ensure_agent_scratch(h, n))) return rc;
  const uint64_t words = bitmap_words((uint64_t)next_id);
  if ((rc = ensure(h, h->bits, (words + 4) * 4))) return rc;
  ABM_CK(h, dev_memset(h->bits.p, 0, words * 4, h->stream));
  ABM_CK(h, dev_memset((uint32_t*)h->counters.p + CTR_ERR, 0, 4, h->stream));
  if (h->tile_mode) { /* group by 8x8 sub-tile: histogram, scan, scatter */
    const uint64_t nsub = (uint64_t)h->nsx * h->nsy;
    uint32_t* scnt = (uint32_t*)h->sub_cnt[dst].p;
    uint32_t* soff = (uint32_t*)h->sub_off[dst].p;
    ABM_CK(h, dev_memset(scnt, 0, nsub * 4, h->stream));
    if (n > 0) {
      if ((rc = ensure(h, h->ingest_sub, (size_t)n * 4))) return rc;
      IngestTileBody ib;
      ib.g = g; ib.nsx = h->nsx; ib.hid = s_id; ib.hx = s_x; ib.hy = s_y; ib.he = s_e; ib.pid = p_id; ib.pxy = p_xy; ib.next_id = next_id;
      ib.cell = (uint32_t*)h->cell[tmp].p; ib.id = (uint32_t*)h->id[tmp].p; ib.energy = (double*)h->energy[tmp].p;
      ib.slot = (uint32_t*)h->slot.p; ib.sub = (uint32_t*)h->ingest_sub.p; ib.sub_cnt = scnt;
      ib.alive_bits = (uint32_t*)h->bits.p; ib.err = (uint32_t*)h->counters.p + CTR_ERR;
      launch_foreach(h->stream, ib, (uint64_t)n);
      if ((rc = check_launch(h, "tile ingest"))) return rc;
    }
    ABM_CK(h, copy_d2d(soff, scnt, nsub * 4, h->stream));
    ABM_CK(h, scan_exclusive_u32(h->stream, soff, nsub, (uint32_t*)h->scan_tmp.p));
    if (n > 0) {
      ScatterTileBody sb;
      sb.cell = (const uint32_t*)h->cell[tmp].p; sb.id = (const uint32_t*)h->id[tmp].p; sb.energy = (const double*)h->energy[tmp].p;
      sb.slot = (const uint32_t*)h->slot.p; sb.sub = (const uint32_t*)h->ingest_sub.p; sb.sub_off = soff;
      sb.out_cell = (uint32_t*)h->cell[dst].p; sb.out_id = (uint32_t*)h->id[dst].p; sb.out_energy = (double*)h->energy[dst].p;
      launch_foreach(h->stream, sb, (uint64_t)n);
      if ((rc = check_launch(h, "tile ingest scatter"))) return rc;
    }
  }
  uint32_t* start = h->tile_mode ? nullptr : (uint32_t*)h->cell_start[dst].p;
  if (!h->tile_mode) ABM_CK(h, dev_memset(start, 0, ((size_t)g.gt + 1) * 4, h->stream));
  if (n > 0 && !h->tile_mode) {
    IngestBody ib;
    ib.g = g; ib.hid = s_id; ib.hx = s_x; ib.hy = s_y; ib.he = s_e; ib.pid = p_id; ib.pxy = p_xy; ib.next_id = next_id;
    ib.cell = (uint32_t*)h->cell[tmp].p; ib.id = (uint32_t*)h->id[tmp].p; ib.energy = (double*)h->energy[tmp].p;
    ib.state = (uint8_t*)h->state.p; ib.slot = (uint32_t*)h->slot.p;
    ib.new_count = start; ib.alive_bits = (uint32_t*)h->bits.p; ib.err = (uint32_t*)h->counters.p + CTR_ERR;
    launch_foreach(h->stream, ib, (uint64_t)n);
    if ((rc = check_launch(h, "ingest"))) return rc;
  }
  if (!h->tile_mode) ABM_CK(h, scan_exclusive_u32(h->stream, start, (uint64_t)g.gt + 1, (uint32_t*)h->scan_tmp.p));
  if (n > 0 && !h->tile_mode) {
    View v = make_view(h);
    v.cell = (uint32_t*)h->cell[tmp].p; v.id = (uint32_t*)h->id[tmp].p; v.energy = (double*)h->energy[tmp].p;
    ScatterBody sb;
    sb.v = v;
    sb.new_cell = (const uint32_t*)h->cell[tmp].p;
    sb.new_start = start;
    sb.out_cell = (uint32_t*)h->cell[dst].p; sb.out_id = (uint32_t*)h->id[dst].p; sb.out_energy = (double*)h->energy[dst].p;
    launch_foreach(h->stream, sb, (uint64_t)n);
    if ((rc = check_launch(h, "ingest scatter"))) return rc;
  }
  if ((rc = read_counters(h))) return rc;
  if (h->h_counters[CTR_ERR]) {
    ABM_CK(h, dev_memset((uint32_t*)h->counters.p + CTR_ERR, 0, 4, h->stream));
    h->agents_set = false; h->n = 0;
    ABM_FAIL(h, ABM_E_BADARG, "agent coordinates must lie inside the grid and ids be unique in 1..next_id-1");
  }
  h->n = (uint64_t)n;
  h->next_id = (uint64_t)next_id;
  h->agents_set = true;
  h->bcap_agreed = false;
  h->s_lists = false; h->s_canon = true; h->s_parent_used = 0; /* stream path: the canonical SoA is the population now */
  /* a fresh population has no routes (pathfinder.agent_paths is empty) */
  if (h->rt_slot_of_id.p) ABM_CK(h, dev_memset(h->rt_slot_of_id.p, 0, h->rt_slot_of_id.bytes, h->stream));
  ABM_CK(h, dev_memset((uint32_t*)h->counters.p + CTR_ROUTE_SLOTS, 0, 4, h->stream));
  h->rt_pool_used = 0;
  return ABM_OK;
}

int ABM_API(set_agents)(abm_handle* h, int64_t n, const int64_t* id, const int64_t* x, const int64_t* y,
                        const double* energy, int64_t next_id) {
  if (!h) return ABM_E_BADARG;
  if (n < 0 || (n > 0 && (!id || !x || !y || !energy))) ABM_FAIL(h, ABM_E_BADARG, "null agent arrays");
  if (next_id < 1 || next_id > 0xFFFFFFFFll || n >= 0xFFFFFFF0ll) ABM_FAIL(h, ABM_E_CAPACITY, "next_id must be in 1..2^32-1");
  int rc = set_device(h);
  if (rc) return rc;
  if ((rc = stage_agents(h, n))) return rc;
  if (n > 0) {
    int64_t* s_id = (int64_t*)h->stage.p;
    ABM_CK(h, copy_h2d_async(s_id, id, (size_t)n * 8, h->stream)); /* ingest_staged ends with a synchronising read-back */
    ABM_CK(h, copy_h2d_async(s_id + n, x, (size_t)n * 8, h->stream));
    ABM_CK(h, copy_h2d_async(s_id + 2 * n, y, (size_t)n * 8, h->stream));
    ABM_CK(h, copy_h2d_async(s_id + 3 * n, energy, (size_t)n * 8, h->stream));
  }
  rc = ingest_staged(h, n, next_id);
  if (rc) stream_sync(h->stream);
  return rc;
}

int ABM_API(set_agents_packed)(abm_handle* h, int64_t n, const uint32_t* id, const uint32_t* xy, const double* energy, int64_t next_id) {
  if (!h) return ABM_E_BADARG;
  if (n < 0 || (n > 0 && (!id || !xy || !energy))) ABM_FAIL(h, ABM_E_BADARG, "null agent arrays");
  if (next_id < 1 || next_id > 0xFFFFFFFFll || n >= 0xFFFFFFF0ll) ABM_FAIL(h, ABM_E_CAPACITY, "next_id must be in 1..2^32-1");
  if (h->cfg.nx > 65535 || h->cfg.ny > 65535) ABM_FAIL(h, ABM_E_BADARG, "packed agent records hold 16-bit coordinates: grid dims must be <= 65535");
  int rc = set_device(h);
  if (rc) return rc;
  if ((rc = stage_agents(h, n))) return rc;
  if (n > 0) { /* three copies enqueued back to back, one synchronisation */
    double* s_e = (double*)h->stage.p;
    uint32_t* s_id = (uint32_t*)(s_e + n);
    ABM_CK(h, copy_h2d_async(s_e, energy, (size_t)n * 8, h->stream));
    ABM_CK(h, copy_h2d_async(s_id, id, (size_t)n * 4, h->stream));
    ABM_CK(h, copy_h2d_async(s_id + n, xy, (size_t)n * 4, h->stream));
  }
  rc = ingest_staged(h, n, next_id, true);
  if (rc) stream_sync(h->stream); /* the host buffers must not be in flight when the caller gets them back */
  return rc;
}

/* the grass bytes alone: the terrain (walkmap, elevation) given to abm_set_grid stays */
int ABM_API(set_grass_packed)(abm_handle* h, const uint8_t* state) {
  if (!h || !state) return ABM_E_BADARG;
  int rc = set_device(h);
  if (rc) return rc;
  const Geo& g = h->g;
  const uint64_t cells = (uint64_t)g.nx * (uint64_t)g.ny;
  if (!h->grid_set) { /* no terrain was given: every cell walkable, elevation 0 (as abm_set_grid with NULL walkmap) */
    ABM_CK(h, dev_memset(h->grass.p, GRASS_PAD, g.gt, h->stream));
    ABM_CK(h, dev_memset(h->walkbits.p, 0, (size_t)g.gt / 8, h->stream));
    h->comp_valid = false;
  }
  const bool open_terrain = !h->grid_set;
  if ((rc = ensure(h, h->stage, cells + 64))) return rc;
  uint32_t* err = (uint32_t*)h->counters.p + CTR_ERR;
  ABM_CK(h, dev_memset(err, 0, 4, h->stream));
  ABM_CK(h, copy_h2d_async(h->stage.p, state, cells, h->stream));
  GrassPackedBody gb;
  gb.g = g; gb.grass = (uint8_t*)h->grass.p; gb.host = (uint8_t*)h->stage.p; gb.to_device = 1; gb.err = err; gb.open_walkbits = open_terrain ? (uint32_t*)h->walkbits.p : nullptr;
  launch_foreach(h->stream, gb, (uint64_t)((g.nx + 3) / 4) * (uint64_t)g.ny);
  if ((rc = check_launch(h, "grass import"))) return rc;
  if ((rc = read_counters(h))) return rc;
  if (h->h_counters[CTR_ERR]) {
    ABM_CK(h, dev_memset(err, 0, 4, h->stream));
    ABM_FAIL(h, ABM_E_BADARG, "countdown must be in 0..126");
  }
  h->grid_set = true;
  return ABM_OK;
}

int ABM_API(get_grass_packed)(abm_handle* h, uint8_t* state) {
  if (!h || !state) return ABM_E_BADARG;
  if (!h->grid_set) ABM_FAIL(h, ABM_E_STATE, "abm_set_grid has not been called");
  int rc = set_device(h);
  if (rc) return rc;
  const Geo& g = h->g;
  const uint64_t cells = (uint64_t)g.nx * (uint64_t)g.ny;
  if ((rc = ensure(h, h->stage, cells + 64))) return rc;
  GrassPackedBody gb;
  gb.g = g; gb.grass = (uint8_t*)h->grass.p; gb.host = (uint8_t*)h->stage.p; gb.to_device = 0; gb.err = nullptr; gb.open_walkbits = nullptr;
  launch_foreach(h->stream, gb, (uint64_t)((g.nx + 3) / 4) * (uint64_t)g.ny);
  if ((rc = check_launch(h, "grass export"))) return rc;
  ABM_CK(h, copy_d2h(state, h->stage.p, cells, h->stream));
  return ABM_OK;
}

/* the random part of initialize_model, built on the device (include/abm.h) */
int ABM_API(set_global_walkmap_bits)(abm_handle* h, const uint32_t* bits) {
  if (!h) return ABM_E_BADARG;
  if (!h->strip) ABM_FAIL(h, ABM_E_STATE, "abm_set_global_walkmap_bits is for strip handles (a single handle already holds the whole walkmap)");
  int rc = set_device(h);
  if (rc) return rc;
  if (!bits) { h->global_bits_set = false; return ABM_OK; }
  const uint64_t words = ((uint64_t)h->cfg.nx * (uint64_t)h->cfg.ny + 31) / 32;
  if ((rc = ensure(h, h->global_bits, words * 4))) return rc;
  ABM_CK(h, copy_h2d(h->global_bits.p, bits, words * 4, h->stream));
  h->global_bits_set = true;
  return ABM_OK;
}

int ABM_API(init_random)(abm_handle* h, int64_t n_sheep, uint64_t init_seed) {
  if (!h) return ABM_E_BADARG;
  const int64_t span = (int64_t)(h->cfg.delta_energy * 5);
  if (n_sheep < 0 || n_sheep >= 0xFFFFFFF0ll) ABM_FAIL(h, ABM_E_CAPACITY, "n_sheep must be below 2^32-16");
  if (span < 1) ABM_FAIL(h, ABM_E_BADARG, "delta_energy * 5 must be at least 1 (energy = rand(1:delta_energy*5) - 1)");
  if (h->strip && h->grid_set && h->has_walkmap && !h->global_bits_set)
    ABM_FAIL(h, ABM_E_STATE, "abm_init_random on a strip handle with terrain needs the walkmap of the whole GridSpace first (abm_set_global_walkmap_bits): random_walkable draws positions over all strips");
  if (h->strip && (h->cfg.nx > 65535 || h->cfg.ny > 65535)) ABM_FAIL(h, ABM_E_BADARG, "strip initialisation files 16-bit coordinates: grid dims must be <= 65535");
  int rc = set_device(h);
  if (rc) return rc;
  if ((rc = stream_invalidate(h))) return rc;
  const Geo& g = h->g;
  uint32_t* err = (uint32_t*)h->counters.p + CTR_ERR;
  ABM_CK(h, dev_memset(err, 0, 4, h->stream));
  if (!h->grid_set) { /* no terrain was given: every cell of the space is walkable, elevation 0 */
    ABM_CK(h, dev_memset(h->grass.p, GRASS_PAD, g.gt, h->stream));
    ABM_CK(h, dev_memset(h->walkbits.p, 0, (size_t)g.gt / 8, h->stream));
    if ((rc = ensure(h, h->elev, (size_t)g.gt * 2))) return rc;
    ABM_CK(h, dev_memset(h->elev.p, 0, (size_t)g.gt * 2, h->stream));
    h->comp_valid = false;
    if (h->strip) { /* the ghost rows of an open terrain are walkable too */
      GhostWalkableBody gw;
      gw.g = g; gw.walkbits = (uint32_t*)h->walkbits.p;
      launch_foreach(h->stream, gw, (uint64_t)g.nx * (uint64_t)g.ny);
    }
  }
  InitGrassBody gb;
  gb.g = g; gb.seed = init_seed; gb.regrowth_time = h->cfg.regrowth_time; gb.open_terrain = h->grid_set ? 0 : 1;
  gb.grass = (uint8_t*)h->grass.p; gb.walkbits = (uint32_t*)h->walkbits.p;
  launch_foreach(h->stream, gb, (uint64_t)g.nx * (uint64_t)(g.own_hi - g.own_lo));
  if ((rc = check_launch(h, "grass initialisation"))) return rc;
  const bool open_terrain = !h->grid_set;
  h->grid_set = true;
  InitSheepBody sb;
  memset(&sb, 0, sizeof sb);
  sb.g = g; sb.seed = init_seed; sb.span = (uint64_t)span; sb.walkbits = (const uint32_t*)h->walkbits.p; sb.err = err;
  if (h->strip) { /* every rank walks through all the sheep's draws and keeps the ones that land on its rows */
    uint32_t* cursor = (uint32_t*)h->counters.p + CTR_CURSOR;
    sb.strip = 1; sb.ny_global = h->cfg.ny; sb.global_bits = (!open_terrain && h->has_walkmap) ? (const uint32_t*)h->global_bits.p : nullptr;
    sb.cursor = cursor;
    ABM_CK(h, dev_memset(cursor, 0, 4, h->stream));
    if (n_sheep > 0) launch_foreach(h->stream, sb, (uint64_t)n_sheep); /* pass 1: how many are mine */
    if ((rc = check_launch(h, "sheep initialisation"))) return rc;
    if ((rc = read_counters(h))) return rc;
    if (h->h_counters[CTR_ERR] & ERR_NOWALKABLE) { ABM_CK(h, dev_memset(err, 0, 4, h->stream)); ABM_FAIL(h, ABM_E_NOWALKABLE, "random_walkable found no walkable cell in 4096 draws"); }
    const int64_t mine = (int64_t)h->h_counters[CTR_CURSOR];
    if ((rc = stage_agents(h, mine))) return rc;
    sb.p_e = (double*)h->stage.p; sb.p_id = (uint32_t*)(sb.p_e + mine); sb.p_xy = sb.p_id + mine;
    ABM_CK(h, dev_memset(cursor, 0, 4, h->stream));
    if (n_sheep > 0) launch_foreach(h->stream, sb, (uint64_t)n_sheep); /* pass 2: file them */
    if ((rc = check_launch(h, "sheep initialisation"))) return rc;
    return ingest_staged(h, mine, n_sheep + 1, true);
  }
  if ((rc = stage_agents(h, n_sheep))) return rc;
  if (n_sheep > 0) {
    sb.s_id = (int64_t*)h->stage.p; sb.s_x = sb.s_id + n_sheep; sb.s_y = sb.s_x + n_sheep; sb.s_e = (double*)(sb.s_y + n_sheep);
    launch_foreach(h->stream, sb, (uint64_t)n_sheep);
    if ((rc = check_launch(h, "sheep initialisation"))) return rc;
    if ((rc = read_counters(h))) return rc;
    if (h->h_counters[CTR_ERR] & ERR_NOWALKABLE) {
      ABM_CK(h, dev_memset(err, 0, 4, h->stream));
      ABM_FAIL(h, ABM_E_NOWALKABLE, "random_walkable found no walkable cell in 4096 draws");
    }
  }
  return ingest_staged(h, n_sheep, n_sheep + 1);
}

int ABM_API(count_agents)(abm_handle* h, int64_t* n) {
  if (!h || !n) return ABM_E_BADARG;
  *n = (int64_t)h->n;
  return ABM_OK;
}

static int export_agents(abm_handle* h, int64_t* n_inout, int64_t* id, int64_t* x, int64_t* y, double* energy, uint32_t* pid, uint32_t* pxy, bool packed) {
  if (!h || !n_inout) return ABM_E_BADARG;
  const int64_t n = (int64_t)h->n;
  if (*n_inout < n) { *n_inout = n; ABM_FAIL(h, ABM_E_SMALL, "agent buffers hold %lld entries, %lld needed", (long long)*n_inout, (long long)n); }
  *n_inout = n;
  if (n == 0) return ABM_OK;
  int rc = set_device(h);
  if (rc) return rc;
  if ((rc = stream_canonical(h))) return rc;
  const uint64_t words = bitmap_words(h->next_id);
  if ((rc = ensure(h, h->bits, (words + 4) * 4))) return rc;
  ABM_CK(h, dev_memset(h->bits.p, 0, words * 4, h->stream));
  SetBitsBody sbb;
  sbb.id = (const uint32_t*)h->id[h->cur].p; sbb.bits = (uint32_t*)h->bits.p;
  launch_foreach(h->stream, sbb, (uint64_t)n);
  if ((rc = rank_prefix(h, words))) return rc;
  if ((rc = ensure(h, h->stage, (size_t)n * 32 + 64))) return rc;
  int64_t* s_id = (int64_t*)h->stage.p;
  int64_t* s_x = s_id + n;
  int64_t* s_y = s_x + n;
  double* s_e = (double*)(s_y + n);
  uint32_t* sp_id = nullptr;
  if (packed) { s_e = (double*)h->stage.p; sp_id = (uint32_t*)(s_e + n); s_id = s_x = s_y = nullptr; }
  ExportBody eb;
  eb.g = h->g;
  eb.cell = (const uint32_t*)h->cell[h->cur].p; eb.id = (const uint32_t*)h->id[h->cur].p; eb.energy = (const double*)h->energy[h->cur].p;
  eb.alive_bits = (const uint32_t*)h->bits.p; eb.prefix = (const uint32_t*)h->bits_prefix.p;
  eb.oid = s_id; eb.ox = s_x; eb.oy = s_y; eb.oe = s_e; eb.pid = sp_id; eb.pxy = sp_id ? sp_id + n : nullptr;
  launch_foreach(h->stream, eb, (uint64_t)n);
  if ((rc = check_launch(h, "export"))) return rc;
  /* the copies are enqueued back to back and waited for once */
  if (packed) {
    if (pid) ABM_CK(h, copy_d2h_async(pid, sp_id, (size_t)n * 4, h->stream));
    if (pxy) ABM_CK(h, copy_d2h_async(pxy, sp_id + n, (size_t)n * 4, h->stream));
  } else {
    if (id) ABM_CK(h, copy_d2h_async(id, s_id, (size_t)n * 8, h->stream));
    if (x) ABM_CK(h, copy_d2h_async(x, s_x, (size_t)n * 8, h->stream));
    if (y) ABM_CK(h, copy_d2h_async(y, s_y, (size_t)n * 8, h->stream));
  }
  if (energy) ABM_CK(h, copy_d2h_async(energy, s_e, (size_t)n * 8, h->stream));
  ABM_CK(h, stream_sync(h->stream));
  return ABM_OK;
}

int ABM_API(get_agents)(abm_handle* h, int64_t* n_inout, int64_t* id, int64_t* x, int64_t* y, double* energy) {
  return export_agents(h, n_inout, id, x, y, energy, nullptr, nullptr, false);
}

int ABM_API(get_agents_packed)(abm_handle* h, int64_t* n_inout, uint32_t* id, uint32_t* xy, double* energy) {
  if (h && (h->cfg.nx > 65535 || h->cfg.ny > 65535)) ABM_FAIL(h, ABM_E_BADARG, "packed agent records hold 16-bit coordinates: grid dims must be <= 65535");
  return export_agents(h, n_inout, nullptr, nullptr, nullptr, energy, id, xy, true);
}

int ABM_API(set_global_state)(abm_handle* h, int64_t tick, int64_t behav, int64_t behav_counter) {
  if (!h) return ABM_E_BADARG;
  if (tick < 0 || tick > 0xFFFFFFFFll) ABM_FAIL(h, ABM_E_BADARG, "tick must fit 32 bits");
  if ((uint64_t)tick != h->tick) { int rc = stream_invalidate(h); if (rc) return rc; } /* the filed steps were drawn for another tick */
  if (h->cfg.walk_type == ABM_ALTERNATED_WALK && (behav < 0 || behav > 3 || behav_counter < 0 || behav_counter > h->cfg.counter))
    ABM_FAIL(h, ABM_E_BADARG, "behav must be in 0..3 and behav_counter in 0..counter");
  h->tick = (uint64_t)tick; h->behav = behav; h->behav_counter = behav_counter;
  return ABM_OK;
}

int ABM_API(get_global_state)(abm_handle* h, int64_t* tick, int64_t* behav, int64_t* behav_counter, int64_t* next_id) {
  if (!h) return ABM_E_BADARG;
  if (tick) *tick = (int64_t)h->tick;
  if (behav) *behav = h->behav;
  if (behav_counter) *behav_counter = h->behav_counter;
  if (next_id) *next_id = (int64_t)h->next_id;
  return ABM_OK;
}

int ABM_API(step)(abm_handle* h, int64_t n_ticks) {
  if (!h) return ABM_E_BADARG;
  if (n_ticks < 0) ABM_FAIL(h, ABM_E_BADARG, "n_ticks must be >= 0");
  if (!h->grid_set || !h->agents_set) ABM_FAIL(h, ABM_E_STATE, "abm_set_grid and abm_set_agents must precede abm_step");
  if (h->poisoned) ABM_FAIL(h, ABM_E_STATE, "this handle lost a peer exchange (time-out) and cannot be stepped any more");
  int rc = set_device(h);
  if (rc) return rc;
  const uint64_t l0 = launches_now();
  h->rounds = 0;
  h->updates = 0;
  h->prof.reset();
  h->tm_total.start(h->stream);
  for (int64_t t = 0; t < n_ticks; ++t) {
    if (h->tick >= 0xFFFFFFFFull) ABM_FAIL(h, ABM_E_CAPACITY, "tick counter exhausted");
    if ((rc = h->stream_mode ? tick_stream(h) : (h->tile_mode ? tick_tile(h) : tick_once(h)))) return rc;
  }
  h->tm_total.stop(h->stream);
  ABM_CK(h, stream_sync(h->stream));
  h->t_total = h->tm_total.ms();
  h->prof.collect();
  h->launches = (double)(launches_now() - l0);
  return ABM_OK;
}

/* this handle's share of a reduction; `any` = it holds at least one agent (min / max of nothing is undefined) */
static int reduce_local(abm_handle* h, int32_t what, double* out, bool* any) {
  int rc;
  *any = h->n > 0;
  if (what == ABM_REDUCE_COUNT_AGENTS) { *out = (double)h->n; return ABM_OK; }
  if (what == ABM_REDUCE_COUNT_GROWN) {
    if (!h->grid_set) ABM_FAIL(h, ABM_E_STATE, "abm_set_grid has not been called");
    unsigned long long* acc = (unsigned long long*)((uint32_t*)h->counters.p + CTR_RED_LO);
    ABM_CK(h, dev_memset(acc, 0, 8, h->stream));
    CountGrownBody cb;
    /* a strip counts its own rows only (whole tile rows: a contiguous range of the tile-major array) */
    const uint64_t first = (uint64_t)(h->g.own_lo / CORE) * (uint64_t)h->g.tiles_x * TILE_CELLS;
    const uint64_t cells = (uint64_t)((h->g.own_hi - h->g.own_lo + CORE - 1) / CORE) * (uint64_t)h->g.tiles_x * TILE_CELLS;
    cb.grass4 = (const uint32_t*)((const uint8_t*)h->grass.p + first); cb.out = acc;
    launch_foreach(h->stream, cb, cells / 4);
    if ((rc = check_launch(h, "count grown"))) return rc;
    unsigned long long v = 0;
    ABM_CK(h, copy_d2h(&v, acc, 8, h->stream));
    *out = (double)v;
    return ABM_OK;
  }
  if (what == ABM_REDUCE_SUM_ENERGY || what == ABM_REDUCE_MAX_ENERGY || what == ABM_REDUCE_MIN_ENERGY) {
    const uint64_t n = h->n;
    if (n == 0) { *out = 0; return ABM_OK; }
    if ((rc = stream_canonical(h))) return rc;
    const uint64_t chunks = (n + ENERGY_CHUNK - 1) / ENERGY_CHUNK;
    if ((rc = ensure(h, h->partials, chunks * 8))) return rc;
    EnergyPartialBody eb;
    eb.energy = (const double*)h->energy[h->cur].p; eb.n = n; eb.what = what; eb.out = (double*)h->partials.p;
    launch_foreach(h->stream, eb, chunks);
    if ((rc = check_launch(h, "energy reduce"))) return rc;
    std::vector<double> part(chunks);
    ABM_CK(h, copy_d2h(part.data(), h->partials.p, chunks * 8, h->stream));
    double v = part[0];
    for (uint64_t i = 1; i < chunks; ++i) {
      if (what == ABM_REDUCE_SUM_ENERGY) v += part[i];
      else if (what == ABM_REDUCE_MAX_ENERGY) v = part[i] > v ? part[i] : v;
      else v = part[i] < v ? part[i] : v;
    }
    *out = v;
    return ABM_OK;
  }
  ABM_FAIL(h, ABM_E_BADARG, "unknown reduction %d", what);
}

int ABM_API(reduce)(abm_handle* h, int32_t what, double* out) {
  if (!h || !out) return ABM_E_BADARG;
  int rc = set_device(h);
  if (rc) return rc;
  double mine = 0;
  bool any = false;
  if ((rc = reduce_local(h, what, &mine, &any))) return rc;
  if (!h->strip || h->n_ranks <= 1) { *out = mine; return ABM_OK; }
  /* strips: the value over the whole GridSpace, combined in rank order from every rank's share (collective call) */
  const int P = h->n_ranks;
  uint64_t snd[2];
  memcpy(&snd[0], &mine, 8);
  snd[1] = any ? 1 : 0;
  std::vector<uint64_t> all((size_t)2 * P);
  if ((rc = comm_allgather_small(h, snd, 2, all.data()))) return rc;
  double v = 0;
  bool first = true;
  for (int r = 0; r < P; ++r) {
    double x;
    memcpy(&x, &all[2 * r], 8);
    if (what == ABM_REDUCE_MAX_ENERGY || what == ABM_REDUCE_MIN_ENERGY) {
      if (!all[2 * r + 1]) continue;
      if (first || (what == ABM_REDUCE_MAX_ENERGY ? x > v : x < v)) v = x;
      first = false;
    } else v += x;
  }
  *out = v;
  return ABM_OK;
}

int ABM_API(reduce_where)(abm_handle* h, int32_t what, const abm_filter* f, double* out) {
  if (!h || !out || !f) return ABM_E_BADARG;
  if (what != ABM_REDUCE_COUNT_AGENTS && what != ABM_REDUCE_SUM_ENERGY && what != ABM_REDUCE_MAX_ENERGY && what != ABM_REDUCE_MIN_ENERGY)
    ABM_FAIL(h, ABM_E_BADARG, "abm_reduce_where takes the agent reductions (count, sum, max, min)");
  int rc = set_device(h);
  if (rc) return rc;
  double mine = 0;
  bool any = false;
  const uint64_t n = h->n;
  if (n > 0) {
    if ((rc = stream_canonical(h))) return rc;
    const uint64_t chunks = (n + ENERGY_CHUNK - 1) / ENERGY_CHUNK;
    if ((rc = ensure(h, h->partials, chunks * 12 + 64))) return rc;
    WherePartialBody wb;
    wb.g = h->g; wb.cell = (const uint32_t*)h->cell[h->cur].p; wb.energy = (const double*)h->energy[h->cur].p; wb.n = n; wb.what = what;
    wb.e_lo = f->energy_lo; wb.e_hi = f->energy_hi; wb.x_lo = f->x_lo - 1; wb.x_hi = f->x_hi - 1;
    wb.y_lo = f->y_lo - 1 - h->g.y_off; wb.y_hi = f->y_hi - 1 - h->g.y_off; /* stored rows (a strip keeps its rows at an offset) */
    wb.out = (double*)h->partials.p; wb.any = (uint32_t*)((double*)h->partials.p + chunks);
    launch_foreach(h->stream, wb, chunks);
    if ((rc = check_launch(h, "filtered reduce"))) return rc;
    std::vector<double> part(chunks);
    std::vector<uint32_t> cnt(chunks);
    ABM_CK(h, copy_d2h(part.data(), wb.out, chunks * 8, h->stream));
    ABM_CK(h, copy_d2h(cnt.data(), wb.any, chunks * 4, h->stream));
    for (uint64_t i = 0; i < chunks; ++i) {
      if (!cnt[i]) continue;
      if (what == ABM_REDUCE_COUNT_AGENTS || what == ABM_REDUCE_SUM_ENERGY) mine += part[i];
      else if (!any || (what == ABM_REDUCE_MAX_ENERGY ? part[i] > mine : part[i] < mine)) mine = part[i];
      any = true;
    }
  }
  if (!h->strip || h->n_ranks <= 1) { *out = mine; return ABM_OK; }
  const int P = h->n_ranks;
  uint64_t snd[2];
  memcpy(&snd[0], &mine, 8);
  snd[1] = any ? 1 : 0;
  std::vector<uint64_t> all((size_t)2 * P);
  if ((rc = comm_allgather_small(h, snd, 2, all.data()))) return rc;
  double v = 0;
  bool first = true;
  for (int r = 0; r < P; ++r) {
    double x;
    memcpy(&x, &all[2 * r], 8);
    if (what == ABM_REDUCE_MAX_ENERGY || what == ABM_REDUCE_MIN_ENERGY) {
      if (!all[2 * r + 1]) continue;
      if (first || (what == ABM_REDUCE_MAX_ENERGY ? x > v : x < v)) v = x;
      first = false;
    } else v += x;
  }
  *out = v;
  return ABM_OK;
}

int ABM_API(get_heat)(abm_handle* h, int32_t which, float* out) {
  if (!h || !out || which < 0 || which > 1) return ABM_E_BADARG;
  if (!h->grid_set) ABM_FAIL(h, ABM_E_STATE, "abm_set_grid has not been called");
  int rc = set_device(h);
  if (rc) return rc;
  const Geo& g = h->g;
  const uint64_t cells = (uint64_t)g.nx * (uint64_t)g.ny;
  if ((rc = ensure(h, h->stage, cells * 4 + 64))) return rc;
  HeatExportBody hb;
  hb.g = g; hb.grass = (const uint8_t*)h->grass.p; hb.elev = (const int16_t*)h->elev.p; hb.which = which;
  hb.inv_rt = h->cfg.regrowth_time > 0 ? 1.0f / (float)h->cfg.regrowth_time : 0.0f; hb.out = (float*)h->stage.p;
  launch_foreach(h->stream, hb, cells);
  if ((rc = check_launch(h, "heat export"))) return rc;
  ABM_CK(h, copy_d2h(out, h->stage.p, cells * 4, h->stream));
  return ABM_OK;
}

int ABM_API(plan_routes)(abm_handle* h, int64_t b, const int64_t* agent_id, const int64_t* goal_x, const int64_t* goal_y) {
  if (!h) return ABM_E_BADARG;
  if (b < 0 || (b > 0 && (!agent_id || !goal_x || !goal_y))) ABM_FAIL(h, ABM_E_BADARG, "null request arrays");
  if (!h->grid_set || !h->agents_set) ABM_FAIL(h, ABM_E_STATE, "abm_set_grid and abm_set_agents must precede abm_plan_routes");
  int rc = set_device(h);
  if (rc) return rc;
  h->expansions = 0;
  h->prof.reset();
  ABM_CK(h, dev_memset((uint32_t*)h->counters.p + CTR_EXP_LO, 0, 8, h->stream));
  if (b == 0) return ABM_OK;
  { /* each agent at most once per batch (the reference plans sequentially; a repeated id would be order-dependent) */
    std::vector<int64_t> srt(agent_id, agent_id + b);
    std::sort(srt.begin(), srt.end());
    for (int64_t i = 1; i < b; ++i) if (srt[i] == srt[i - 1]) ABM_FAIL(h, ABM_E_BADARG, "agent id %lld appears twice in one abm_plan_routes batch", (long long)srt[i]);
  }
  const uint64_t n = h->n;
  if ((rc = stream_canonical(h))) return rc;
  if ((rc = ensure(h, h->idx_of_id, (h->next_id + 1) * 4))) return rc;
  if ((rc = ensure_route_ids(h, h->next_id))) return rc;
  ABM_CK(h, dev_memset(h->idx_of_id.p, 0xFF, (h->next_id + 1) * 4, h->stream));
  if (n > 0) {
    IdIndexBody ii;
    ii.id = (const uint32_t*)h->id[h->cur].p; ii.idx_of_id = (uint32_t*)h->idx_of_id.p;
    launch_foreach(h->stream, ii, n);
  }
  if ((rc = ensure(h, h->stage, (size_t)b * 24 + 64))) return rc;
  if ((rc = ensure(h, h->as_id, (size_t)b * 4))) return rc;
  if ((rc = ensure(h, h->as_src, (size_t)b * 4))) return rc;
  if ((rc = ensure(h, h->as_dst, (size_t)b * 4))) return rc;
  int64_t* s_id = (int64_t*)h->stage.p;
  int64_t* s_x = s_id + b;
  int64_t* s_y = s_x + b;
  ABM_CK(h, copy_h2d(s_id, agent_id, (size_t)b * 8, h->stream));
  ABM_CK(h, copy_h2d(s_x, goal_x, (size_t)b * 8, h->stream));
  ABM_CK(h, copy_h2d(s_y, goal_y, (size_t)b * 8, h->stream));
  ABM_CK(h, dev_memset((uint32_t*)h->counters.p + CTR_ERR, 0, 4, h->stream));
  PlanRequestBody pr;
  pr.g = h->g; pr.hid = s_id; pr.gx = s_x; pr.gy = s_y; pr.idx_of_id = (const uint32_t*)h->idx_of_id.p;
  pr.id = (const uint32_t*)h->id[h->cur].p; pr.cell = (const uint32_t*)h->cell[h->cur].p; pr.n = (uint32_t)n; pr.next_id = h->next_id;
  pr.aid = (uint32_t*)h->as_id.p; pr.src = (uint32_t*)h->as_src.p; pr.dst = (uint32_t*)h->as_dst.p; pr.err = (uint32_t*)h->counters.p + CTR_ERR;
  launch_foreach(h->stream, pr, (uint64_t)b);
  if ((rc = check_launch(h, "plan requests"))) return rc;
  if ((rc = read_counters(h))) return rc;
  if (h->h_counters[CTR_ERR]) {
    ABM_CK(h, dev_memset((uint32_t*)h->counters.p + CTR_ERR, 0, 4, h->stream));
    ABM_FAIL(h, ABM_E_BADARG, "abm_plan_routes: unknown agent id or goal outside the grid");
  }
  h->tm_total.start(h->stream);
  rc = plan_batch(h, (uint64_t)b, (const uint32_t*)h->as_id.p, (const uint32_t*)h->as_src.p, (const uint32_t*)h->as_dst.p);
  h->tm_total.stop(h->stream);
  if (rc) return rc;
  ABM_CK(h, stream_sync(h->stream));
  h->t_total = h->tm_total.ms();
  h->prof.collect();
  return ABM_OK;
}

int ABM_API(get_paths)(abm_handle* h, int64_t b, const int64_t* agent_id, int64_t* offset, int64_t* cells_inout, int64_t* px, int64_t* py, int64_t* cost) {
  if (!h) return ABM_E_BADARG;
  if (b < 0 || !offset || !cells_inout || (b > 0 && !agent_id)) ABM_FAIL(h, ABM_E_BADARG, "null argument");
  int rc = set_device(h);
  if (rc) return rc;
  if (b == 0) { offset[0] = 0; *cells_inout = 0; return ABM_OK; }
  if ((rc = ensure_route_ids(h, h->next_id))) return rc;
  if ((rc = ensure(h, h->rt_off, 8))) return rc;
  if ((rc = ensure(h, h->rt_len, 4))) return rc;
  if ((rc = ensure(h, h->rt_pos, 4))) return rc;
  if ((rc = ensure(h, h->rt_cost, 8))) return rc;
  if ((rc = ensure(h, h->rt_pool, 4))) return rc;
  if ((rc = ensure(h, h->stage, (size_t)b * 32 + 64))) return rc;
  int64_t* s_id = (int64_t*)h->stage.p;
  int64_t* s_len = s_id + b;
  int64_t* s_cost = s_len + b;
  int64_t* s_off = s_cost + b;
  ABM_CK(h, copy_h2d(s_id, agent_id, (size_t)b * 8, h->stream));
  PathInfoBody pi;
  pi.rs = route_store(h); pi.r_cost = (const int64_t*)h->rt_cost.p; pi.hid = s_id; pi.ids_cap = h->rt_ids_cap;
  pi.out_len = s_len; pi.out_cost = s_cost;
  launch_foreach(h->stream, pi, (uint64_t)b);
  if ((rc = check_launch(h, "path info"))) return rc;
  std::vector<int64_t> len(b), off(b + 1);
  ABM_CK(h, copy_d2h(len.data(), s_len, (size_t)b * 8, h->stream));
  if (cost) ABM_CK(h, copy_d2h(cost, s_cost, (size_t)b * 8, h->stream));
  off[0] = 0;
  for (int64_t i = 0; i < b; ++i) off[i + 1] = off[i] + len[i];
  const int64_t total = off[b];
  if (*cells_inout < total) { *cells_inout = total; ABM_FAIL(h, ABM_E_SMALL, "path buffers hold fewer than %lld cells", (long long)total); }
  *cells_inout = total;
  memcpy(offset, off.data(), (size_t)(b + 1) * 8);
  if (total == 0) return ABM_OK;
  if (!px || !py) ABM_FAIL(h, ABM_E_BADARG, "null path buffers");
  DevBuf out;
  if ((rc = ensure(h, out, (size_t)total * 16))) return rc;
  ABM_CK(h, copy_h2d(s_off, off.data(), (size_t)b * 8, h->stream));
  PathCopyBody pc;
  pc.g = h->g; pc.rs = route_store(h); pc.hid = s_id; pc.ids_cap = h->rt_ids_cap; pc.offset = s_off;
  pc.px = (int64_t*)out.p; pc.py = (int64_t*)out.p + total;
  launch_foreach(h->stream, pc, (uint64_t)b);
  if ((rc = check_launch(h, "path copy"))) return rc;
  ABM_CK(h, copy_d2h(px, pc.px, (size_t)total * 8, h->stream));
  ABM_CK(h, copy_d2h(py, pc.py, (size_t)total * 8, h->stream));
  return ABM_OK;
}

int ABM_API(set_transport)(abm_handle* h, const abm_transport* t) {
  if (!h || !t || !t->sendrecv || !t->allreduce_sum_u64 || !t->allgather_u32) return ABM_E_BADARG;
  h->tr = *t;
  h->has_tr = true;
  return ABM_OK;
}

int ABM_API(nccl_unique_id)(void* out128) {
#ifndef ABM_EMU
  if (!out128) return ABM_E_BADARG;
  if (!g_nccl.load()) { g_create_error = "libnccl.so.2 could not be loaded"; return ABM_E_CUDA; }
  ncclUniqueId id;
  if (g_nccl.GetUniqueId(&id) != ncclSuccess) { g_create_error = "ncclGetUniqueId failed"; return ABM_E_CUDA; }
  static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
  memcpy(out128, &id, 128);
  return ABM_OK;
#else
  (void)out128;
  return ABM_E_STATE;
#endif
}

int ABM_API(comm_init_nccl)(abm_handle* h, const void* unique_id128) {
  if (!h || !unique_id128) return ABM_E_BADARG;
#ifndef ABM_EMU
  if (!h->strip || h->n_ranks < 2) ABM_FAIL(h, ABM_E_STATE, "abm_comm_init_nccl needs a strip handle with n_ranks > 1");
  int rc = set_device(h);
  if (rc) return rc;
  if (!g_nccl.load()) ABM_FAIL(h, ABM_E_CUDA, "libnccl.so.2 could not be loaded");
  ncclUniqueId id;
  memcpy(&id, unique_id128, 128);
  ncclComm_t c;
  ABM_NCCL(h, g_nccl.CommInitRank(&c, h->n_ranks, id, h->my_rank));
  h->nccl_comm = (void*)c;
  return ABM_OK;
#else
  ABM_FAIL(h, ABM_E_STATE, "NCCL is not part of the emulation build");
#endif
}

int ABM_API(p2p_export)(abm_handle* h, void* handle64) {
  if (!h || !handle64) return ABM_E_BADARG;
#ifndef ABM_EMU
  if (!h->strip || h->n_ranks < 2 || h->n_ranks > P2P_MAXP) ABM_FAIL(h, ABM_E_STATE, "abm_p2p_export needs a strip handle with 2..%d ranks", (int)P2P_MAXP);
  if (!h->agents_set) ABM_FAIL(h, ABM_E_STATE, "abm_set_agents must precede abm_p2p_export (the mailbox is sized from the population)");
  int rc = set_device(h);
  if (rc) return rc;
  if (!h->p2p_mine) {
    /* Every rank addresses its neighbours' mailboxes with its OWN layout, so the layout must be the same on all ranks: it is
     * derived from the grid alone (never from this rank's population).  A boundary sub-tile row may hold 16 agents per
     * cell on average; offspring per tick and rank: 1/32 of a rank's share of the cells. */
    h->p2p_capmax = std::max<uint64_t>(1u << 16, (uint64_t)h->cfg.nx * SUB * 16);
    h->p2p_birthmax = std::max<uint64_t>(1u << 16, (uint64_t)h->cfg.nx * (uint64_t)h->cfg.ny / (uint64_t)h->n_ranks / 32);
    h->p2p_lay.agents_bytes = (strip_msg_bytes(h->nsx, h->p2p_capmax) + 255) & ~(uint64_t)255;
    h->p2p_lay.state_bytes = (h->p2p_capmax + 255) & ~(uint64_t)255;
    h->p2p_lay.births_bytes = ((1 + h->p2p_birthmax) * 4 + 255) & ~(uint64_t)255;
    ABM_CK(h, cudaMalloc((void**)&h->p2p_mine, h->p2p_lay.total_bytes()));
    ABM_CK(h, cudaMemset(h->p2p_mine, 0, h->p2p_lay.total_bytes()));
    ABM_CK(h, cudaStreamSynchronize(cudaStreamLegacy));
  }
  cudaIpcMemHandle_t mh;
  ABM_CK(h, cudaIpcGetMemHandle(&mh, h->p2p_mine));
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
  memcpy(handle64, &mh, 64);
  return ABM_OK;
#else
  ABM_FAIL(h, ABM_E_STATE, "peer-memory transport is not part of the emulation build");
#endif
}

#ifndef ABM_EMU
/* ABM_P2P_TIMEOUT_S (1..3600, default 30): how long a kernel waits for a neighbour's flag before it gives up (per device) */
static int p2p_apply_timeout(abm_handle* h) {
  long secs = 30;
  if (const char* e = getenv("ABM_P2P_TIMEOUT_S")) { const long q = atol(e); if (q >= 1 && q <= 3600) secs = q; }
  int khz = 0;
  if (cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, h->device) != cudaSuccess || khz <= 0) khz = 2000000;
  const long long cycles = (long long)secs * (long long)khz * 1000ll;
  ABM_CK(h, cudaMemcpyToSymbol(abm::abm_flag_timeout_cycles, &cycles, sizeof cycles));
  return 0;
}
#endif

int ABM_API(p2p_connect)(abm_handle* h, const void* handles, int32_t n) {
  if (!h || !handles) return ABM_E_BADARG;
#ifndef ABM_EMU
  if (!h->p2p_mine || n != h->n_ranks) ABM_FAIL(h, ABM_E_STATE, "abm_p2p_connect needs abm_p2p_export first and one handle per rank");
  int rc = set_device(h);
  if (rc) return rc;
  if ((rc = p2p_apply_timeout(h))) return rc;
  for (int r = 0; r < n; ++r) {
    if (r == h->my_rank) { h->p2p_peer[r] = h->p2p_mine; continue; }
    cudaIpcMemHandle_t mh;
    memcpy(&mh, (const uint8_t*)handles + 64 * r, 64);
    void* p = nullptr;
    ABM_CK(h, cudaIpcOpenMemHandle(&p, mh, cudaIpcMemLazyEnablePeerAccess));
    h->p2p_peer[r] = (uint8_t*)p;
  }
  h->p2p_on = true;
  return ABM_OK;
#else
  (void)n;
  ABM_FAIL(h, ABM_E_STATE, "peer-memory transport is not part of the emulation build");
#endif
}

int ABM_API(set_profiling)(abm_handle* h, int32_t on) {
  if (!h) return ABM_E_BADARG;
  h->prof.on = on != 0;
  return ABM_OK;
}

int ABM_API(get_kernel_stat)(abm_handle* h, int32_t idx, char* name, int32_t name_cap, double* ms, int64_t* launches, int64_t* units) {
  if (!h || idx < 0) return ABM_E_BADARG;
  if ((size_t)idx >= h->prof.stats.size()) return ABM_E_SMALL;
  const KStat& k = h->prof.stats[idx];
  if (name && name_cap > 0) { snprintf(name, (size_t)name_cap, "%s", k.name.c_str()); }
  if (ms) *ms = k.ms;
  if (launches) *launches = (int64_t)k.launches;
  if (units) *units = (int64_t)k.units;
  return ABM_OK;
}

int ABM_API(get_timing)(abm_handle* h, double* out, int32_t n) {
  if (!h || !out) return ABM_E_BADARG;
  const double v[7] = {h->t_total, h->t_agents, h->t_grass, h->rounds, h->launches, h->expansions, h->updates};
  for (int i = 0; i < n; ++i) out[i] = i < 7 ? v[i] : 0.0;
  return ABM_OK;
}

}  // extern "C"

#if defined(ABM_PHASE_CLOCKS) && !defined(ABM_EMU)
/* tuning builds only: the summed phase clocks of the instrumented kernels (32 counters), reset after the read */
extern "C" int abm_debug_phase_clocks(unsigned long long* out) {
  cudaDeviceSynchronize();
  if (cudaMemcpyFromSymbol(out, abm::abm_phase_clk, sizeof(unsigned long long) * 32) != cudaSuccess) return ABM_E_CUDA;
  unsigned long long zero[32] = {0};
  return cudaMemcpyToSymbol(abm::abm_phase_clk, zero, sizeof zero) == cudaSuccess ? 0 : ABM_E_CUDA;
}
#endif
