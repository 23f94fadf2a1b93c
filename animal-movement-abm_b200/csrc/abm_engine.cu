/*
 * abm_engine.cu — host side of libabm.so: owns device memory, orders the kernels of one tick, implements the
 * C ABI of include/abm.h.  Compiled by nvcc for sm_100a (the product).  tests/emu compiles the same file with
 * g++ -DABM_EMU to exercise this orchestration on a box without a GPU (see abm_platform.h); that build is
 * never shipped nor loaded by the package.
 */
#include <math.h>
#include <stdio.h>
#ifndef ABM_EMU
#include <dlfcn.h>
#include <nccl.h>
#endif

#include <algorithm>
#include <new>
#include <string>
#include <vector>

#include "abm_kernels.h"
#include "abm_astar.h"
#include "abm_tile.h"
#include "abm_stream.h"
#include "abm_p2p.h"

#ifdef ABM_EMU
#define ABM_API(name) emu_##name
namespace abm {
uint64_t g_emu_launches = 0;
uint64_t g_emu_shuffle_seed = 12345;
}
extern "C" void emu_set_shuffle_seed(uint64_t s) { abm::g_emu_shuffle_seed = s; }
#else
#define ABM_API(name) abm_##name
namespace abm {
unsigned long long g_launches = 0;
}
#endif

using namespace abm;

namespace {

static std::string g_create_error;

/* ------------------------------------------------------------------------------------------ device scan */
#ifdef ABM_EMU
static abm_err_t scan_exclusive_u32(abm_stream_t, uint32_t* data, uint64_t n, uint32_t*) {
  uint32_t run = 0;
  for (uint64_t i = 0; i < n; ++i) { const uint32_t v = data[i]; data[i] = run; run += v; }
  ++g_emu_launches;
  return 0;
}
static uint64_t scan_tmp_words(uint64_t) { return 1; }
#else
enum { SCAN_THREADS = 256, SCAN_ITEMS = 16, SCAN_TILE = SCAN_THREADS * SCAN_ITEMS };

__device__ __forceinline__ uint32_t block_exclusive_scan(uint32_t v, uint32_t* total_out) {
  __shared__ uint32_t warp_sums[SCAN_THREADS / 32];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  uint32_t inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += t;
  }
  if (lane == 31) warp_sums[wid] = inc;
  __syncthreads();
  if (wid == 0) {
    uint32_t w = lane < SCAN_THREADS / 32 ? warp_sums[lane] : 0;
#pragma unroll
    for (int o = 1; o < SCAN_THREADS / 32; o <<= 1) {
      const uint32_t t = __shfl_up_sync(0xffffffffu, w, o);
      if (lane >= o) w += t;
    }
    if (lane < SCAN_THREADS / 32) warp_sums[lane] = w; /* inclusive over warps */
  }
  __syncthreads();
  const uint32_t before = wid ? warp_sums[wid - 1] : 0;
  *total_out = warp_sums[SCAN_THREADS / 32 - 1];
  return before + inc - v;
}

/* pass 1: one partial sum per 4096-element tile */
__global__ void __launch_bounds__(SCAN_THREADS) k_scan_reduce(const uint32_t* __restrict__ in, uint64_t n, uint32_t* __restrict__ sums) {
  const uint64_t base = (uint64_t)blockIdx.x * SCAN_TILE;
  uint32_t s = 0;
  if (base + SCAN_TILE <= n) {
    const uint4* p = reinterpret_cast<const uint4*>(in + base);
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS / 4; ++i) { const uint4 q = p[i * SCAN_THREADS + threadIdx.x]; s += q.x + q.y + q.z + q.w; }
  } else {
    for (uint64_t i = base + threadIdx.x; i < n; i += SCAN_THREADS) s += in[i];
  }
  uint32_t total;
  (void)block_exclusive_scan(s, &total);
  if (threadIdx.x == 0) sums[blockIdx.x] = total;
}

/* pass 2: exclusive scan inside each tile plus the tile's offset */
__global__ void __launch_bounds__(SCAN_THREADS) k_scan_apply(uint32_t* __restrict__ data, uint64_t n, const uint32_t* __restrict__ offsets) {
  const uint64_t base = (uint64_t)blockIdx.x * SCAN_TILE + (uint64_t)threadIdx.x * SCAN_ITEMS;
  uint32_t v[SCAN_ITEMS];
  if (base + SCAN_ITEMS <= n) {
    const uint4* p = reinterpret_cast<const uint4*>(data + base);
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS / 4; ++i) { const uint4 q = p[i]; v[4 * i] = q.x; v[4 * i + 1] = q.y; v[4 * i + 2] = q.z; v[4 * i + 3] = q.w; }
  } else {
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; ++i) v[i] = base + i < n ? data[base + i] : 0u;
  }
  uint32_t s = 0;
#pragma unroll
  for (int i = 0; i < SCAN_ITEMS; ++i) s += v[i];
  uint32_t total;
  uint32_t run = block_exclusive_scan(s, &total) + (offsets ? offsets[blockIdx.x] : 0u);
#pragma unroll
  for (int i = 0; i < SCAN_ITEMS; ++i) { const uint32_t t = v[i]; v[i] = run; run += t; }
  if (base + SCAN_ITEMS <= n) {
    uint4* p = reinterpret_cast<uint4*>(data + base);
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS / 4; ++i) p[i] = make_uint4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
  } else {
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; ++i) if (base + i < n) data[base + i] = v[i];
  }
}

static uint64_t scan_tmp_words(uint64_t n) {
  uint64_t w = 0;
  while (n > SCAN_TILE) { n = (n + SCAN_TILE - 1) / SCAN_TILE; w += (n + 3) & ~3ull; }
  return w + 4;
}

/* in-place exclusive prefix sum; data must be 16-byte aligned */
static abm_err_t scan_exclusive_u32(abm_stream_t s, uint32_t* data, uint64_t n, uint32_t* tmp) {
  if (n == 0) return cudaSuccess;
  const uint64_t nb = (n + SCAN_TILE - 1) / SCAN_TILE;
  if (nb == 1) {
    ++g_launches;
    k_scan_apply<<<1, SCAN_THREADS, 0, s>>>(data, n, nullptr);
    return cudaGetLastError();
  }
  g_launches += 2;
  k_scan_reduce<<<(unsigned)nb, SCAN_THREADS, 0, s>>>(data, n, tmp);
  abm_err_t e = scan_exclusive_u32(s, tmp, nb, tmp + ((nb + 3) & ~3ull));
  if (e != cudaSuccess) return e;
  k_scan_apply<<<(unsigned)nb, SCAN_THREADS, 0, s>>>(data, n, tmp);
  return cudaGetLastError();
}
#endif

/* ------------------------------------------------------------------------------------------- buffers */
struct DevBuf {
  void* p = nullptr;
  size_t bytes = 0;
  ~DevBuf() { dev_free(p); }
};

struct Timer {
#ifndef ABM_EMU
  cudaEvent_t a = nullptr, b = nullptr;
  void init() { cudaEventCreate(&a); cudaEventCreate(&b); }
  void destroy() { if (a) cudaEventDestroy(a); if (b) cudaEventDestroy(b); }
  void start(abm_stream_t s) { cudaEventRecord(a, s); }
  void stop(abm_stream_t s) { cudaEventRecord(b, s); }
  double ms() { float f = 0; cudaEventSynchronize(b); cudaEventElapsedTime(&f, a, b); return f; }
#else
  void init() {}
  void destroy() {}
  void start(abm_stream_t) {}
  void stop(abm_stream_t) {}
  double ms() { return 0; }
#endif
};

/* per-kernel device timing of the last abm_step (CUDA events on the handle's stream), for bench.py's roofline */
struct KStat { std::string name; double ms = 0; uint64_t launches = 0, units = 0; };
struct Prof {
  bool on = false;
  std::vector<KStat> stats;
#ifndef ABM_EMU
  struct Rec { int idx; cudaEvent_t a, b; bool own_a; };
  std::vector<Rec> recs;
  unsigned long long launches_at_end = 0; /* a scope that starts with no launch since the previous one ended shares its event */
  bool last_ended = false;
  cudaStream_t last_stream = nullptr; /* ... and only when both are on the same stream */
  std::vector<cudaEvent_t> pool;
  cudaEvent_t get() { if (pool.empty()) { cudaEvent_t e; cudaEventCreate(&e); return e; } cudaEvent_t e = pool.back(); pool.pop_back(); return e; }
#endif
  int index(const char* name) {
    for (size_t i = 0; i < stats.size(); ++i) if (stats[i].name == name) return (int)i;
    KStat k; k.name = name; stats.push_back(k); return (int)stats.size() - 1;
  }
  void reset() { for (size_t i = 0; i < stats.size(); ++i) { stats[i].ms = 0; stats[i].launches = 0; stats[i].units = 0; } }
  void begin(abm_stream_t s, const char* name, uint64_t units, uint64_t launches) {
    const int i = index(name);
    stats[i].launches += launches; stats[i].units += units;
#ifndef ABM_EMU
    if (!on) return;
    if (recs.size() >= 8192) collect();
    Rec r; r.idx = i; r.b = get();
    if (!recs.empty() && last_ended && launches_at_end == g_launches && last_stream == s) { r.a = recs.back().b; r.own_a = false; }
    else { r.a = get(); r.own_a = true; cudaEventRecord(r.a, s); }
    last_ended = false;
    recs.push_back(r);
#else
    (void)s;
#endif
  }
  void end(abm_stream_t s) {
#ifndef ABM_EMU
    if (on && !recs.empty()) { cudaEventRecord(recs.back().b, s); launches_at_end = g_launches; last_ended = true; last_stream = s; }
#else
    (void)s;
#endif
  }
  void collect() {
#ifndef ABM_EMU
    for (size_t i = 0; i < recs.size(); ++i) {
      float f = 0;
      cudaEventSynchronize(recs[i].b);
      cudaEventElapsedTime(&f, recs[i].a, recs[i].b);
      stats[recs[i].idx].ms += f;
      if (recs[i].own_a) pool.push_back(recs[i].a);
      pool.push_back(recs[i].b);
    }
    recs.clear();
    last_ended = false;
#endif
  }
  void destroy() {
#ifndef ABM_EMU
    collect();
    for (size_t i = 0; i < pool.size(); ++i) cudaEventDestroy(pool[i]);
    pool.clear();
#endif
  }
};
struct ProfScope {
  Prof& p; abm_stream_t s;
  ProfScope(Prof& p_, abm_stream_t s_, const char* name, uint64_t units, uint64_t launches = 1) : p(p_), s(s_) { p.begin(s, name, units, launches); }
  ~ProfScope() { p.end(s); }
};

enum { CTR_BIRTH_TOTAL = 12, CTR_PEND_A = 0, CTR_PEND_B = 1, CTR_BIRTHS = 2, CTR_ERR = 3, CTR_RED_LO = 4, CTR_RED_HI = 5, CTR_PLANS = 6, CTR_ROUTE_SLOTS = 7,
       CTR_EXP_LO = 8, CTR_EXP_HI = 9, CTR_CURSOR = 10, CTR_BIRTH_FLAGS = 11, CTR_PEND_FIRST = 13, CTR_GO = 14, CTR_PACK_DONE = 16, CTR_POOL_LO = 20, CTR_POOL_HI = 21, CTR_NEXT_REQ = 22, CTR_BIRTH_DONE = 24 /* 8 words */, CTR_WORDS = 32 };

}  // namespace

struct abm_handle {
  abm_config cfg;
  Geo g;
  Params params;
  abm_stream_t stream = 0;
#ifndef ABM_EMU
  /* peer-memory transport (abm_p2p.h): this rank's mailbox, the peers' mailboxes mapped through CUDA IPC */
  uint8_t* p2p_mine = nullptr;
  uint8_t* p2p_peer[P2P_MAXP] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  P2PLayout p2p_lay;
  bool p2p_on = false;
  uint32_t p2p_seq[4] = {0, 0, 0, 0};
  uint64_t p2p_capmax = 0, p2p_birthmax = 0;
#endif
  int device = 0;

  DevBuf cell[2], id[2], energy[2]; /* ping-pong cell-sorted SoA */
  DevBuf cell_start[2];             /* gt + 1 (+ padding) */
  int cur = 0;
  DevBuf state, slot, new_cell, pend[2], births;
  DevBuf grass, walkbits, elev, global_bits; /* global_bits: strips, the whole GridSpace's walkmap (abm_set_global_walkmap_bits) */
  bool global_bits_set = false, has_walkmap = false;
  DevBuf bits, bits_prefix, scan_tmp, stage, lut, greg_order, greg_dxy, greg_table, counters, partials;
  /* ALTERNATED_WALK / A*: id-order ranks, behaviour segments, plan requests, route followers, the route store */
  DevBuf rank, by_rank, seg_behav, plan_agent, plan_goal, remote_head, remote_next, remote_dest, idx_of_id;
  DevBuf as_g, as_meta, as_heap, as_src, as_dst, as_id, as_status, as_len, as_cost, as_off, as_slot_gen, as_todo, as_comp;
  bool comp_valid = false;        /* as_comp holds the walkmap's component labels */
  uint64_t astar_budget = 128ull << 30; /* ABM_ASTAR_GB: device bytes the search slots may take (and at most three quarters of what is free:
                                         * 7/8 left too little for the buffers allocated afterwards) */
  DevBuf rt_slot_of_id, rt_off, rt_len, rt_pos, rt_cost, rt_pool;
  uint64_t as_slots = 0, as_heap_cap = 0, as_heap_want = 0, rt_slots_cap = 0, rt_pool_used = 0, rt_ids_cap = 0;
  /* tile path (abm_tile.h): sub-tile runs of the SoA (ping-pong with the SoA), offspring fix-up lists */
  bool tile_mode = false;
  uint64_t birth_bitmap_from = 32768; /* ABM_BIRTH_BITMAP: gathered offspring capacity above which ids are ranked through a bitmap */
  bool optimistic = true;         /* enqueue the commit behind a device-side guard instead of reading counters first (ABM_OPTIMISTIC=0 disables) */
  uint64_t bf_est = 0;            /* expected upper bound of reproducing agents, learnt from the previous tick */
  int32_t fixed_agent_rounds = 2; /* per-agent rounds enqueued without asking the host (ABM_AGENT_ROUNDS sets the floor) */
  int32_t min_agent_rounds = 2;   /* the guard failing (somebody still open) adds a round; 256 clean ticks in a row take one off.
                                   * Every rank sees the same guard, so strips stay in step. */
  int32_t clean_ticks = 0;
  uint64_t kf_grid = 1024;        /* CTAs of those launches: an upper bound learnt from the previous tick */
  int32_t tile_rounds = 8;        /* in-window rounds of the tile resolve kernel (ABM_TILE_ROUNDS overrides; tests use 1) */
  int32_t tile_margin = HALO;     /* halo cells actually taken into a core's window (ABM_TILE_MARGIN overrides) */
  DevBuf okey;                    /* Schedulers.randomly: order key per agent (own + ghost), recomputed every tick */
  /* offspring ids off the critical path: the reproducing agents are known after propose, so their order keys are gathered
   * and ranked on a second stream while the tick resolves and commits (ABM_EARLY_BIRTHS=0: after the commit, as before) */
  bool poisoned = false;          /* a peer exchange timed out: the state is garbage, every later call fails */
  bool early_births = true;
  abm_stream_t side = 0;
  abm_event_t ev_proposed = 0, ev_ranked = 0, ev_rows = 0, ev_halo = 0;
  /* peer-memory strips: exchanges beside compute on the side stream.  Tile path: the halo exchange beside propose — measured
   * 0.641 vs 0.664 ms at 2 GPUs but 0.747 vs 0.712 ms at 8, so it is off unless ABM_OVERLAP_HALO=1.  Stream path: the list
   * exchange beside the interior rows' kernel — 1.019 vs 1.045 ms at 8 GPUs (C5), on unless ABM_OVERLAP_LISTS=0. */
  bool overlap_halo = false, overlap_lists = true;
  DevBuf parents;                 /* [count][order keys of this handle's reproducing agents] */
  int32_t nsx = 0, nsy = 0;
  DevBuf sub_off[2], sub_cnt[2], birth_pos, birth_parent, ingest_sub, mig_cnt, mig_a, mig_lc;
  bool use_migrants = false; /* ABM_MIGRANTS=1: border crossers are listed per destination core and the commit reads only its
                              * own core + that list; measured slower than scanning the halo (profiles/README.md), so off */
  /* row strips: ring neighbours, transport, boundary-row staging ([0] = towards / from the previous strip, [1] = next) */
  bool strip = false;
  int32_t n_ranks = 1, my_rank = 0;
  bool has_prev = false, has_next = false;
  abm_transport tr;
  bool has_tr = false;
  void* nccl_comm = nullptr;
  DevBuf sx_pre[2], rx_pre[2], sx_buf[2], rx_buf[2], sx_st[2], rx_st[2], red_buf, birth_ids;
  uint64_t n_ghost = 0, bcap = 4096, next_bcap = 4096;
  bool bcap_agreed = false; /* capacity (agents) of a boundary message, the same on every rank */
  /* stream path (abm_stream.h): arrival lists of the current and the next tick, parents bitmaps, counters */
  bool stream_mode = false;      /* walk rule without conflicts on a tile-path grid (ABM_STREAM=0 disables) */
  bool s_lists = false;          /* the lists of tick `tick` are filed */
  bool s_canon = true;           /* cell/id/energy[cur] hold the population (the getters' view) */
  bool s_relayout = false;       /* file again with more slots per core */
  int s_cur = 0;
  uint32_t s_cap = 0, s_cap_forced = 0; /* record slots per core (ABM_STREAM_CAP forces a value: tests) */
  uint64_t s_ovf_cap = 0, s_n_ovf = 0, s_n_layout = 0, s_words[2] = {0, 0}, s_parent_cap = 0, s_parent_used = 0, s_msg_cap = 0;
  DevBuf s_rec[2], s_cnt[2], s_ovf_rec[2], s_ovf_tile[2], s_bits[2], s_ctr, s_parents, s_gather, s_tx[2], s_rx[2];
  std::vector<uint8_t> hstage[4];
  uint32_t* h_counters = nullptr; /* host mirror (pinned on CUDA) */

  uint64_t n = 0, next_id = 1, tick = 0;
  int64_t behav = 0, behav_counter = 0;
  bool grid_set = false, agents_set = false;
  double t_total = 0, t_agents = 0, t_grass = 0, rounds = 0, launches = 0, expansions = 0, updates = 0;
  Timer tm_total;
  Prof prof;
  std::string err;
};

namespace {

#define ABM_FAIL(h, code, ...) do { char b_[512]; snprintf(b_, sizeof b_, __VA_ARGS__); (h)->err = b_; return (code); } while (0)
#define ABM_CK(h, call) do { abm_err_t e_ = (call); if (e_ != 0) { ABM_FAIL(h, ABM_E_CUDA, "%s failed: %s", #call, err_string(e_)); } } while (0)

static int ensure(abm_handle* h, DevBuf& b, size_t bytes) {
  bytes = (bytes + 255) & ~(size_t)255;
  if (b.bytes >= bytes && b.p) return 0;
  dev_free(b.p);
  b.p = nullptr; b.bytes = 0;
  size_t want = bytes + bytes / 4; /* head room so that a growing population does not reallocate every tick */
  if (dev_malloc(&b.p, want) != 0) {
    b.p = nullptr;
    want = bytes;
    if (dev_malloc(&b.p, want) != 0) { b.p = nullptr; ABM_FAIL(h, ABM_E_NOMEM, "device allocation of %zu bytes failed", bytes); }
  }
  b.bytes = want;
  return 0;
}

static int set_device(abm_handle* h) {
#ifndef ABM_EMU
  ABM_CK(h, cudaSetDevice(h->device));
#else
  (void)h;
#endif
  return 0;
}

static uint64_t launches_now() {
#ifdef ABM_EMU
  return g_emu_launches;
#else
  return g_launches;
#endif
}

static int check_launch(abm_handle* h, const char* what) {
#ifndef ABM_EMU
  abm_err_t e = cudaGetLastError();
  if (e != cudaSuccess) ABM_FAIL(h, ABM_E_CUDA, "launch of %s failed: %s", what, cudaGetErrorString(e));
#else
  (void)h; (void)what;
#endif
  return 0;
}

static int build_greg_table(abm_handle* h) {
  const int r = h->cfg.walk_type == ABM_RANDOM_WALK_GREGARIOUS ? h->cfg.visual_distance : 1;
  const int side = 2 * r + 1;
  int rc = ensure(h, h->greg_table, (size_t)side * side * 8 * sizeof(double));
  if (rc) return rc;
  GregTableBody gb;
  gb.lut = (const double*)h->lut.p; gb.r = r; gb.table = (double*)h->greg_table.p;
  launch_foreach(h->stream, gb, (uint64_t)side * side);
  ABM_CK(h, stream_sync(h->stream));
  return 0;
}

/* the scheduler's order keys (abm_rng.h) of tick `tick`, when nextid(model) is `next_id` at its start */
static SchedKey sched_key(const abm_handle* h, uint64_t tick, uint64_t next_id) {
  SchedKey sk;
  memset(&sk, 0, sizeof sk);
  if (h->cfg.scheduler == ABM_SCHED_RANDOMLY) { sk.on = 1; sk.hb = sched_half_bits(next_id); sched_round_keys(h->cfg.seed, tick, sk.k); }
  return sk;
}
/* ids are below id_bound: what their order keys are below (bitmaps and ranks over the schedule are indexed by key) */
static uint64_t key_space(const abm_handle* h, uint64_t id_bound) {
  return h->cfg.scheduler == ABM_SCHED_RANDOMLY ? std::max<uint64_t>(id_bound, 1ull << (2 * sched_half_bits(id_bound))) : id_bound;
}
/* Schedulers.randomly: room for the order keys of `capacity` agents (call before make_view binds v.okey) ... */
static int order_keys_room(abm_handle* h, uint64_t capacity) {
  return h->cfg.scheduler == ABM_SCHED_RANDOMLY ? ensure(h, h->okey, (capacity + 1) * 4) : 0;
}
/* ... and the keys of agents [first, first + count) of the current SoA for the tick that is about to run */
static void order_keys(abm_handle* h, uint64_t first, uint64_t count) {
  if (h->cfg.scheduler != ABM_SCHED_RANDOMLY || count == 0) return;
  OrderKeyBody ok;
  ok.sk = sched_key(h, h->tick, h->next_id); ok.id = (const uint32_t*)h->id[h->cur].p + first; ok.okey = (uint32_t*)h->okey.p + first;
  launch_foreach(h->stream, ok, count);
}

static View make_view(abm_handle* h) {
  View v;
  memset(&v, 0, sizeof v);
  v.g = h->g;
  v.p = h->params;
  v.p.tick = h->tick;
  v.p.sk = sched_key(h, h->tick, h->next_id);
  v.n = (uint32_t)h->n;
  v.cell = (uint32_t*)h->cell[h->cur].p;
  v.id = (uint32_t*)h->id[h->cur].p;
  v.okey = v.p.sk.on ? (const uint32_t*)h->okey.p : v.id;
  v.energy = (double*)h->energy[h->cur].p;
  v.state = (uint8_t*)h->state.p;
  v.slot = (uint32_t*)h->slot.p;
  v.cell_start = (const uint32_t*)h->cell_start[h->cur].p;
  v.grass = (uint8_t*)h->grass.p;
  v.walkbits = (const uint32_t*)h->walkbits.p;
  v.lut = (const double*)h->lut.p;
  v.greg_order = (const uint8_t*)h->greg_order.p;
  v.greg_table = (const double*)h->greg_table.p;
  v.greg_dxy = (const int8_t*)h->greg_dxy.p;
  v.err = (uint32_t*)h->counters.p + CTR_ERR;
  if (h->cfg.walk_type == ABM_ALTERNATED_WALK) {
    v.remote_head = (uint32_t*)h->remote_head.p;
    v.remote_next = (uint32_t*)h->remote_next.p;
    v.remote_dest = (uint32_t*)h->remote_dest.p;
  }
  return v;
}

static int read_counters(abm_handle* h) {
  ABM_CK(h, copy_d2h(h->h_counters, h->counters.p, CTR_WORDS * sizeof(uint32_t), h->stream));
  return 0;
}

static int device_error_to_code(abm_handle* h) {
  const uint32_t e = h->h_counters[CTR_ERR];
  if (e & ERR_NOWALKABLE) ABM_FAIL(h, ABM_E_NOWALKABLE, "random_walkmap: an agent has no walkable neighbour (the reference throws on rand(1:0))");
  if (e & ERR_P2P_TIMEOUT) { h->poisoned = true; ABM_FAIL(h, ABM_E_CUDA, "peer exchange timed out: a neighbouring strip did not deliver its message (the kernels went on without it: this handle is unusable now)"); }
  if (e & ERR_STRIP_CAP) ABM_FAIL(h, ABM_E_CAPACITY, "a boundary row outgrew the agreed halo message capacity (%llu agents) within one tick", (unsigned long long)h->bcap);
  if (e & ERR_INTERNAL) ABM_FAIL(h, ABM_E_BADARG, "device-side validation failed (out-of-range coordinate, id, countdown or elevation, or duplicate id)");
  return 0;
}

static uint64_t bitmap_words(uint64_t next_id) { return (next_id >> 5) + 1; }

/* exclusive popcount prefix of h->bits over `words` words into h->bits_prefix */
static int rank_prefix(abm_handle* h, uint64_t words) {
  int rc;
  if ((rc = ensure(h, h->bits_prefix, (words + 4) * 4))) return rc;
  if ((rc = ensure(h, h->scan_tmp, scan_tmp_words(std::max<uint64_t>(words, (uint64_t)h->g.gt + 1)) * 4))) return rc;
  PopcountBody pb;
  pb.bits = (const uint32_t*)h->bits.p;
  pb.out = (uint32_t*)h->bits_prefix.p;
  launch_foreach(h->stream, pb, words);
  ABM_CK(h, scan_exclusive_u32(h->stream, (uint32_t*)h->bits_prefix.p, words, (uint32_t*)h->scan_tmp.p));
  return 0;
}

static int ensure_agent_scratch(abm_handle* h, uint64_t n) {
  int rc;
  const uint64_t m = std::max<uint64_t>(n, 1);
  if ((rc = ensure(h, h->state, m))) return rc;
  if ((rc = ensure(h, h->slot, m * 4))) return rc;
  if ((rc = ensure(h, h->new_cell, m * 4))) return rc;
  if ((rc = ensure(h, h->pend[0], m * 4))) return rc;
  if ((rc = ensure(h, h->pend[1], m * 4))) return rc;
  if ((rc = ensure(h, h->births, m * sizeof(BirthRec)))) return rc;
  return 0;
}

static int ensure_soa(abm_handle* h, int which, uint64_t n) {
  int rc;
  const uint64_t m = std::max<uint64_t>(n, 1);
  if ((rc = ensure(h, h->cell[which], m * 4))) return rc;
  if ((rc = ensure(h, h->id[which], m * 4))) return rc;
  if ((rc = ensure(h, h->energy[which], m * 8))) return rc;
  return 0;
}


/* grows a buffer keeping its contents (new bytes zeroed) */
static int grow_keep(abm_handle* h, DevBuf& b, size_t bytes) {
  bytes = (bytes + 255) & ~(size_t)255;
  if (b.bytes >= bytes && b.p) return 0;
  void* np_ = nullptr;
  const size_t want = bytes + bytes / 2;
  if (dev_malloc(&np_, want) != 0) ABM_FAIL(h, ABM_E_NOMEM, "device allocation of %zu bytes failed", want);
  if (b.p && b.bytes) {
    abm_err_t e = copy_d2d(np_, b.p, b.bytes, h->stream);
    if (e == 0) e = stream_sync(h->stream);
    if (e != 0) { dev_free(np_); ABM_FAIL(h, ABM_E_CUDA, "device copy failed: %s", err_string(e)); }
  }
  dev_free(b.p);
  b.p = np_; b.bytes = want;
  return 0;
}

static RouteStore route_store(abm_handle* h) {
  RouteStore rs;
  rs.slot_of_id = (uint32_t*)h->rt_slot_of_id.p; rs.r_off = (uint64_t*)h->rt_off.p; rs.r_len = (uint32_t*)h->rt_len.p;
  rs.r_pos = (uint32_t*)h->rt_pos.p; rs.pool = (uint32_t*)h->rt_pool.p;
  return rs;
}

static int ensure_route_ids(abm_handle* h, uint64_t ids) {
  if (ids <= h->rt_ids_cap) return 0;
  int rc = grow_keep(h, h->rt_slot_of_id, (ids + 1) * 4);
  if (rc) return rc;
  h->rt_ids_cap = h->rt_slot_of_id.bytes / 4 - 1;
  return 0;
}

static AStarGrid astar_grid(abm_handle* h) {
  AStarGrid a;
  a.g = h->g; a.walkbits = (const uint32_t*)h->walkbits.p; a.elev = (const int16_t*)h->elev.p;
  a.metric = h->cfg.astar_metric; a.penalty = h->cfg.astar_penalty;
  return a;
}

static size_t device_free_bytes() {
#ifndef ABM_EMU
  size_t f = 0, t = 0;
  if (cudaMemGetInfo(&f, &t) != cudaSuccess) return (size_t)1 << 30;
  return f;
#else
  return (size_t)1 << 30;
#endif
}

#include "abm_host_routes.inl"

/* one tick: every agent's custom_agent_step! in id order, then custom_model_step! */
static int tick_once(abm_handle* h) {
  int rc;
  const uint64_t n = h->n;
  const uint64_t gt = h->g.gt;
  if ((rc = ensure_agent_scratch(h, n))) return rc;
  uint32_t* ctr = (uint32_t*)h->counters.p;
  ABM_CK(h, dev_memset(ctr, 0, 3 * sizeof(uint32_t), h->stream)); /* keeps the sticky error word */
  if ((rc = order_keys_room(h, n))) return rc;
  order_keys(h, 0, n);
  View v = make_view(h);

  if (n > 0 && h->cfg.walk_type == ABM_ALTERNATED_WALK) {
    if ((rc = alternated_propose(h, v, ctr))) return rc;
  } else if (n > 0) {
    ProposeBody pb;
    pb.v = v;
    pb.pending = (uint32_t*)h->pend[0].p;
    pb.pending_count = ctr + CTR_PEND_A;
    pb.birth_flags = nullptr; pb.parent_list = nullptr; pb.parent_cap = 0;
    { ProfScope ps(h->prof, h->stream, "propose", n); launch_foreach(h->stream, pb, n); }
    if ((rc = check_launch(h, "propose"))) return rc;
  }
  /* rank-ordered fixed point over the agents whose outcome depends on lower ids */
  const bool may_pend = h->cfg.walk_type != ABM_RANDOM_WALKMAP && n > 0;
  if (may_pend) {
    if ((rc = read_counters(h))) return rc;
    uint64_t cnt = h->h_counters[CTR_PEND_A];
    int in = 0;
    int stalls = 0;
    while (cnt > 0) {
      ABM_CK(h, dev_memset(ctr + (in ? CTR_PEND_A : CTR_PEND_B), 0, sizeof(uint32_t), h->stream));
      ResolveBody rb;
      rb.v = v;
      rb.in = (const uint32_t*)h->pend[in].p;
      rb.in_count = ctr + (in ? CTR_PEND_B : CTR_PEND_A);
      rb.out = (uint32_t*)h->pend[in ^ 1].p;
      rb.out_count = ctr + (in ? CTR_PEND_A : CTR_PEND_B);
      { ProfScope ps(h->prof, h->stream, "resolve", cnt); launch_foreach(h->stream, rb, cnt); }
      if ((rc = check_launch(h, "resolve"))) return rc;
      if ((rc = read_counters(h))) return rc;
      const uint64_t left = h->h_counters[in ? CTR_PEND_A : CTR_PEND_B];
      h->rounds += 1;
      /* the lowest pending id is always decidable, so a round without progress can only be a stale read */
      stalls = left == cnt ? stalls + 1 : 0;
      if (stalls > 8) ABM_FAIL(h, ABM_E_INTERNAL, "conflict resolution made no progress with %llu agents pending", (unsigned long long)left);
      cnt = left;
      in ^= 1;
    }
  }

  /* commit: eat / energy / reproduce and the counting sort into next tick's order */
  const int nxt = h->cur ^ 1;
  uint32_t* new_start = (uint32_t*)h->cell_start[nxt].p;
  { ProfScope ps(h->prof, h->stream, "memset_cell_counts", gt + 1); ABM_CK(h, dev_memset(new_start, 0, (gt + 1) * 4, h->stream)); }
  const uint64_t words = bitmap_words(key_space(h, h->next_id));
  const bool repro = h->cfg.reproduce != 0;
  if (repro) {
    if ((rc = ensure(h, h->bits, (words + 4) * 4))) return rc;
    ABM_CK(h, dev_memset(h->bits.p, 0, words * 4, h->stream));
  }
  if (n > 0) {
    CommitBody cb;
    cb.v = v;
    cb.new_count = new_start;
    cb.birth_bits = (uint32_t*)h->bits.p;
    cb.births = (BirthRec*)h->births.p;
    cb.birth_count = ctr + CTR_BIRTHS;
    cb.new_cell = (uint32_t*)h->new_cell.p;
    { ProfScope ps(h->prof, h->stream, "commit", n); launch_foreach(h->stream, cb, n); }
    if ((rc = check_launch(h, "commit"))) return rc;
  }
  if ((rc = ensure(h, h->scan_tmp, scan_tmp_words(gt + 1) * 4))) return rc;
  { ProfScope ps(h->prof, h->stream, "cell_scan", gt + 1, 3); ABM_CK(h, scan_exclusive_u32(h->stream, new_start, gt + 1, (uint32_t*)h->scan_tmp.p)); }
  if ((rc = read_counters(h))) return rc;
  if ((rc = device_error_to_code(h))) return rc;
  uint32_t n_new32 = 0;
  ABM_CK(h, copy_d2h(&n_new32, new_start + gt, 4, h->stream));
  const uint64_t n_new = n_new32;
  const uint64_t nb = h->h_counters[CTR_BIRTHS];
  if (h->next_id + nb > 0xFFFFFFFFull) ABM_FAIL(h, ABM_E_CAPACITY, "agent id space (2^32-1) exhausted");
  if ((rc = ensure_soa(h, nxt, n_new))) return rc;
  if (n > 0) {
    ScatterBody sb;
    sb.v = v;
    sb.new_cell = (const uint32_t*)h->new_cell.p;
    sb.new_start = new_start;
    sb.out_cell = (uint32_t*)h->cell[nxt].p;
    sb.out_id = (uint32_t*)h->id[nxt].p;
    sb.out_energy = (double*)h->energy[nxt].p;
    { ProfScope ps(h->prof, h->stream, "scatter", n); launch_foreach(h->stream, sb, n); }
    if ((rc = check_launch(h, "scatter"))) return rc;
    if (nb > 0) {
      if ((rc = rank_prefix(h, words))) return rc;
      BirthScatterBody bb;
      bb.v = v;
      bb.births = (const BirthRec*)h->births.p;
      bb.birth_bits = (const uint32_t*)h->bits.p;
      bb.birth_prefix = (const uint32_t*)h->bits_prefix.p;
      bb.new_cell = sb.new_cell;
      bb.new_start = new_start;
      bb.next_id = (uint32_t)h->next_id;
      bb.out_cell = sb.out_cell;
      bb.out_id = sb.out_id;
      bb.out_energy = sb.out_energy;
      { ProfScope ps(h->prof, h->stream, "birth_scatter", nb); launch_foreach(h->stream, bb, nb); }
      if ((rc = check_launch(h, "birth scatter"))) return rc;
    }
  }
  if (n > 0 && h->cfg.walk_type == ABM_ALTERNATED_WALK) {
    RemoteResetBody rr;
    rr.v = v;
    launch_foreach(h->stream, rr, n);
  }
  /* custom_model_step!: grass regrowth */
  {
    RegrowBody gb;
    gb.grass4 = (uint32_t*)h->grass.p;
    gb.walkbits = (const uint32_t*)h->walkbits.p;
    gb.rt = (uint32_t)h->cfg.regrowth_time;
    gb.use_walk = h->cfg.walkmap_regrowth;
    { ProfScope ps(h->prof, h->stream, "regrow", gt); launch_foreach(h->stream, gb, gt / 16); }
    if ((rc = check_launch(h, "regrow"))) return rc;
  }
  h->cur = nxt;
  h->n = n_new;
  h->next_id += nb;
  h->tick += 1;
  h->updates += (double)n;
  return 0;
}

static SubIndex sub_index(abm_handle* h, int which) {
  SubIndex si;
  si.off = (const uint32_t*)h->sub_off[which].p; si.cnt = (const uint32_t*)h->sub_cnt[which].p; si.nsx = h->nsx; si.nsy = h->nsy;
  return si;
}

#include "abm_host_strips.inl"

#include "abm_host_stream.inl"

/* The offspring ids' ranks, prepared on the side stream while the tick resolves (tick_tile).  In: [count][order keys] of this
 * handle's reproducing agents (ProposeBody; count = CTR_BIRTH_FLAGS).  Out: every strip's list side by side in h->birth_ids,
 * the bitmap of all of them over the key space in h->bits and its popcount prefix in h->bits_prefix; ev_ranked is recorded
 * behind the last kernel.  Strips exchange the lists through their peer-memory mailboxes (the birth channel). */
struct ParentCountBody {
  const uint32_t* count; uint32_t cap; uint32_t* msg;
  ABM_HD void operator()(uint64_t) const { const uint32_t c = *count; msg[0] = c < cap ? c : cap; }
};
static int early_births_enqueue(abm_handle* h, uint64_t cap, uint64_t words) {
  const int P = h->n_ranks;
  uint32_t* ctr = (uint32_t*)h->counters.p;
  uint32_t* msg = (uint32_t*)h->parents.p;
  uint32_t* all = (uint32_t*)h->birth_ids.p;
  const abm_stream_t s = h->side;
  event_record(h->ev_proposed, h->stream);
  stream_wait_event(s, h->ev_proposed);
  bool gathered = false;
#ifndef ABM_EMU
  if (h->strip && h->p2p_on && P > 1) {
    const uint32_t seq = ++h->p2p_seq[P2P_CH_BIRTHS];
    const int parity = (int)(seq & 1u);
    P2PBirthSendCta sd;
    sd.birth_parent = msg + 1; sd.birth_count = ctr + CTR_BIRTH_FLAGS; sd.cap = (uint32_t)cap; sd.seq = seq; sd.done = ctr + CTR_BIRTH_DONE;
    P2PBirthRecvCta rv;
    rv.seq = seq; rv.cap = (uint32_t)cap; rv.all = all; rv.err = ctr + CTR_ERR;
    for (int r = 0; r < P; ++r) {
      sd.dst[r] = (uint32_t*)(h->p2p_peer[r] + h->p2p_lay.births_off(h->my_rank, parity));
      sd.flag[r] = (uint32_t*)(h->p2p_peer[r] + h->p2p_lay.flag_off(P2P_CH_BIRTHS, h->my_rank, parity));
      rv.src[r] = (const uint32_t*)(h->p2p_mine + h->p2p_lay.births_off(r, parity));
      rv.flag[r] = (const uint32_t*)(h->p2p_mine + h->p2p_lay.flag_off(P2P_CH_BIRTHS, r, parity));
    }
    launch_cta(s, sd, (uint64_t)P * BIRTH_CTAS, 256, 16);
    launch_cta(s, rv, (uint64_t)P * BIRTH_CTAS, 256, 16);
    gathered = true;
  }
#endif
  if (!gathered) { /* a single handle (or a ring of one): its own list is the whole gathering */
    ParentCountBody pc;
    pc.count = ctr + CTR_BIRTH_FLAGS; pc.cap = (uint32_t)cap; pc.msg = msg;
    launch_foreach(s, pc, 1);
    ABM_CK(h, copy_d2d(all, msg, (1 + cap) * 4, s));
  }
  ABM_CK(h, dev_memset(h->bits.p, 0, words * 4, s));
  BirthMarkBody bm;
  bm.all = all; bm.n_ranks = P; bm.cap = (uint32_t)cap; bm.bits = (uint32_t*)h->bits.p;
  launch_foreach(s, bm, (uint64_t)P * std::max<uint64_t>(cap, 1));
  PopcountBody pb;
  pb.bits = (const uint32_t*)h->bits.p; pb.out = (uint32_t*)h->bits_prefix.p;
  launch_foreach(s, pb, words);
  ABM_CK(h, scan_exclusive_u32(s, (uint32_t*)h->bits_prefix.p, words, (uint32_t*)h->scan_tmp.p));
  event_record(h->ev_ranked, s);
  return check_launch(h, "early offspring ranks");
}

/* one tick on the tile path (abm_tile.h): propose -> window resolve -> per-agent windows for the rest -> commit */
static int tick_tile(abm_handle* h) {
  int rc;
  const uint64_t n = h->n;
  const uint64_t cores = (uint64_t)(h->g.nx / CORE) * (uint64_t)((h->g.own_hi - h->g.own_lo) / CORE);
  const uint64_t m = std::max<uint64_t>(n, 1);
  if ((rc = ensure(h, h->state, m))) return rc;
  if ((rc = ensure(h, h->pend[0], m * 4))) return rc;
  if ((rc = ensure(h, h->pend[1], m * 4))) return rc;
  uint32_t* ctr = (uint32_t*)h->counters.p;
  ABM_CK(h, dev_memset(ctr, 0, 3 * sizeof(uint32_t), h->stream));
  ABM_CK(h, dev_memset(ctr + CTR_CURSOR, 0, 2 * sizeof(uint32_t), h->stream));
  if (h->strip) {
    if (!h->bcap_agreed) { /* first tick after abm_set_agents: agree on the message capacity from the real row sizes */
      if ((rc = strip_alloc(h))) return rc;
      const int row_s[2] = {h->g.own_lo / SUB, h->g.own_hi / SUB - 1};
      for (int d = 0; d < 2; ++d) {
        GatherCountsBody gc;
        gc.sub_cnt = (const uint32_t*)h->sub_cnt[h->cur].p; gc.row = row_s[d]; gc.nsx = h->nsx; gc.out = (uint32_t*)h->sx_pre[d].p;
        launch_foreach(h->stream, gc, (uint64_t)h->nsx + 1);
        ABM_CK(h, scan_exclusive_u32(h->stream, (uint32_t*)h->sx_pre[d].p, (uint64_t)h->nsx + 1, (uint32_t*)h->scan_tmp.p));
      }
      std::vector<uint64_t> st0;
      if ((rc = strip_stats(h, ctr + CTR_PEND_A, st0))) return rc;
      uint64_t rmax = 0;
      for (int q = 0; q < h->n_ranks; ++q) rmax = std::max(rmax, st0[1 + 2 * h->n_ranks + q]);
      h->bcap = h->next_bcap = strip_cap_for(rmax);
      h->bcap_agreed = true;
    }
    if ((rc = strip_prepare(h))) return rc; /* before anybody takes pointers */
  }
  if ((rc = order_keys_room(h, n + (h->strip ? 2 * h->bcap : 0)))) return rc;
  order_keys(h, 0, n);
  View v = make_view(h);
  const SubIndex si = sub_index(h, h->cur);
  const int wt = h->cfg.walk_type;
  const bool needs_resolve = wt == ABM_RANDOM_WALK_ATTRACTOR || wt == ABM_RANDOM_WALK_GREGARIOUS || (wt == ABM_RANDOM_WALK && h->cfg.randomwalk_ifempty);
  uint64_t cnt = 0;
  /* Will this tick commit behind the device-side guard?  Then the offspring ids can be prepared early: every reproducing
   * agent is known after propose, so the parents' order keys are gathered from all strips and ranked (bitmap over the key
   * space + popcount prefix) on the side stream while the tick resolves and commits; what stays behind the commit is one
   * small kernel.  Peer-memory strips and single handles only (an NCCL collective must not run beside the main stream's). */
  const bool optimistic = h->optimistic && h->bf_est > 0 && (!h->strip || strip_stats_on_device(h));
  bool early = optimistic && h->early_births && h->cfg.reproduce && (!h->strip || h->n_ranks == 1);
#ifndef ABM_EMU
  if (optimistic && h->early_births && h->cfg.reproduce && h->strip && h->p2p_on) early = true;
#endif
  const int P_all = h->n_ranks;
  const uint64_t ecap = h->bf_est, ewords = bitmap_words(key_space(h, h->next_id));
  if (early) { /* every buffer the side stream touches exists before anything is enqueued */
    if ((rc = ensure(h, h->parents, (ecap + 2) * 4))) return rc;
    if ((rc = ensure(h, h->birth_ids, (1 + ecap) * 4 * (size_t)(P_all + 2) + 64))) return rc;
    if ((rc = ensure(h, h->bits, (ewords + 4) * 4))) return rc;
    if ((rc = ensure(h, h->bits_prefix, (ewords + 4) * 4))) return rc;
    if ((rc = ensure(h, h->scan_tmp, scan_tmp_words(std::max<uint64_t>(ewords, (uint64_t)h->g.gt + 1)) * 4))) return rc;
#ifndef ABM_EMU
    if (h->strip && h->p2p_on && (1 + ecap) * 4 > h->p2p_lay.births_bytes) early = false; /* the mailbox is smaller: gather after the commit */
#endif
  }
  /* peer-memory strips: the two boundary rows are proposed first and travel (side stream) while propose does the rest */
  bool halo_beside = false;
#ifndef ABM_EMU
  halo_beside = h->strip && h->p2p_on && h->n_ranks > 1 && h->overlap_halo;
#endif
  if (halo_beside) {
    ProposeRowsCta pr;
    pr.v = v; pr.si = si; pr.row[0] = h->g.own_lo / SUB; pr.row[1] = h->g.own_hi / SUB - 1;
    pr.chunks = (h->nsx + TILE_THREADS / 32 - 1) / (TILE_THREADS / 32);
    launch_cta(h->stream, pr, 2 * (uint64_t)pr.chunks, TILE_THREADS, 16);
    event_record(h->ev_rows, h->stream);
    stream_wait_event(h->side, h->ev_rows);
    std::swap(h->stream, h->side); /* the exchange (and its profiling scope) is enqueued on the side stream */
    rc = strip_exchange_agents(h);
    if (!rc) order_keys(h, n, 2 * h->bcap);
    event_record(h->ev_halo, h->stream);
    std::swap(h->stream, h->side);
    if (rc) return rc;
  }
  {
    ProposeBody pb;
    pb.v = v; pb.pending = nullptr; pb.pending_count = nullptr; pb.birth_flags = ctr + CTR_BIRTH_FLAGS;
    pb.parent_list = early ? (uint32_t*)h->parents.p + 1 : nullptr; pb.parent_cap = (uint32_t)ecap;
    if (n > 0) {
      { ProfScope ps(h->prof, h->stream, "propose", n); launch_foreach(h->stream, pb, n); }
      if ((rc = check_launch(h, "propose"))) return rc;
    }
  }
  if (early && (rc = early_births_enqueue(h, ecap, ewords))) return rc;
  if (halo_beside) stream_wait_event(h->stream, h->ev_halo); /* the ghost rows are in place before anything looks across the border */
  const int cores_x = h->g.nx / CORE, core_rows = (h->g.own_hi - h->g.own_lo) / CORE;
  auto launch_tiles = [&](int row0, int nrows, int row_step = 1) -> int { /* window resolve of the core rows row0, row0 + step, ... */
    if (nrows <= 0) return 0;
    ResolveWindow rw;
    rw.v = v; rw.si = si; rw.mode = 0; rw.cores_x = cores_x; rw.core_y0 = h->g.own_lo / CORE + row0; rw.row_step = row_step; rw.max_rounds = h->tile_rounds; rw.nthreads = TILE_THREADS;
    rw.margin = std::max<int>(h->tile_margin, wt == ABM_RANDOM_WALK_GREGARIOUS ? std::min<int>(h->cfg.visual_distance + 1, HALO) : 2);
    rw.in = nullptr; rw.out = (uint32_t*)h->pend[0].p; rw.out_count = ctr + CTR_PEND_A;
    const uint64_t grid = (uint64_t)nrows * (uint64_t)cores_x;
    { ProfScope ps(h->prof, h->stream, "resolve_tiles", grid * 256, 1); launch_cta(h->stream, rw, grid, TILE_THREADS, ResolveWindow::smem_bytes_core()); }
    return check_launch(h, "tile resolve");
  };
  if (h->strip && !halo_beside && (rc = strip_exchange_agents(h))) return rc; /* neighbours' boundary rows become ghost rows behind the own agents */
  if (h->strip && !halo_beside) order_keys(h, n, 2 * h->bcap); /* the ghosts' order keys (slots beyond the real counts are never read) */
  if (n + h->n_ghost > 0 && needs_resolve && (rc = launch_tiles(0, core_rows))) return rc;
  /* (measured and dropped: resolving the two boundary core rows first and shipping their decisions beside the interior
   * rows' resolve — the two 128-CTA launches cost more than the exchange they hide: 0.706 vs 0.666 ms at 2 GPUs) */
  /* figures the host needs before it can go on: open agents, reproducing agents (sizes of the output SoA); in strip
   * mode from every rank, through one all-reduce of device-resident counters */
  uint64_t birth_flags = 0, birth_cap = 0, global_cnt = 0, strip_fmax = 0;
  std::vector<uint64_t> stats;
  const int P = h->n_ranks;
  auto take_stats = [&](const uint32_t* d_pending) -> int {
    if (!h->strip) {
      int r = read_counters(h);
      if (r) return r;
      cnt = h->h_counters[d_pending - ctr];
      global_cnt = cnt;
      birth_flags = h->h_counters[CTR_BIRTH_FLAGS];
      return 0;
    }
    int r = strip_stats(h, d_pending, stats, needs_resolve);
    if (r) return r;
    global_cnt = stats[0];
    cnt = stats[1 + h->my_rank];
    birth_flags = stats[1 + P + h->my_rank];
    uint64_t fmax = 0, rmax = 0;
    for (int q = 0; q < P; ++q) { fmax = std::max(fmax, stats[1 + P + q]); rmax = std::max(rmax, stats[1 + 2 * P + q]); }
    birth_cap = 3 * fmax + 64; /* a rank commits its own and its neighbours' arriving parents */
    h->next_bcap = rmax * 2 > h->bcap ? strip_cap_for(rmax) : h->bcap; /* every rank sees the same figures */
    return 0;
  };
  int in = 0;
  const int q = wt == ABM_RANDOM_WALK_GREGARIOUS ? h->cfg.visual_distance + 1 : 2;
  auto launch_agents = [&](uint64_t grid, uint32_t* d_out) -> int { /* one Jacobi round of per-agent windows over pend[in] */
    ABM_CK(h, dev_memset(d_out, 0, sizeof(uint32_t), h->stream));
    ResolveWindow rw;
    rw.v = v; rw.si = si; rw.mode = 1; rw.cores_x = 0; rw.core_y0 = 0; rw.row_step = 1; rw.max_rounds = 1; rw.nthreads = AGENT_THREADS; rw.margin = HALO;
    rw.grid = (uint32_t)grid; rw.in = (const uint32_t*)h->pend[in].p; rw.in_count = ctr + (in ? CTR_PEND_B : CTR_PEND_A);
    rw.out = (uint32_t*)h->pend[in ^ 1].p; rw.out_count = d_out;
    { ProfScope ps(h->prof, h->stream, "resolve_agents", grid); launch_cta(h->stream, rw, grid, AGENT_THREADS, ResolveWindow::smem_bytes(2 * q + 1, 2 * q + 1)); }
    return check_launch(h, "agent resolve");
  };
  if (needs_resolve && n + h->n_ghost > 0) {
    /* the first rounds of per-agent windows are enqueued blind: their list lengths stay on the device, the launch is an
     * upper bound learnt from the previous tick (grid-stride inside), and the host reads everything back once */
    ABM_CK(h, copy_d2d(ctr + CTR_PEND_FIRST, ctr + CTR_PEND_A, sizeof(uint32_t), h->stream));
    for (int r = 0; r < h->fixed_agent_rounds; ++r) {
      /* strips: the owners' decisions reach the ghost copies before the first blind round and once more with the counters
       * below (and before the last of three or more rounds); an agent that waits for a ghost in between simply stays
       * open and the guard sees it */
      if (h->strip && (r == 0 || (r >= 2 && r == h->fixed_agent_rounds - 1)) && (rc = strip_exchange_states(h))) return rc;
      if ((rc = launch_agents(std::max<uint64_t>(h->kf_grid, 1), ctr + (in ? CTR_PEND_A : CTR_PEND_B)))) return rc;
      in ^= 1;
      h->rounds += 1;
    }
  }
  const int nxt = h->cur ^ 1;
  const uint64_t words = bitmap_words(key_space(h, h->next_id));
  /* commit + grass + re-grouping (+ offspring ids on one GPU); `go` = device flag checked by the kernels, or null */
  auto commit_phase = [&](uint64_t bflags, const uint32_t* go) -> int {
    const uint64_t arrivals_cap = h->strip ? 2 * h->bcap : 0; /* ghost agents (and their offspring) may land here */
    int r;
    if ((r = ensure_soa(h, nxt, n + bflags + 2 * arrivals_cap))) return r;
    if ((r = ensure(h, h->birth_pos, (bflags + arrivals_cap + 1) * 4))) return r;
    if ((r = ensure(h, h->birth_parent, (bflags + arrivals_cap + 1) * 4))) return r;
    if ((bflags > 0 || h->strip) && !(early && go)) {
      if ((r = ensure(h, h->bits, (words + 4) * 4))) return r;
      if (!h->strip) ABM_CK(h, dev_memset(h->bits.p, 0, words * 4, h->stream));
    }
    CommitTile ct;
    ct.mig_cnt = nullptr; ct.mig_a = nullptr; ct.mig_lc = nullptr;
    if (h->use_migrants) { /* decisions are final here (or `go` is 0): list the agents that cross a core border at their destination */
      if ((r = ensure(h, h->mig_cnt, cores * 4))) return r;
      if ((r = ensure(h, h->mig_a, cores * MIG_CAP * 4))) return r;
      if ((r = ensure(h, h->mig_lc, cores * MIG_CAP * 2))) return r;
      ABM_CK(h, dev_memset(h->mig_cnt.p, 0, cores * 4, h->stream));
      MigrantScanBody ms;
      ms.g = h->g; ms.cores_x = h->g.nx / CORE; ms.core_y0 = h->g.own_lo / CORE; ms.core_rows = (h->g.own_hi - h->g.own_lo) / CORE;
      ms.cell = (const uint32_t*)h->cell[h->cur].p; ms.state = (const uint8_t*)h->state.p;
      ms.mig_cnt = (uint32_t*)h->mig_cnt.p; ms.mig_a = (uint32_t*)h->mig_a.p; ms.mig_lc = (uint16_t*)h->mig_lc.p;
      ms.go = go; ms.n_own = (uint32_t)n; ms.ghost_cap = (uint32_t)h->bcap;
      ms.ghost_total[0] = h->strip ? (const uint32_t*)h->rx_pre[0].p + h->nsx : nullptr;
      ms.ghost_total[1] = h->strip ? (const uint32_t*)h->rx_pre[1].p + h->nsx : nullptr;
      { ProfScope ps(h->prof, h->stream, "migrant_scan", n + h->n_ghost); launch_foreach(h->stream, ms, n + h->n_ghost); }
      ct.mig_cnt = ms.mig_cnt; ct.mig_a = ms.mig_a; ct.mig_lc = ms.mig_lc;
    }
    ct.v = v; ct.si = si; ct.cores_x = h->g.nx / CORE; ct.core_y0 = h->g.own_lo / CORE;
    ct.out_cell = (uint32_t*)h->cell[nxt].p; ct.out_id = (uint32_t*)h->id[nxt].p; ct.out_energy = (double*)h->energy[nxt].p;
    ct.new_off = (uint32_t*)h->sub_off[nxt].p; ct.new_cnt = (uint32_t*)h->sub_cnt[nxt].p;
    ct.cursor = ctr + CTR_CURSOR;
    ct.birth_pos = (uint32_t*)h->birth_pos.p; ct.birth_parent = (uint32_t*)h->birth_parent.p; ct.birth_count = ctr + CTR_BIRTHS;
    ct.birth_bits = (h->strip || (early && go)) ? nullptr : (uint32_t*)h->bits.p; /* strips (and the early gather) rank the parents from their lists */
    ct.go = go;
    { ProfScope ps(h->prof, h->stream, "commit_tiles", n); launch_cta(h->stream, ct, cores, COMMIT_THREADS, CommitTile::smem_bytes()); }
    return check_launch(h, "tile commit");
  };
  auto finish_tick = [&](uint64_t n_new, uint64_t nb_total) -> int {
    if (h->next_id + nb_total > 0xFFFFFFFFull) ABM_FAIL(h, ABM_E_CAPACITY, "agent id space (2^32-1) exhausted");
    h->kf_grid = std::min<uint64_t>(std::max<uint64_t>(2ull * h->h_counters[CTR_PEND_FIRST] + 256, 512), std::max<uint64_t>(n + h->n_ghost, 1));
    h->bf_est = 2ull * h->h_counters[CTR_BIRTH_FLAGS] + 1024;
    h->cur = nxt; h->n = n_new; h->next_id += nb_total; h->n_ghost = 0; h->tick += 1; h->updates += (double)n;
    return 0;
  };
  if (optimistic) {
    /* steady state: no read-back before the commit.  A one-thread guard tells the commit and the offspring kernels
     * whether everybody is decided and the output arrays (sized from last tick's births) are large enough; when not,
     * they do nothing and the careful path below redoes this part with exact figures.  Strips: the guard reads the
     * all-reduced counters, so every rank decides alike and the collectives stay matched. */
    const uint64_t bcapn = h->bf_est;
    uint32_t* d_pend = ctr + (in ? CTR_PEND_B : CTR_PEND_A);
    if (h->strip) {
      const uint32_t gcap = (uint32_t)std::min<uint64_t>(bcapn, 0xFFFFFFFFull);
      bool guarded = false;
#ifndef ABM_EMU
      guarded = h->p2p_on;
#endif
      if ((rc = strip_stats_enqueue(h, d_pend, needs_resolve, ctr + CTR_GO, gcap))) return rc;
      if (!guarded) {
        StripGuardBody sg;
        sg.stats = (const unsigned long long*)h->red_buf.p; sg.n_ranks = P; sg.birth_cap = gcap; sg.go = ctr + CTR_GO;
        launch_foreach(h->stream, sg, 1);
      }
    } else {
      TickGuardBody tg;
      tg.pending = d_pend; tg.birth_flags = ctr + CTR_BIRTH_FLAGS; tg.birth_cap = (uint32_t)std::min<uint64_t>(bcapn, 0xFFFFFFFFull); tg.go = ctr + CTR_GO;
      launch_foreach(h->stream, tg, 1);
    }
    if ((rc = commit_phase(bcapn, ctr + CTR_GO))) return rc;
    if (early) { /* the ranks are ready (or about to be): offspring ids = next_id + rank of the parent's order key */
      stream_wait_event(h->stream, h->ev_ranked);
      const uint64_t arrivals_cap = h->strip ? 2 * h->bcap : 0;
      BirthAssignBody ba;
      ba.all = (const uint32_t*)h->birth_ids.p; ba.n_ranks = P; ba.cap = (uint32_t)ecap; ba.next_id = (uint32_t)h->next_id;
      ba.birth_pos = (const uint32_t*)h->birth_pos.p; ba.birth_parent = (const uint32_t*)h->birth_parent.p; ba.birth_count = ctr + CTR_BIRTHS;
      ba.rank_in = nullptr; ba.birth_bits = (const uint32_t*)h->bits.p; ba.birth_prefix = (const uint32_t*)h->bits_prefix.p;
      ba.out_id = (uint32_t*)h->id[nxt].p; ba.total_out = ctr + CTR_BIRTH_TOTAL;
      { ProfScope ps(h->prof, h->stream, "birth_ids", bcapn); launch_foreach(h->stream, ba, bcapn + arrivals_cap + 1); }
      if ((rc = check_launch(h, "birth ids"))) return rc;
      if (h->strip) {
        birth_cap = 3 * bcapn + 64;
        stats.assign((size_t)(1 + 3 * P), 0);
        ABM_CK(h, copy_d2h(stats.data(), h->red_buf.p, stats.size() * 8, h->stream));
      }
    } else if (h->strip) {
      birth_cap = 3 * bcapn + 64;
      if ((rc = strip_birth_ids(h, birth_cap, nxt, ctr + CTR_BIRTH_TOTAL))) return rc;
      stats.assign((size_t)(1 + 3 * P), 0);
      ABM_CK(h, copy_d2h(stats.data(), h->red_buf.p, stats.size() * 8, h->stream));
    } else if (h->cfg.reproduce) {
      if ((rc = rank_prefix(h, words))) return rc;
      BirthIdBody bi;
      bi.birth_pos = (const uint32_t*)h->birth_pos.p; bi.birth_parent = (const uint32_t*)h->birth_parent.p;
      bi.birth_bits = (const uint32_t*)h->bits.p; bi.birth_prefix = (const uint32_t*)h->bits_prefix.p;
      bi.next_id = (uint32_t)h->next_id; bi.out_id = (uint32_t*)h->id[nxt].p; bi.birth_count = ctr + CTR_BIRTHS;
      { ProfScope ps(h->prof, h->stream, "birth_ids", bcapn); launch_foreach(h->stream, bi, bcapn); }
      if ((rc = check_launch(h, "birth ids"))) return rc;
    }
    if ((rc = read_counters(h))) return rc;
    if ((rc = device_error_to_code(h))) return rc;
    if (h->h_counters[CTR_GO]) {
      if (needs_resolve && ++h->clean_ticks >= 256) { h->clean_ticks = 0; if (h->fixed_agent_rounds > h->min_agent_rounds) --h->fixed_agent_rounds; }
      if (!h->strip) return finish_tick(h->h_counters[CTR_CURSOR], h->h_counters[CTR_BIRTHS]);
      uint64_t fmax = 0, rmax = 0;
      for (int r = 0; r < P; ++r) { fmax = std::max(fmax, stats[1 + P + r]); rmax = std::max(rmax, stats[1 + 2 * P + r]); }
      if (h->h_counters[CTR_BIRTHS] > birth_cap) ABM_FAIL(h, ABM_E_INTERNAL, "more offspring than the agreed message capacity");
      h->bcap = rmax * 2 > h->bcap ? strip_cap_for(rmax) : h->bcap; /* every rank sees the same figures */
      if ((rc = finish_tick(h->h_counters[CTR_CURSOR], h->h_counters[CTR_BIRTH_TOTAL]))) return rc;
      h->bf_est = 2 * fmax + 1024; /* strips: the estimate must be the same on every rank */
      return 0;
    }
    /* not ready (agents still open, or more births than expected): nothing was touched; go on carefully */
    if (early) stream_wait_event(h->stream, h->ev_ranked); /* the side stream's buffers are about to be reused */
    if (needs_resolve) { h->clean_ticks = 0; if (h->fixed_agent_rounds < 6 && h->fixed_agent_rounds >= 1) ++h->fixed_agent_rounds; }
  }
  if ((rc = take_stats(ctr + (in ? CTR_PEND_B : CTR_PEND_A)))) return rc;
  if (h->strip && (rc = read_counters(h))) return rc;
  h->kf_grid = std::min<uint64_t>(std::max<uint64_t>(2ull * h->h_counters[CTR_PEND_FIRST] + 256, 512), std::max<uint64_t>(n + h->n_ghost, 1));
  {
    int stalls = 0;
    while (true) { /* chains that left their window: one small window per agent, Jacobi rounds */
      /* strips: take_stats also carried the owners' decisions to the ghost copies; every rank stays in the loop until all
       * ranks are done */
      if (global_cnt == 0) break;
      uint32_t* d_out = ctr + (in ? CTR_PEND_A : CTR_PEND_B);
      if (cnt > 0) { if ((rc = launch_agents(cnt, d_out))) return rc; }
      else ABM_CK(h, dev_memset(d_out, 0, sizeof(uint32_t), h->stream));
      const uint64_t before = global_cnt;
      if ((rc = take_stats(d_out))) return rc;
      h->rounds += 1;
      stalls = global_cnt == before ? stalls + 1 : 0;
      if (stalls > 64) ABM_FAIL(h, ABM_E_INTERNAL, "conflict resolution made no progress with %llu agents pending", (unsigned long long)global_cnt);
      in ^= 1;
    }
  }
  if ((rc = commit_phase(birth_flags, nullptr))) return rc;
  if (h->strip && (rc = strip_birth_ids(h, birth_cap, nxt, ctr + CTR_BIRTH_TOTAL))) return rc;
  if ((rc = read_counters(h))) return rc;
  if ((rc = device_error_to_code(h))) return rc;
  const uint64_t n_new = h->h_counters[CTR_CURSOR], nb = h->h_counters[CTR_BIRTHS];
  uint64_t nb_total = nb;
  if (h->strip) {
    nb_total = h->h_counters[CTR_BIRTH_TOTAL];
    if (nb > birth_cap) ABM_FAIL(h, ABM_E_INTERNAL, "more offspring (%llu) than the agreed message capacity", (unsigned long long)nb);
    h->bcap = h->next_bcap;
    strip_fmax = (birth_cap - 64) / 3;
  } else if (nb > 0) {
    if ((rc = rank_prefix(h, words))) return rc;
    BirthIdBody bi;
    bi.birth_pos = (const uint32_t*)h->birth_pos.p; bi.birth_parent = (const uint32_t*)h->birth_parent.p;
    bi.birth_bits = (const uint32_t*)h->bits.p; bi.birth_prefix = (const uint32_t*)h->bits_prefix.p;
    bi.next_id = (uint32_t)h->next_id; bi.out_id = (uint32_t*)h->id[nxt].p; bi.birth_count = nullptr;
    { ProfScope ps(h->prof, h->stream, "birth_ids", nb); launch_foreach(h->stream, bi, nb); }
    if ((rc = check_launch(h, "birth ids"))) return rc;
  }
  if ((rc = finish_tick(n_new, nb_total))) return rc;
  if (h->strip) h->bf_est = 2 * strip_fmax + 1024; /* the same on every rank */
  return 0;
}

static int validate_config(const abm_config* c, std::string& why) {
  char b[256];
#define BAD(...) do { snprintf(b, sizeof b, __VA_ARGS__); why = b; return ABM_E_BADARG; } while (0)
  if (c->struct_size != (int32_t)sizeof(abm_config)) BAD("abm_config.struct_size %d != %zu", c->struct_size, sizeof(abm_config));
  if (c->nx < 1 || c->ny < 1) BAD("grid dims must be positive");
  const uint64_t tx = ((uint64_t)c->nx + 31) / 32, ty = ((uint64_t)c->ny + 31) / 32;
  if (tx * ty * 1024 >= 0xFFFFFFF0ull) BAD("grid too large for 32-bit cell indices");
  if (c->walk_type < 0 || c->walk_type > ABM_RANDOM_WALKMAP) BAD("walk_type %d out of range", c->walk_type);
  if (c->regrowth_time < 0 || c->regrowth_time > 126) BAD("regrowth_time must be in 0..126");
  if (c->walk_type == ABM_RANDOM_WALK_GREGARIOUS) {
    if (!c->periodic) BAD("RANDOM_WALK_GREGARIOUS needs periodic=true (the reference indexes 8 neighbours, src/agent_actions.jl:334)");
    if (c->visual_distance < 1 || c->visual_distance > GREG_MAX_R) BAD("visual_distance must be in 1..%d", (int)GREG_MAX_R);
    if (c->nx < 2 * c->visual_distance + 3 || c->ny < 2 * c->visual_distance + 3) BAD("gregarious walk needs dims >= 2*visual_distance+3");
  }
  if (c->walk_type == ABM_RANDOM_WALK_ATTRACTOR) {
    const double ax = 2 * c->attractor_x, ay = 2 * c->attractor_y;
    if (ax != floor(ax) || ay != floor(ay) || fabs(ax) > 1e9 || fabs(ay) > 1e9) BAD("2*attractor must be integral (griddims ./ 2)");
  }
  if (c->walk_type == ABM_ALTERNATED_WALK && c->counter < 1) BAD("counter must be >= 1");
  /* on a periodic grid with a dimension below 3 two of the 8 neighbour offsets wrap onto the same cell: the batched A* relaxes
   * the 8 neighbours side by side and does not reproduce the reference's one-after-the-other result there */
  if (c->walk_type == ABM_ALTERNATED_WALK && c->periodic && (c->nx < 3 || c->ny < 3)) BAD("ALTERNATED_WALK (A* routes) on a periodic grid needs dims >= 3");
  if (!(c->prob_random_walk >= 0 && c->prob_random_walk <= 1)) BAD("prob_random_walk must be in [0,1]");
  if (c->astar_metric != ABM_ASTAR_DIRECT && c->astar_metric != ABM_ASTAR_MAX) BAD("astar_metric out of range");
  if (c->scheduler != ABM_SCHED_BY_ID && c->scheduler != ABM_SCHED_RANDOMLY) BAD("scheduler out of range");
  if (c->n_ranks > 1 || c->strip_ny > 0) { /* row strip */
    const int P = c->n_ranks > 1 ? c->n_ranks : 1;
    if (c->rank < 0 || c->rank >= P) BAD("rank out of range");
    if (c->strip_ny < 32 || c->strip_ny % 32 || c->strip_y0 % 32 || c->strip_y0 < 0 || c->strip_y0 + c->strip_ny > c->ny) BAD("strip rows must be multiples of 32 inside the grid");
    if (c->nx % 32 || c->nx < 64) BAD("strip mode needs nx to be a multiple of 32, at least 64");
    if (c->walk_type == ABM_ALTERNATED_WALK) BAD("ALTERNATED_WALK (model-global counter, route teleports) is not available in strip mode");
    if (P == 1 && c->periodic && c->strip_ny != c->ny) BAD("a single periodic strip must cover the whole grid");
  }
#undef BAD
  return 0;
}

}  // namespace

/* ================================================================================================ C ABI */
#include "abm_capi.inl"
#include "abm_group.inl"
