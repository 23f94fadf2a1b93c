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
    if (!recs.empty() && last_ended && launches_at_end == g_launches) { r.a = recs.back().b; r.own_a = false; }
    else { r.a = get(); r.own_a = true; cudaEventRecord(r.a, s); }
    last_ended = false;
    recs.push_back(r);
#else
    (void)s;
#endif
  }
  void end(abm_stream_t s) {
#ifndef ABM_EMU
    if (on && !recs.empty()) { cudaEventRecord(recs.back().b, s); launches_at_end = g_launches; last_ended = true; }
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
  DevBuf grass, walkbits, elev;
  DevBuf bits, bits_prefix, scan_tmp, stage, lut, greg_order, greg_dxy, greg_table, counters, partials;
  /* ALTERNATED_WALK / A*: id-order ranks, behaviour segments, plan requests, route followers, the route store */
  DevBuf rank, by_rank, seg_behav, plan_agent, plan_goal, remote_head, remote_next, remote_dest, idx_of_id;
  DevBuf as_g, as_meta, as_heap, as_src, as_dst, as_id, as_status, as_len, as_cost, as_off, as_slot_gen, as_todo, as_comp;
  bool comp_valid = false;        /* as_comp holds the walkmap's component labels */
  uint64_t astar_budget = 64ull << 30; /* ABM_ASTAR_GB: device bytes the search slots may take (and at most half of what is free) */
  DevBuf rt_slot_of_id, rt_off, rt_len, rt_pos, rt_cost, rt_pool;
  uint64_t as_slots = 0, as_heap_cap = 0, rt_slots_cap = 0, rt_pool_used = 0, rt_ids_cap = 0;
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

static View make_view(abm_handle* h) {
  View v;
  memset(&v, 0, sizeof v);
  v.g = h->g;
  v.p = h->params;
  v.p.tick = h->tick;
  v.n = (uint32_t)h->n;
  v.cell = (uint32_t*)h->cell[h->cur].p;
  v.id = (uint32_t*)h->id[h->cur].p;
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
  if (e & ERR_P2P_TIMEOUT) ABM_FAIL(h, ABM_E_CUDA, "peer exchange timed out: a neighbouring strip did not deliver its message");
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
  if (B > 0xFFFFFFF0ull) ABM_FAIL(h, ABM_E_BADARG, "too many route requests in one batch");
  const uint64_t gt = h->g.gt;
  const uint64_t heap_cap = std::min<uint64_t>(8ull * (uint64_t)h->g.nx * h->g.ny + 16, 1ull << 20);
  const uint64_t per_slot = gt * 6 + heap_cap * 8;
  uint64_t max_resident = 148ull * 32; /* one warp-sized CTA per slot, at most 32 CTAs per SM */
  if (const char* ms = getenv("ABM_ASTAR_SLOTS")) max_resident = (uint64_t)std::max(atoll(ms), 1ll);
  if (h->as_slots == 0 || h->as_heap_cap != heap_cap || h->as_slots < std::min<uint64_t>(B, max_resident)) {
    const uint64_t budget = std::min<uint64_t>(device_free_bytes() / 2, h->astar_budget);
    uint64_t slots = std::max<uint64_t>(1, std::min<uint64_t>(std::min<uint64_t>(B, max_resident), budget / per_slot));
    if (slots > h->as_slots || h->as_heap_cap != heap_cap) {
      if ((rc = ensure(h, h->as_g, slots * gt * 4))) return rc;
      if ((rc = ensure(h, h->as_meta, slots * gt * 2))) return rc;
      if ((rc = ensure(h, h->as_heap, slots * heap_cap * 8))) return rc;
      if ((rc = ensure(h, h->as_slot_gen, slots * 4))) return rc;
      ABM_CK(h, dev_memset(h->as_meta.p, 0, slots * gt * 2, h->stream));
      ABM_CK(h, dev_memset(h->as_slot_gen.p, 0, slots * 4, h->stream));
      h->as_slots = slots; h->as_heap_cap = heap_cap;
    }
  }
  const uint64_t S = std::min<uint64_t>(std::min<uint64_t>(h->as_slots, B), max_resident);
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
    sb.heap_cap = (uint32_t)heap_cap; sb.slot_gen = (uint32_t*)h->as_slot_gen.p;
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
    if (todo.empty()) { for (uint64_t i = 0; i < B; ++i) if (st[i] == AS_POOL_FULL) again.push_back((uint32_t)i); }
    else for (uint32_t i : todo) if (st[i] == AS_POOL_FULL) again.push_back(i);
    for (uint64_t i = 0; i < B; ++i)
      if (st[i] == AS_HEAP_FULL) ABM_FAIL(h, ABM_E_CAPACITY, "A* open list exceeded %llu entries", (unsigned long long)heap_cap);
    if (again.empty()) { h->rt_pool_used = cursor; break; }
    /* the cursor ran past the end: everything that fitted stays (and stays counted), the rest runs again in a larger pool */
    h->rt_pool_used = std::min<uint64_t>(cursor, pool_cap);
    if ((rc = grow_keep(h, h->rt_pool, (2 * pool_cap + (cursor - std::min<uint64_t>(cursor, pool_cap)) + 1) * 4))) return rc;
    pool_cap = h->rt_pool.bytes / 4;
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
  const uint64_t words = bitmap_words(h->next_id);
  if ((rc = ensure(h, h->bits, (words + 4) * 4))) return rc;
  if ((rc = ensure(h, h->rank, n * 4))) return rc;
  if ((rc = ensure(h, h->by_rank, n * 4))) return rc;
  if ((rc = ensure(h, h->remote_next, n * 4))) return rc;
  if ((rc = ensure(h, h->remote_dest, n * 4))) return rc;
  if ((rc = ensure_route_ids(h, h->next_id))) return rc;
  v = make_view(h);
  ABM_CK(h, dev_memset(h->bits.p, 0, words * 4, h->stream));
  SetBitsBody sbb;
  sbb.id = v.id; sbb.bits = (uint32_t*)h->bits.p;
  { ProfScope ps(h->prof, h->stream, "rank_by_id", n, 4); launch_foreach(h->stream, sbb, n);
    if ((rc = rank_prefix(h, words))) return rc;
    RankBody rb;
    rb.id = v.id; rb.alive_bits = (const uint32_t*)h->bits.p; rb.prefix = (const uint32_t*)h->bits_prefix.p;
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

/* one tick: every agent's custom_agent_step! in id order, then custom_model_step! */
static int tick_once(abm_handle* h) {
  int rc;
  const uint64_t n = h->n;
  const uint64_t gt = h->g.gt;
  if ((rc = ensure_agent_scratch(h, n))) return rc;
  uint32_t* ctr = (uint32_t*)h->counters.p;
  ABM_CK(h, dev_memset(ctr, 0, 3 * sizeof(uint32_t), h->stream)); /* keeps the sticky error word */
  View v = make_view(h);

  if (n > 0 && h->cfg.walk_type == ABM_ALTERNATED_WALK) {
    if ((rc = alternated_propose(h, v, ctr))) return rc;
  } else if (n > 0) {
    ProposeBody pb;
    pb.v = v;
    pb.pending = (uint32_t*)h->pend[0].p;
    pb.pending_count = ctr + CTR_PEND_A;
    pb.birth_flags = nullptr;
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
  const uint64_t words = bitmap_words(h->next_id);
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
    const uint64_t idw = bitmap_words(h->next_id);
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
  View v = make_view(h);
  const SubIndex si = sub_index(h, h->cur);
  const int wt = h->cfg.walk_type;
  const bool needs_resolve = wt == ABM_RANDOM_WALK_ATTRACTOR || wt == ABM_RANDOM_WALK_GREGARIOUS || (wt == ABM_RANDOM_WALK && h->cfg.randomwalk_ifempty);
  uint64_t cnt = 0;
  if (n > 0) {
    ProposeBody pb;
    pb.v = v; pb.pending = nullptr; pb.pending_count = nullptr; pb.birth_flags = ctr + CTR_BIRTH_FLAGS;
    { ProfScope ps(h->prof, h->stream, "propose", n); launch_foreach(h->stream, pb, n); }
    if ((rc = check_launch(h, "propose"))) return rc;
  }
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
  if (h->strip && (rc = strip_exchange_agents(h))) return rc; /* neighbours' boundary rows become ghost rows behind the own agents */
  if (n + h->n_ghost > 0 && needs_resolve && (rc = launch_tiles(0, core_rows))) return rc;
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
  const uint64_t words = bitmap_words(h->next_id);
  /* commit + grass + re-grouping (+ offspring ids on one GPU); `go` = device flag checked by the kernels, or null */
  auto commit_phase = [&](uint64_t bflags, const uint32_t* go) -> int {
    const uint64_t arrivals_cap = h->strip ? 2 * h->bcap : 0; /* ghost agents (and their offspring) may land here */
    int r;
    if ((r = ensure_soa(h, nxt, n + bflags + 2 * arrivals_cap))) return r;
    if ((r = ensure(h, h->birth_pos, (bflags + arrivals_cap + 1) * 4))) return r;
    if ((r = ensure(h, h->birth_parent, (bflags + arrivals_cap + 1) * 4))) return r;
    if (bflags > 0 || h->strip) {
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
    ct.birth_bits = h->strip ? nullptr : (uint32_t*)h->bits.p; /* strips rank the parents globally instead */
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
  if (h->optimistic && h->bf_est > 0 && (!h->strip || strip_stats_on_device(h))) {
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
    if (h->strip) {
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
  if (!(c->prob_random_walk >= 0 && c->prob_random_walk <= 1)) BAD("prob_random_walk must be in [0,1]");
  if (c->astar_metric != ABM_ASTAR_DIRECT && c->astar_metric != ABM_ASTAR_MAX) BAD("astar_metric out of range");
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

const char* ABM_API(last_error)(abm_handle* h) { return h ? h->err.c_str() : g_create_error.c_str(); }

void ABM_API(destroy)(abm_handle* h) {
  if (!h) return;
#ifndef ABM_EMU
  cudaSetDevice(h->device);
  if (h->stream) { cudaStreamSynchronize(h->stream); }
  for (int r = 0; r < P2P_MAXP; ++r) if (h->p2p_peer[r] && h->p2p_peer[r] != h->p2p_mine) cudaIpcCloseMemHandle(h->p2p_peer[r]);
  if (h->p2p_mine) cudaFree(h->p2p_mine);
  if (h->nccl_comm && g_nccl.CommDestroy) g_nccl.CommDestroy((ncclComm_t)h->nccl_comm);
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
    if (cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaMallocHost((void**)&h->h_counters, CTR_WORDS * sizeof(uint32_t)) != cudaSuccess) {
      g_create_error = "stream / pinned memory creation failed";
      delete h; return ABM_E_CUDA;
    }
  }
#else
  h->h_counters = (uint32_t*)calloc(CTR_WORDS, sizeof(uint32_t));
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
  if (const char* mg = getenv("ABM_MIGRANTS")) h->use_migrants = atoi(mg) != 0;
  if (const char* ag = getenv("ABM_ASTAR_GB")) { const int q = atoi(ag); if (q >= 1 && q <= 160) h->astar_budget = (uint64_t)q << 30; }
  if (const char* bb = getenv("ABM_BIRTH_BITMAP")) h->birth_bitmap_from = (uint64_t)std::max(atoll(bb), 0ll);
  if (const char* op = getenv("ABM_OPTIMISTIC")) h->optimistic = atoi(op) != 0;
  if (const char* ar = getenv("ABM_AGENT_ROUNDS")) { const int q = atoi(ar); if (q >= 0 && q <= 16) h->fixed_agent_rounds = h->min_agent_rounds = q; }
  if (const char* tm = getenv("ABM_TILE_MARGIN")) { const int q = atoi(tm); if (q >= 2 && q <= HALO) h->tile_margin = q; }
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
  if (!fg || !cd) ABM_FAIL(h, ABM_E_BADARG, "fully_grown and countdown are required");
  int rc = set_device(h);
  if (rc) return rc;
  const Geo& g = h->g;
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
    ABM_CK(h, copy_h2d(s_cd, cd + off, cells * 8, h->stream));
    ABM_CK(h, copy_h2d(s_fg, fg + off, cells, h->stream));
    if (wm) ABM_CK(h, copy_h2d(s_wm, wm + off, cells, h->stream));
    if (el) ABM_CK(h, copy_h2d(s_el, el + off, cells * 8, h->stream));
    GridImportBody ib;
    ib.g = g; ib.fg = s_fg; ib.cd = s_cd; ib.wm = wm ? s_wm : nullptr; ib.el = el ? s_el : nullptr;
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
static int ingest_staged(abm_handle* h, int64_t n, int64_t next_id) {
  int rc;
  const Geo& g = h->g;
  int64_t* s_id = (int64_t*)h->stage.p;
  int64_t* s_x = s_id + n;
  int64_t* s_y = s_x + n;
  double* s_e = (double*)(s_y + n);
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
      ib.g = g; ib.nsx = h->nsx; ib.hid = s_id; ib.hx = s_x; ib.hy = s_y; ib.he = s_e; ib.next_id = next_id;
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
    ib.g = g; ib.hid = s_id; ib.hx = s_x; ib.hy = s_y; ib.he = s_e; ib.next_id = next_id;
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
    ABM_CK(h, copy_h2d(s_id, id, (size_t)n * 8, h->stream));
    ABM_CK(h, copy_h2d(s_id + n, x, (size_t)n * 8, h->stream));
    ABM_CK(h, copy_h2d(s_id + 2 * n, y, (size_t)n * 8, h->stream));
    ABM_CK(h, copy_h2d(s_id + 3 * n, energy, (size_t)n * 8, h->stream));
  }
  return ingest_staged(h, n, next_id);
}

/* the random part of initialize_model, built on the device (include/abm.h) */
int ABM_API(init_random)(abm_handle* h, int64_t n_sheep, uint64_t init_seed) {
  if (!h) return ABM_E_BADARG;
  if (h->strip) ABM_FAIL(h, ABM_E_BADARG, "abm_init_random is not available on a strip handle (placement needs the whole walkmap)");
  const int64_t span = (int64_t)(h->cfg.delta_energy * 5);
  if (n_sheep < 0 || n_sheep >= 0xFFFFFFF0ll) ABM_FAIL(h, ABM_E_CAPACITY, "n_sheep must be below 2^32-16");
  if (span < 1) ABM_FAIL(h, ABM_E_BADARG, "delta_energy * 5 must be at least 1 (energy = rand(1:delta_energy*5) - 1)");
  int rc = set_device(h);
  if (rc) return rc;
  const Geo& g = h->g;
  uint32_t* err = (uint32_t*)h->counters.p + CTR_ERR;
  ABM_CK(h, dev_memset(err, 0, 4, h->stream));
  if (!h->grid_set) { /* no terrain was given: every cell of the space is walkable, elevation 0 */
    ABM_CK(h, dev_memset(h->grass.p, GRASS_PAD, g.gt, h->stream));
    ABM_CK(h, dev_memset(h->walkbits.p, 0, (size_t)g.gt / 8, h->stream));
    if ((rc = ensure(h, h->elev, (size_t)g.gt * 2))) return rc;
    ABM_CK(h, dev_memset(h->elev.p, 0, (size_t)g.gt * 2, h->stream));
    h->comp_valid = false;
  }
  InitGrassBody gb;
  gb.g = g; gb.seed = init_seed; gb.regrowth_time = h->cfg.regrowth_time; gb.open_terrain = h->grid_set ? 0 : 1;
  gb.grass = (uint8_t*)h->grass.p; gb.walkbits = (uint32_t*)h->walkbits.p;
  launch_foreach(h->stream, gb, (uint64_t)g.nx * (uint64_t)g.ny);
  if ((rc = check_launch(h, "grass initialisation"))) return rc;
  h->grid_set = true;
  if ((rc = stage_agents(h, n_sheep))) return rc;
  if (n_sheep > 0) {
    InitSheepBody sb;
    sb.g = g; sb.seed = init_seed; sb.span = (uint64_t)span; sb.walkbits = (const uint32_t*)h->walkbits.p;
    sb.s_id = (int64_t*)h->stage.p; sb.s_x = sb.s_id + n_sheep; sb.s_y = sb.s_x + n_sheep; sb.s_e = (double*)(sb.s_y + n_sheep);
    sb.err = err;
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

int ABM_API(get_agents)(abm_handle* h, int64_t* n_inout, int64_t* id, int64_t* x, int64_t* y, double* energy) {
  if (!h || !n_inout) return ABM_E_BADARG;
  const int64_t n = (int64_t)h->n;
  if (*n_inout < n) { *n_inout = n; ABM_FAIL(h, ABM_E_SMALL, "agent buffers hold %lld entries, %lld needed", (long long)*n_inout, (long long)n); }
  *n_inout = n;
  if (n == 0) return ABM_OK;
  int rc = set_device(h);
  if (rc) return rc;
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
  ExportBody eb;
  eb.g = h->g;
  eb.cell = (const uint32_t*)h->cell[h->cur].p; eb.id = (const uint32_t*)h->id[h->cur].p; eb.energy = (const double*)h->energy[h->cur].p;
  eb.alive_bits = (const uint32_t*)h->bits.p; eb.prefix = (const uint32_t*)h->bits_prefix.p;
  eb.oid = s_id; eb.ox = s_x; eb.oy = s_y; eb.oe = s_e;
  launch_foreach(h->stream, eb, (uint64_t)n);
  if ((rc = check_launch(h, "export"))) return rc;
  if (id) ABM_CK(h, copy_d2h(id, s_id, (size_t)n * 8, h->stream));
  if (x) ABM_CK(h, copy_d2h(x, s_x, (size_t)n * 8, h->stream));
  if (y) ABM_CK(h, copy_d2h(y, s_y, (size_t)n * 8, h->stream));
  if (energy) ABM_CK(h, copy_d2h(energy, s_e, (size_t)n * 8, h->stream));
  return ABM_OK;
}

int ABM_API(set_global_state)(abm_handle* h, int64_t tick, int64_t behav, int64_t behav_counter) {
  if (!h) return ABM_E_BADARG;
  if (tick < 0 || tick > 0xFFFFFFFFll) ABM_FAIL(h, ABM_E_BADARG, "tick must fit 32 bits");
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
  int rc = set_device(h);
  if (rc) return rc;
  const uint64_t l0 = launches_now();
  h->rounds = 0;
  h->updates = 0;
  h->prof.reset();
  h->tm_total.start(h->stream);
  for (int64_t t = 0; t < n_ticks; ++t) {
    if (h->tick >= 0xFFFFFFFFull) ABM_FAIL(h, ABM_E_CAPACITY, "tick counter exhausted");
    if ((rc = h->tile_mode ? tick_tile(h) : tick_once(h))) return rc;
  }
  h->tm_total.stop(h->stream);
  ABM_CK(h, stream_sync(h->stream));
  h->t_total = h->tm_total.ms();
  h->prof.collect();
  h->launches = (double)(launches_now() - l0);
  return ABM_OK;
}

int ABM_API(reduce)(abm_handle* h, int32_t what, double* out) {
  if (!h || !out) return ABM_E_BADARG;
  int rc = set_device(h);
  if (rc) return rc;
  if (what == ABM_REDUCE_COUNT_AGENTS) { *out = (double)h->n; return ABM_OK; }
  if (what == ABM_REDUCE_COUNT_GROWN) {
    if (!h->grid_set) ABM_FAIL(h, ABM_E_STATE, "abm_set_grid has not been called");
    unsigned long long* acc = (unsigned long long*)((uint32_t*)h->counters.p + CTR_RED_LO);
    ABM_CK(h, dev_memset(acc, 0, 8, h->stream));
    CountGrownBody cb;
    cb.grass4 = (const uint32_t*)h->grass.p; cb.out = acc;
    launch_foreach(h->stream, cb, (uint64_t)h->g.gt / 4);
    if ((rc = check_launch(h, "count grown"))) return rc;
    unsigned long long v = 0;
    ABM_CK(h, copy_d2h(&v, acc, 8, h->stream));
    *out = (double)v;
    return ABM_OK;
  }
  if (what == ABM_REDUCE_SUM_ENERGY || what == ABM_REDUCE_MAX_ENERGY || what == ABM_REDUCE_MIN_ENERGY) {
    const uint64_t n = h->n;
    if (n == 0) { *out = 0; return ABM_OK; }
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
    h->p2p_capmax = std::max<uint64_t>(1u << 16, h->n / 4 + 4096);   /* agents per boundary row the mailbox can carry */
    h->p2p_birthmax = std::max<uint64_t>(1u << 16, h->n / 8 + 4096);
    h->p2p_lay.agents_bytes = (strip_msg_bytes(h->nsx, h->p2p_capmax) + 255) & ~(uint64_t)255;
    h->p2p_lay.state_bytes = (h->p2p_capmax + 255) & ~(uint64_t)255;
    h->p2p_lay.births_bytes = ((1 + h->p2p_birthmax) * 4 + 255) & ~(uint64_t)255;
    ABM_CK(h, cudaMalloc((void**)&h->p2p_mine, h->p2p_lay.total_bytes()));
    ABM_CK(h, cudaMemset(h->p2p_mine, 0, h->p2p_lay.total_bytes()));
    ABM_CK(h, cudaDeviceSynchronize());
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

int ABM_API(p2p_connect)(abm_handle* h, const void* handles, int32_t n) {
  if (!h || !handles) return ABM_E_BADARG;
#ifndef ABM_EMU
  if (!h->p2p_mine || n != h->n_ranks) ABM_FAIL(h, ABM_E_STATE, "abm_p2p_connect needs abm_p2p_export first and one handle per rank");
  int rc = set_device(h);
  if (rc) return rc;
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
