/*
 * abm_group.inl — single-process multi-GPU: one call steps N row strips on N devices.
 *
 * The reference's caller is ONE Julia process (scripts/v0.3.jl:124: abmvideo -> step!(model, agent_step!, model_step!, n));
 * it cannot start a process per GPU.  An abm_group owns N strip handles (one per device) of one GridSpace and drives them
 * from the calling thread: every collective call (step, reduce, init) runs the members on N worker threads, and the
 * strips exchange their boundaries over NVLink peer memory mapped directly inside the process (cudaDeviceEnablePeerAccess;
 * no CUDA IPC, no NCCL, no MPI) — or, when the devices cannot reach each other (and in the emulated build), through an
 * in-process mailbox in host memory behind the same abm_transport callbacks a multi-process host would supply.
 * Whole-GridSpace arrays go in and come out: the group cuts them into strips (with their ghost rows) and merges them back.
 * Part of abm_engine.cu (one translation unit); not a stand-alone header.
 */
#include <condition_variable>
#include <functional>
#include <mutex>
#include <thread>

namespace {

/* in-process ring transport: the members' worker threads meet at a barrier and copy from each other's host buffers */
struct GroupBus {
  int n = 0;
  std::mutex m;
  std::condition_variable cv;
  int arrived = 0;
  uint64_t gen = 0;
  std::vector<const void*> a, b;
  std::vector<uint64_t> la, lb;
  void barrier() {
    std::unique_lock<std::mutex> lk(m);
    const uint64_t g = gen;
    if (++arrived == n) { arrived = 0; ++gen; cv.notify_all(); }
    else cv.wait(lk, [&] { return gen != g; });
  }
};
struct GroupPort { GroupBus* bus; int rank; };

static int bus_sendrecv(void* user, const void* to_prev, uint64_t b_tp, const void* to_next, uint64_t b_tn, void* from_prev, uint64_t b_fp, void* from_next, uint64_t b_fn) {
  GroupPort* p = (GroupPort*)user;
  GroupBus& B = *p->bus;
  const int r = p->rank, prev = (r + B.n - 1) % B.n, next = (r + 1) % B.n;
  B.a[r] = to_prev; B.la[r] = b_tp; B.b[r] = to_next; B.lb[r] = b_tn;
  B.barrier();
  int rc = 0;
  if (b_fp) { if (B.lb[prev] != b_fp) rc = 1; else memcpy(from_prev, B.b[prev], b_fp); }
  if (b_fn) { if (B.la[next] != b_fn) rc = 1; else memcpy(from_next, B.a[next], b_fn); }
  B.barrier();
  return rc;
}
static int bus_allreduce(void* user, uint64_t* values, int32_t n) {
  GroupPort* p = (GroupPort*)user;
  GroupBus& B = *p->bus;
  B.a[p->rank] = values;
  B.barrier();
  std::vector<uint64_t> sum((size_t)n, 0);
  for (int r = 0; r < B.n; ++r) for (int i = 0; i < n; ++i) sum[i] += ((const uint64_t*)B.a[r])[i];
  B.barrier();
  memcpy(values, sum.data(), (size_t)n * 8);
  B.barrier();
  return 0;
}
static int bus_allgather(void* user, const uint32_t* send, uint32_t* recv, uint32_t n) {
  GroupPort* p = (GroupPort*)user;
  GroupBus& B = *p->bus;
  B.a[p->rank] = send;
  B.barrier();
  for (int r = 0; r < B.n; ++r) memcpy(recv + (size_t)r * n, B.a[r], (size_t)n * 4);
  B.barrier();
  return 0;
}

}  // namespace

struct abm_group {
  abm_config cfg;
  int n = 0;
  std::vector<abm_handle*> h;
  std::vector<int> y0, rows;
  GroupBus bus;
  std::vector<GroupPort> port;
  bool peer_memory = false;
  std::string err;
};

namespace {

#define GRP_FAIL(g, code, ...) do { char b_[512]; snprintf(b_, sizeof b_, __VA_ARGS__); (g)->err = b_; return (code); } while (0)

/* fn(member index) on one worker thread per member; the first failure (and its text) is reported */
static int group_parallel(abm_group* g, const std::function<int(int)>& fn) {
  std::vector<int> rc((size_t)g->n, 0);
  std::vector<std::thread> th;
  for (int i = 1; i < g->n; ++i) th.emplace_back([&, i] { rc[i] = fn(i); });
  rc[0] = fn(0);
  for (size_t i = 0; i < th.size(); ++i) th[i].join();
  for (int i = 0; i < g->n; ++i) if (rc[i]) { g->err = "device " + std::to_string(g->h[i]->device) + ": " + g->h[i]->err; return rc[i]; }
  return 0;
}

/* rows y0-32 .. y0+n+32 of a whole-grid array (wrapped when periodic, `fill` outside otherwise) */
template <class T>
static void strip_rows(const T* a, int nx, int ny, int y0, int n, bool periodic, T fill, std::vector<T>& out) {
  const int G = CORE;
  out.assign((size_t)(n + 2 * G) * nx, fill);
  for (int r = 0; r < n + 2 * G; ++r) {
    int y = y0 - G + r;
    if (periodic) y = ((y % ny) + ny) % ny; else if (y < 0 || y >= ny) continue;
    memcpy(&out[(size_t)r * nx], a + (size_t)y * nx, (size_t)nx * sizeof(T));
  }
}

}  // namespace

extern "C" {

const char* ABM_API(group_last_error)(abm_group* g) { return g ? g->err.c_str() : g_create_error.c_str(); }

void ABM_API(group_destroy)(abm_group* g) {
  if (!g) return;
  for (size_t i = 0; i < g->h.size(); ++i) {
#ifndef ABM_EMU
    if (g->h[i] && g->peer_memory) for (int r = 0; r < P2P_MAXP; ++r) g->h[i]->p2p_peer[r] = nullptr; /* not IPC mappings: nothing to close */
#endif
    ABM_API(destroy)(g->h[i]);
  }
  delete g;
}

int ABM_API(group_create)(const abm_config* cfg, int32_t n_devices, const int32_t* device_ids, abm_group** out) {
  if (!cfg || !out || n_devices < 1 || n_devices > 8) { g_create_error = "abm_group_create: 1..8 devices"; return ABM_E_BADARG; }
  *out = nullptr;
  if (cfg->nx % CORE || cfg->ny % CORE || cfg->ny / CORE < n_devices) { g_create_error = "abm_group_create: grid dims must be multiples of 32 with at least one 32-row core row per device"; return ABM_E_BADARG; }
#ifndef ABM_EMU
  /* one strip per device: while a strip's kernel waits for its neighbour's message, an allocation by the neighbour's thread
   * would synchronise the whole device (cudaMalloc / cudaFree do) and never return */
  for (int i = 0; i < n_devices; ++i) for (int j = 0; j < i; ++j)
    if ((device_ids ? device_ids[i] : i) == (device_ids ? device_ids[j] : j)) { g_create_error = "abm_group_create: every strip needs its own device"; return ABM_E_BADARG; }
#endif
  abm_group* g = new (std::nothrow) abm_group();
  if (!g) return ABM_E_NOMEM;
  g->cfg = *cfg; g->n = n_devices;
  g->bus.n = n_devices; g->bus.a.assign(n_devices, nullptr); g->bus.b.assign(n_devices, nullptr); g->bus.la.assign(n_devices, 0); g->bus.lb.assign(n_devices, 0);
  g->port.resize(n_devices);
  const int cores = cfg->ny / CORE;
  for (int i = 0; i < n_devices; ++i) {
    const int lo = (int)((int64_t)cores * i / n_devices), hi = (int)((int64_t)cores * (i + 1) / n_devices);
    abm_config c = *cfg;
    c.strip_y0 = lo * CORE; c.strip_ny = (hi - lo) * CORE; c.n_ranks = n_devices; c.rank = i;
    c.device = device_ids ? device_ids[i] : i;
    abm_handle* h = nullptr;
    const int rc = ABM_API(create)(&c, &h);
    if (rc) { ABM_API(group_destroy)(g); return rc; }
    g->h.push_back(h); g->y0.push_back(c.strip_y0); g->rows.push_back(c.strip_ny);
    g->port[i].bus = &g->bus; g->port[i].rank = i;
    abm_transport t;
    t.user = &g->port[i]; t.sendrecv = bus_sendrecv; t.allreduce_sum_u64 = bus_allreduce; t.allgather_u32 = bus_allgather;
    ABM_API(set_transport)(h, &t); /* the fallback; peer memory replaces the data path below when the devices allow it */
  }
  *out = g;
  return ABM_OK;
}

abm_handle* ABM_API(group_member)(abm_group* g, int32_t i) { return (g && i >= 0 && i < g->n) ? g->h[i] : nullptr; }

/* after the population is known (the mailboxes are allocated from the grid): map every member's mailbox into the others */
static int group_connect(abm_group* g) {
#ifndef ABM_EMU
  if (g->peer_memory || g->n < 2 || getenv("ABM_GROUP_HOST_TRANSPORT")) return 0;
  for (int i = 0; i < g->n; ++i) for (int j = 0; j < g->n; ++j) {
    if (i == j || g->h[i]->device == g->h[j]->device) continue;
    int can = 0;
    if (cudaDeviceCanAccessPeer(&can, g->h[i]->device, g->h[j]->device) != cudaSuccess || !can) return 0; /* stay on the host transport */
  }
  uint8_t dummy[64];
  for (int i = 0; i < g->n; ++i) {
    int rc = ABM_API(p2p_export)(g->h[i], dummy); /* allocates the mailbox (the IPC handle is not needed inside one process) */
    if (rc) { g->err = g->h[i]->err; return rc; }
  }
  for (int i = 0; i < g->n; ++i) {
    cudaSetDevice(g->h[i]->device);
    if (int rc = p2p_apply_timeout(g->h[i])) { g->err = g->h[i]->err; return rc; }
    for (int j = 0; j < g->n; ++j) {
      if (i != j && g->h[i]->device != g->h[j]->device) {
        cudaError_t e = cudaDeviceEnablePeerAccess(g->h[j]->device, 0);
        if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) { cudaGetLastError(); return 0; }
        cudaGetLastError();
      }
      g->h[i]->p2p_peer[j] = g->h[j]->p2p_mine;
    }
  }
  for (int i = 0; i < g->n; ++i) g->h[i]->p2p_on = true;
  g->peer_memory = true;
#else
  (void)g;
#endif
  return 0;
}

int ABM_API(group_set_grid)(abm_group* g, const uint8_t* fg, const int64_t* cd, const uint8_t* wm, const int64_t* el) {
  if (!g) return ABM_E_BADARG;
  const int nx = g->cfg.nx, ny = g->cfg.ny;
  const bool per = g->cfg.periodic != 0;
  std::vector<uint32_t> bits;
  if (wm) { /* random_walkable of abm_init_random draws over the whole GridSpace: every member gets the whole walkmap as bits */
    bits.assign(((size_t)nx * ny + 31) / 32, 0u);
    for (size_t l = 0; l < (size_t)nx * ny; ++l) if (wm[l]) bits[l >> 5] |= 1u << (l & 31);
  }
  return group_parallel(g, [&](int i) -> int {
    std::vector<uint8_t> sfg, swm; std::vector<int64_t> scd, sel;
    if (fg) strip_rows<uint8_t>(fg, nx, ny, g->y0[i], g->rows[i], per, 0, sfg);
    if (cd) strip_rows<int64_t>(cd, nx, ny, g->y0[i], g->rows[i], per, 0, scd);
    if (wm) strip_rows<uint8_t>(wm, nx, ny, g->y0[i], g->rows[i], per, 0, swm);
    if (el) strip_rows<int64_t>(el, nx, ny, g->y0[i], g->rows[i], per, 0, sel);
    int rc = ABM_API(set_grid)(g->h[i], fg ? sfg.data() : nullptr, cd ? scd.data() : nullptr, wm ? swm.data() : nullptr, el ? sel.data() : nullptr);
    if (!rc) rc = ABM_API(set_global_walkmap_bits)(g->h[i], wm ? bits.data() : nullptr);
    return rc;
  });
}

int ABM_API(group_set_agents)(abm_group* g, int64_t n, const int64_t* id, const int64_t* x, const int64_t* y, const double* energy, int64_t next_id) {
  if (!g || n < 0 || (n > 0 && (!id || !x || !y || !energy))) return ABM_E_BADARG;
  std::vector<std::vector<int64_t>> bi(g->n), bx(g->n), by(g->n);
  std::vector<std::vector<double>> be(g->n);
  for (int64_t k = 0; k < n; ++k) {
    if (y[k] < 1 || y[k] > g->cfg.ny) GRP_FAIL(g, ABM_E_BADARG, "agent %lld: y = %lld outside the grid", (long long)id[k], (long long)y[k]);
    int m = 0;
    while (m + 1 < g->n && y[k] - 1 >= g->y0[m] + g->rows[m]) ++m;
    bi[m].push_back(id[k]); bx[m].push_back(x[k]); by[m].push_back(y[k]); be[m].push_back(energy[k]);
  }
  int rc = group_parallel(g, [&](int i) -> int {
    return ABM_API(set_agents)(g->h[i], (int64_t)bi[i].size(), bi[i].data(), bx[i].data(), by[i].data(), be[i].data(), next_id);
  });
  if (rc) return rc;
  return group_connect(g);
}

int ABM_API(group_init_random)(abm_group* g, int64_t n_sheep, uint64_t init_seed) {
  if (!g) return ABM_E_BADARG;
  int rc = group_parallel(g, [&](int i) -> int { return ABM_API(init_random)(g->h[i], n_sheep, init_seed); });
  if (rc) return rc;
  return group_connect(g);
}

int ABM_API(group_step)(abm_group* g, int64_t n_ticks) {
  if (!g) return ABM_E_BADARG;
  return group_parallel(g, [&](int i) -> int { return ABM_API(step)(g->h[i], n_ticks); });
}

int ABM_API(group_set_global_state)(abm_group* g, int64_t tick, int64_t behav, int64_t behav_counter) {
  if (!g) return ABM_E_BADARG;
  return group_parallel(g, [&](int i) -> int { return ABM_API(set_global_state)(g->h[i], tick, behav, behav_counter); });
}
int ABM_API(group_get_global_state)(abm_group* g, int64_t* tick, int64_t* behav, int64_t* behav_counter, int64_t* next_id) {
  if (!g) return ABM_E_BADARG;
  return ABM_API(get_global_state)(g->h[0], tick, behav, behav_counter, next_id); /* the same on every member */
}

int ABM_API(group_count_agents)(abm_group* g, int64_t* n) {
  if (!g || !n) return ABM_E_BADARG;
  *n = 0;
  for (int i = 0; i < g->n; ++i) *n += (int64_t)g->h[i]->n;
  return ABM_OK;
}

/* collective inside: every member contributes its share and returns the value over the whole GridSpace */
int ABM_API(group_reduce)(abm_group* g, int32_t what, double* out) {
  if (!g || !out) return ABM_E_BADARG;
  std::vector<double> v((size_t)g->n, 0.0);
  const int rc = group_parallel(g, [&](int i) -> int { return ABM_API(reduce)(g->h[i], what, &v[i]); });
  if (rc) return rc;
  *out = v[0];
  return ABM_OK;
}

/* all agents of the GridSpace in id order */
int ABM_API(group_get_agents)(abm_group* g, int64_t* n_inout, int64_t* id, int64_t* x, int64_t* y, double* energy) {
  if (!g || !n_inout) return ABM_E_BADARG;
  int64_t total = 0;
  std::vector<int64_t> off((size_t)g->n + 1, 0);
  for (int i = 0; i < g->n; ++i) { off[i] = total; total += (int64_t)g->h[i]->n; }
  off[g->n] = total;
  if (*n_inout < total) { *n_inout = total; GRP_FAIL(g, ABM_E_SMALL, "agent buffers hold fewer than %lld entries", (long long)total); }
  *n_inout = total;
  if (total == 0) return ABM_OK;
  std::vector<int64_t> ti((size_t)total), tx((size_t)total), ty((size_t)total);
  std::vector<double> te((size_t)total);
  int rc = group_parallel(g, [&](int i) -> int {
    int64_t cap = off[i + 1] - off[i];
    return ABM_API(get_agents)(g->h[i], &cap, ti.data() + off[i], tx.data() + off[i], ty.data() + off[i], te.data() + off[i]);
  });
  if (rc) return rc;
  std::vector<int64_t> order((size_t)total);
  for (int64_t k = 0; k < total; ++k) order[k] = k;
  std::sort(order.begin(), order.end(), [&](int64_t a, int64_t b) { return ti[a] < ti[b]; });
  for (int64_t k = 0; k < total; ++k) {
    const int64_t s = order[k];
    if (id) id[k] = ti[s];
    if (x) x[k] = tx[s];
    if (y) y[k] = ty[s];
    if (energy) energy[k] = te[s];
  }
  return ABM_OK;
}

int ABM_API(group_get_grid)(abm_group* g, uint8_t* fg, int64_t* cd) {
  if (!g) return ABM_E_BADARG;
  const int nx = g->cfg.nx;
  return group_parallel(g, [&](int i) -> int {
    const size_t cells = (size_t)(g->rows[i] + 2 * CORE) * nx;
    std::vector<uint8_t> sfg(fg ? cells : 0); std::vector<int64_t> scd(cd ? cells : 0);
    const int rc = ABM_API(get_grid)(g->h[i], fg ? sfg.data() : nullptr, cd ? scd.data() : nullptr);
    if (rc) return rc;
    const size_t own = (size_t)g->rows[i] * nx, skip = (size_t)CORE * nx, at = (size_t)g->y0[i] * nx;
    if (fg) memcpy(fg + at, sfg.data() + skip, own);
    if (cd) memcpy(cd + at, scd.data() + skip, own * 8);
    return 0;
  });
}

}  // extern "C"
