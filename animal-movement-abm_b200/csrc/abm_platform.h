/*
 * abm_platform.h — the one place that knows whether the engine sources are being compiled by nvcc for sm_100a
 * (the product, libabm.so) or by g++ with -DABM_EMU (tests/emu only).
 *
 * ABM_EMU is NOT a CPU fallback of the product: libabm.so is never built with it, and nothing in the package
 * loads the emulation library.  It exists because the build container has no GPU: the kernel bodies are plain
 * functors over raw pointers, so the `-m "not gpu"` tests run the very same bodies and the very same host
 * orchestration (allocation, launch order, worklists, re-binning) with a sequential "launcher" that visits the
 * thread indices in a shuffled order.  That catches logic and ordering bugs before GPU minutes are spent.
 */
#ifndef ABM_PLATFORM_H_
#define ABM_PLATFORM_H_

#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#ifdef ABM_EMU
/* ------------------------------------------------------------------ emulation (tests only) --------------- */
#define ABM_HD inline
#define ABM_D inline
#define ABM_HOSTDEV inline
#include <algorithm>
#include <vector>

typedef int abm_stream_t;
typedef int abm_err_t;

namespace abm {

static inline uint32_t atomic_add(uint32_t* p, uint32_t v) { uint32_t o = *p; *p = o + v; return o; }
static inline uint32_t atomic_or(uint32_t* p, uint32_t v) { uint32_t o = *p; *p = o | v; return o; }
static inline uint32_t atomic_cas(uint32_t* p, uint32_t cmp, uint32_t v) { uint32_t o = *p; if (o == cmp) *p = v; return o; }
static inline uint32_t atomic_max(uint32_t* p, uint32_t v) { uint32_t o = *p; if (v > o) *p = v; return o; }
static inline uint32_t atomic_exch(uint32_t* p, uint32_t v) { uint32_t o = *p; *p = v; return o; }
static inline unsigned long long atomic_add64(unsigned long long* p, unsigned long long v) { unsigned long long o = *p; *p = o + v; return o; }
static inline uint8_t load_state(const uint8_t* p) { return *(const volatile uint8_t*)p; }
static inline int popc(uint32_t v) { return __builtin_popcount(v); }
/* position of the (k + 1)-th set bit of m (k < popc(m)) */
static inline int nth_set_bit(uint32_t m, int k) { int d = 0; for (;; m >>= 1, ++d) { if (m & 1u) { if (k == 0) return d; --k; } } }

struct EmuRand {
  uint64_t s;
  uint32_t next() { s = s * 6364136223846793005ull + 1442695040888963407ull; return (uint32_t)(s >> 33); }
};
extern uint64_t g_emu_launches;
extern uint64_t g_emu_shuffle_seed;

/* visits i in [0,n) in a pseudo-random order (block-shuffled so that big launches stay cheap) */
template <class F>
static inline void launch_foreach(abm_stream_t, const F& f, uint64_t n, const char* = 0) {
  ++g_emu_launches;
  if (n == 0) return;
  EmuRand r = {g_emu_shuffle_seed + n * 0x9E3779B97F4A7C15ull + g_emu_launches};
  const uint64_t B = 61; /* prime block; blocks visited in a strided permutation, lanes reversed at random */
  const uint64_t nb = (n + B - 1) / B;
  uint64_t stride = 1;
  if (nb > 2) { stride = (r.next() % (nb - 1)) + 1; while (std::__gcd(stride, nb) != 1) ++stride; }
  uint64_t b = r.next() % nb;
  for (uint64_t k = 0; k < nb; ++k, b = (b + stride) % nb) {
    const uint64_t lo = b * B, hi = std::min(n, lo + B);
    if (r.next() & 1) { for (uint64_t i = hi; i-- > lo;) f(i); }
    else { for (uint64_t i = lo; i < hi; ++i) f(i); }
  }
}

/* CTA-structured kernels: the body is a sequence of phases `ABM_CTA_FOR(t, NT) { ... } ABM_CTA_SYNC();` — on the
 * device every thread runs each phase once (t = threadIdx.x) and phases are separated by __syncthreads(); here a
 * phase is a loop over the thread indices (visited forwards or backwards), and CTAs run one after another in a
 * shuffled order.  Code outside the phases must be uniform across the CTA. */
#ifdef ABM_EMU_REVERSE /* tools/sanitize.sh: the threads of a phase visited backwards — a result that changes is a race on the device */
#define ABM_CTA_FOR(t, NT) for (int t = (int)(NT) - 1; t >= 0; --t)
#else
#define ABM_CTA_FOR(t, NT) for (int t = 0; t < (int)(NT); ++t)
#endif
#define ABM_CTA_SYNC() ((void)0)
/* a stretch of phases run by the first warp alone: `if (ABM_IN_WARP0) { ABM_WARP0_FOR(t) { ... } ABM_WARP_SYNC(); ... }` */
#define ABM_IN_WARP0 true
#ifdef ABM_EMU_REVERSE
#define ABM_WARP0_FOR(t) for (int t = 31; t >= 0; --t)
#else
#define ABM_WARP0_FOR(t) for (int t = 0; t < 32; ++t)
#endif
#define ABM_WARP_SYNC() ((void)0)
#define ABM_PHASE_BEGIN() ((void)0)
#define ABM_PHASE(k) ((void)0)
static inline uint32_t smem_min(uint32_t* p, uint32_t v) { uint32_t o = *p; if (v < o) *p = v; return o; }
static inline uint32_t smem_max(uint32_t* p, uint32_t v) { uint32_t o = *p; if (v > o) *p = v; return o; }
static inline uint32_t smem_add(uint32_t* p, uint32_t v) { uint32_t o = *p; *p = o + v; return o; }
static inline void warp_min64_into(unsigned long long* dst, unsigned long long v, int /*lane*/) { if (v < *dst) *dst = v; }
static inline uint32_t smem_or(uint32_t* p, uint32_t v) { uint32_t o = *p; *p = o | v; return o; }
/* minimum over the 32 lanes of a warp that all execute this call, folded into *dst; returns the minimum known to the
 * caller (lanes run one after another here, so that is the running value) */
static inline uint32_t warp_min_into(uint32_t* dst, uint32_t v, int /*lane*/) { if (v < *dst) *dst = v; return *dst; }
/* true when the predicate holds on every lane of the warp (lanes run one after another here: the lane's own answer) */
static inline bool warp_all(bool pred) { return pred; }
/* lanes that share one gregarious search (abm_tile.h): one here, so group_min is the identity */
enum { GREG_LANES = 1 };
static inline uint32_t group_min(uint32_t v) { return v; }
/* base[key] += 1 for every calling lane with valid set; returns the lane's own slot (see the CUDA version) */
static inline uint32_t agg_add(uint32_t* base, uint32_t key, bool valid) { return valid ? atomic_add(&base[key], 1u) : 0u; }

/* peer-memory flags (abm_p2p.h) do not exist in the emulation: transports there go through the host */
static inline void flag_publish(uint32_t*, uint32_t) {}
static inline void flag_fence() {}
static inline void flag_wait(const uint32_t*, uint32_t, uint32_t*, uint32_t) {}

template <class F>
static inline void launch_cta(abm_stream_t, const F& f, uint64_t n_ctas, int /*threads*/, size_t smem_bytes) {
  ++g_emu_launches;
  if (n_ctas == 0) return;
  std::vector<uint32_t> smem((smem_bytes + 3) / 4 + 1);
  EmuRand r = {g_emu_shuffle_seed * 31 + n_ctas * 0x9E3779B97F4A7C15ull + g_emu_launches};
  uint64_t stride = 1;
  if (n_ctas > 2) { stride = (r.next() % (n_ctas - 1)) + 1; while (std::__gcd(stride, n_ctas) != 1) ++stride; }
  uint64_t b = r.next() % n_ctas;
  for (uint64_t k = 0; k < n_ctas; ++k, b = (b + stride) % n_ctas) {
    std::fill(smem.begin(), smem.end(), 0xA5A5A5A5u); /* shared memory starts uninitialised */
    f((uint32_t)b, smem.data());
  }
}

template <int THREADS, int MINB, class F>
static inline void launch_cta_lb(abm_stream_t s, const F& f, uint64_t n_ctas, size_t smem_bytes) { launch_cta(s, f, n_ctas, THREADS, smem_bytes); }

static inline abm_err_t dev_malloc(void** p, size_t bytes) { *p = calloc(bytes ? bytes : 1, 1); return *p ? 0 : 2; }
static inline void dev_free(void* p) { free(p); }
static inline abm_err_t dev_memset(void* p, int v, size_t bytes, abm_stream_t) { memset(p, v, bytes); return 0; }
static inline abm_err_t copy_h2d(void* d, const void* h, size_t bytes, abm_stream_t) { memcpy(d, h, bytes); return 0; }
static inline abm_err_t copy_d2h(void* h, const void* d, size_t bytes, abm_stream_t) { memcpy(h, d, bytes); return 0; }
static inline abm_err_t copy_d2d(void* d, const void* s, size_t bytes, abm_stream_t) { memmove(d, s, bytes); return 0; }
static inline abm_err_t copy_h2d_async(void* d, const void* h, size_t bytes, abm_stream_t) { memcpy(d, h, bytes); return 0; }
static inline abm_err_t copy_d2h_async(void* h, const void* d, size_t bytes, abm_stream_t) { memcpy(h, d, bytes); return 0; }
static inline abm_err_t stream_sync(abm_stream_t) { return 0; }
/* a second stream and the events that order it against the first: the emulation runs everything in program order */
typedef int abm_event_t;
static inline abm_err_t side_stream_create(abm_stream_t* s) { *s = 0; return 0; }
static inline void side_stream_destroy(abm_stream_t) {}
static inline abm_err_t event_create(abm_event_t* e) { *e = 0; return 0; }
static inline void event_destroy(abm_event_t) {}
static inline void event_record(abm_event_t, abm_stream_t) {}
static inline void stream_wait_event(abm_stream_t, abm_event_t) {}
static inline const char* err_string(abm_err_t e) { return e ? "emulated allocation failure" : "ok"; }

}  // namespace abm

#else
/* ------------------------------------------------------------------ CUDA, sm_100a (the product) ---------- */
#include <cuda_runtime.h>
#include <atomic>
#define ABM_HD __device__ __forceinline__
#define ABM_D __device__ __forceinline__
#define ABM_HOSTDEV __host__ __device__ __forceinline__

typedef cudaStream_t abm_stream_t;
typedef cudaError_t abm_err_t;

namespace abm {

ABM_D uint32_t atomic_add(uint32_t* p, uint32_t v) { return atomicAdd(p, v); }
ABM_D uint32_t atomic_or(uint32_t* p, uint32_t v) { return atomicOr(p, v); }
ABM_D uint32_t atomic_cas(uint32_t* p, uint32_t cmp, uint32_t v) { return atomicCAS(p, cmp, v); }
ABM_D uint32_t atomic_max(uint32_t* p, uint32_t v) { return atomicMax(p, v); }
ABM_D uint32_t atomic_exch(uint32_t* p, uint32_t v) { return atomicExch(p, v); }
ABM_D unsigned long long atomic_add64(unsigned long long* p, unsigned long long v) { return atomicAdd(p, v); }
/* per-agent decision bytes are written once by their owner and polled by neighbours in the same launch:
 * bypass L1 so that a decision published by another SM becomes visible (ld.global.cg = cache at L2 only) */
ABM_D uint8_t load_state(const uint8_t* p) { return __ldcg(p); }
ABM_D int popc(uint32_t v) { return __popc(v); }
/* position of the (k + 1)-th set bit of m (k < popc(m)) */
ABM_D int nth_set_bit(uint32_t m, int k) { /* m has at most 8 bits (a Moore neighbourhood): three popcount steps, a third of __fns */
  int d = 0;
  int c = __popc(m & 0xFu); if (k >= c) { k -= c; d = 4; m >>= 4; }
  c = __popc(m & 0x3u); if (k >= c) { k -= c; d += 2; m >>= 2; }
  c = (int)(m & 1u); if (k >= c) d += 1;
  return d;
}

extern unsigned long long g_launches;

/* -DABM_PHASE_CLOCKS (tuning builds only, tools/phase_probe.py): SM cycles between the marks of a CTA-structured kernel,
 * taken by thread 0 and summed over all CTAs into abm_phase_clk[] (read and reset through abm_debug_phase_clocks) */
#ifdef ABM_PHASE_CLOCKS
__device__ unsigned long long abm_phase_clk[32];
#define ABM_PHASE_BEGIN() long long abm_phase_t0_ = clock64()
#define ABM_PHASE(k) do { if (threadIdx.x == 0) { const long long c_ = clock64(); atomicAdd(&abm_phase_clk[k], (unsigned long long)(c_ - abm_phase_t0_)); abm_phase_t0_ = c_; } } while (0)
#else
#define ABM_PHASE_BEGIN() ((void)0)
#define ABM_PHASE(k) ((void)0)
#endif

template <class F>
__global__ void __launch_bounds__(256) k_foreach(const F f, uint64_t n) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) f(i);
}

template <class F>
static inline void launch_foreach(abm_stream_t s, const F& f, uint64_t n, const char* = 0) {
  if (n == 0) return;
  ++g_launches;
  const uint64_t blocks = (n + 255) / 256;
  k_foreach<F><<<(unsigned)blocks, 256, 0, s>>>(f, n);
}

#define ABM_CTA_FOR(t, NT) for (int t = (int)threadIdx.x, abm_once_ = 1; abm_once_ && t < (int)(NT); abm_once_ = 0)
#define ABM_CTA_SYNC() __syncthreads()
#define ABM_IN_WARP0 (threadIdx.x < 32)
#define ABM_WARP0_FOR(t) for (int t = (int)threadIdx.x, abm_once_ = 1; abm_once_; abm_once_ = 0)
#define ABM_WARP_SYNC() __syncwarp()
ABM_D uint32_t smem_min(uint32_t* p, uint32_t v) { return atomicMin(p, v); }
ABM_D uint32_t smem_max(uint32_t* p, uint32_t v) { return atomicMax(p, v); }
ABM_D uint32_t smem_add(uint32_t* p, uint32_t v) { return atomicAdd(p, v); }
/* 64-bit minimum over the 32 converged lanes of a warp (two 32-bit reductions), folded into *dst by lane 0 */
ABM_D void warp_min64_into(unsigned long long* dst, unsigned long long v, int lane) {
  const uint32_t hi = (uint32_t)(v >> 32);
  const uint32_t mhi = __reduce_min_sync(0xffffffffu, hi);
  const uint32_t mlo = __reduce_min_sync(0xffffffffu, hi == mhi ? (uint32_t)v : 0xFFFFFFFFu);
  const unsigned long long m = ((unsigned long long)mhi << 32) | mlo;
  if (lane == 0 && m < *dst) *dst = m;
}
ABM_D uint32_t smem_or(uint32_t* p, uint32_t v) { return atomicOr(p, v); }
/* minimum over the 32 lanes of a warp that all execute this call (converged), folded into *dst by lane 0 */
ABM_D uint32_t warp_min_into(uint32_t* dst, uint32_t v, int lane) {
  const uint32_t m = __reduce_min_sync(0xffffffffu, v);
  if (lane == 0 && m < *dst) *dst = m;
  return m; /* the same value in every lane */
}

/* GREG_LANES consecutive lanes share one gregarious search (abm_tile.h); their minimum, in every lane of the group */
enum { GREG_LANES = 4 };
ABM_D uint32_t group_min(uint32_t v) {
  v = min(v, __shfl_xor_sync(0xffffffffu, v, 1));
  v = min(v, __shfl_xor_sync(0xffffffffu, v, 2));
  return v;
}
/* true when the predicate holds on every lane of the warp; every lane must call it (it also reconverges the warp) */
ABM_D bool warp_all(bool pred) { return __all_sync(0xffffffffu, pred) != 0; }
/* base[key] += 1 for every lane with `valid`, one atomic per distinct key in the warp; returns the lane's own slot.
 * Every lane of the warp must call it (the *_sync intrinsics wait for all 32). */
ABM_D uint32_t agg_add(uint32_t* base, uint32_t key, bool valid) {
  const uint32_t lane = threadIdx.x & 31u;
  const uint32_t peers = __match_any_sync(0xffffffffu, valid ? key : 0xFFFFFFFFu);
  const int leader = __ffs((int)peers) - 1;
  uint32_t b = 0;
  if (valid && (int)lane == leader) b = atomicAdd(&base[key], (uint32_t)__popc(peers));
  b = __shfl_sync(peers, b, leader);
  return b + (uint32_t)__popc(peers & ((1u << lane) - 1u));
}

/* publish a sequence number in (peer) memory after everything this thread's CTA wrote before the preceding barrier */
ABM_D void flag_publish(uint32_t* flag, uint32_t seq) {
  if (!flag) return;
  __threadfence_system();
  *(volatile uint32_t*)flag = seq;
}
ABM_D void flag_fence() { __threadfence_system(); }
/* spin until the flag reaches seq; a lost neighbour becomes an error bit after abm_flag_timeout_cycles (~30 s by default,
 * ABM_P2P_TIMEOUT_S at abm_p2p_connect / abm_group_create; the handle is then unusable), not a hang */
__device__ long long abm_flag_timeout_cycles = 60000000000ll;
ABM_D void flag_wait(const uint32_t* flag, uint32_t seq, uint32_t* err, uint32_t err_bit) {
  if (!flag) return;
  const long long t0 = clock64(), limit = abm_flag_timeout_cycles;
  while ((int32_t)(*(const volatile uint32_t*)flag - seq) < 0) {
    if (clock64() - t0 > limit) { atomicOr(err, err_bit); break; }
    __nanosleep(32);
  }
  __threadfence_system();
}

template <class F>
__global__ void __launch_bounds__(256) k_cta(const F f) {
  extern __shared__ __align__(16) uint32_t abm_smem[];
  f(blockIdx.x, abm_smem);
}

template <class F>
static inline void launch_cta(abm_stream_t s, const F& f, uint64_t n_ctas, int threads, size_t smem_bytes) {
  if (n_ctas == 0) return;
  ++g_launches;
  static std::atomic<unsigned long long> configured{0}; /* per kernel instantiation: one bit per device (the attributes are per device) */
  int dev = 0;
  cudaGetDevice(&dev);
  const unsigned long long bit = 1ull << (dev & 63);
  if (!(configured.load(std::memory_order_relaxed) & bit)) {
    cudaFuncSetAttribute(k_cta<F>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    cudaFuncSetAttribute(k_cta<F>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    configured.fetch_or(bit, std::memory_order_relaxed);
  }
  k_cta<F><<<(unsigned)n_ctas, threads, smem_bytes, s>>>(f);
}

/* the same with the register budget of THREADS x MINB resident threads per SM (occupancy of latency-bound kernels) */
template <class F, int THREADS, int MINB>
__global__ void __launch_bounds__(THREADS, MINB) k_cta_lb(const F f) {
  extern __shared__ __align__(16) uint32_t abm_smem[];
  f(blockIdx.x, abm_smem);
}
template <int THREADS, int MINB, class F>
static inline void launch_cta_lb(abm_stream_t s, const F& f, uint64_t n_ctas, size_t smem_bytes) {
  if (n_ctas == 0) return;
  ++g_launches;
  cudaFuncSetAttribute(k_cta_lb<F, THREADS, MINB>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
  k_cta_lb<F, THREADS, MINB><<<(unsigned)n_ctas, THREADS, smem_bytes, s>>>(f);
}

static inline abm_err_t dev_malloc(void** p, size_t bytes) {
  abm_err_t e = cudaMalloc(p, bytes ? bytes : 1);
  if (e == cudaSuccess) e = cudaMemset(*p, 0, bytes ? bytes : 1);
  /* the handle's stream is non-blocking: make the legacy-stream memset land before anyone writes the buffer.  Only the
   * legacy stream is waited for, not the whole device (other handles' kernels may be running). */
  if (e == cudaSuccess) e = cudaStreamSynchronize(cudaStreamLegacy);
  return e;
}
static inline void dev_free(void* p) { if (p) cudaFree(p); }
static inline abm_err_t dev_memset(void* p, int v, size_t bytes, abm_stream_t s) { return cudaMemsetAsync(p, v, bytes, s); }
static inline abm_err_t copy_h2d(void* d, const void* h, size_t bytes, abm_stream_t s) {
  abm_err_t e = cudaMemcpyAsync(d, h, bytes, cudaMemcpyHostToDevice, s);
  return e != cudaSuccess ? e : cudaStreamSynchronize(s);
}
static inline abm_err_t copy_d2h(void* h, const void* d, size_t bytes, abm_stream_t s) {
  abm_err_t e = cudaMemcpyAsync(h, d, bytes, cudaMemcpyDeviceToHost, s);
  return e != cudaSuccess ? e : cudaStreamSynchronize(s);
}
/* enqueue only: the caller synchronises the stream before it returns to the host program (host buffers are never retained) */
static inline abm_err_t copy_h2d_async(void* d, const void* h, size_t bytes, abm_stream_t s) { return cudaMemcpyAsync(d, h, bytes, cudaMemcpyHostToDevice, s); }
static inline abm_err_t copy_d2h_async(void* h, const void* d, size_t bytes, abm_stream_t s) { return cudaMemcpyAsync(h, d, bytes, cudaMemcpyDeviceToHost, s); }
static inline abm_err_t copy_d2d(void* d, const void* src, size_t bytes, abm_stream_t s) {
  return cudaMemcpyAsync(d, src, bytes, cudaMemcpyDeviceToDevice, s);
}
static inline abm_err_t stream_sync(abm_stream_t s) { return cudaStreamSynchronize(s); }
/* a second, higher-priority stream for small work that overlaps the tick's big kernels, and the events that order it */
typedef cudaEvent_t abm_event_t;
static inline abm_err_t side_stream_create(abm_stream_t* s) {
  int lo = 0, hi = 0;
  cudaDeviceGetStreamPriorityRange(&lo, &hi);
  return cudaStreamCreateWithPriority(s, cudaStreamNonBlocking, hi);
}
static inline void side_stream_destroy(abm_stream_t s) { if (s) { cudaStreamSynchronize(s); cudaStreamDestroy(s); } }
static inline abm_err_t event_create(abm_event_t* e) { return cudaEventCreateWithFlags(e, cudaEventDisableTiming); }
static inline void event_destroy(abm_event_t e) { if (e) cudaEventDestroy(e); }
static inline void event_record(abm_event_t e, abm_stream_t s) { cudaEventRecord(e, s); }
static inline void stream_wait_event(abm_stream_t s, abm_event_t e) { cudaStreamWaitEvent(s, e, 0); }
static inline const char* err_string(abm_err_t e) { return cudaGetErrorString(e); }

}  // namespace abm
#endif

#endif /* ABM_PLATFORM_H_ */
