/*
 * abm_rng.h — counter-based randomness of the engine (contract stated in include/abm.h).
 *
 * Replaces the reference's `model.rng` (MersenneTwister, scripts/v0.1.jl:49) and the global-RNG call sites
 * (src/agent_actions.jl:213 `sample(1:3)`, :328 `wsample`) by Philox4x32-10 keyed (seed; tick, agent id, draw
 * index), so that a draw no longer depends on how many agents were stepped before.
 * Sampler semantics follow Julia 1.9.1 `Random` for a generic AbstractRNG.
 */
#ifndef ABM_RNG_H_
#define ABM_RNG_H_

#include "abm_platform.h"

namespace abm {

ABM_HOSTDEV uint32_t mulhi32(uint32_t a, uint32_t b) {
#if defined(__CUDA_ARCH__)
  return __umulhi(a, b);
#else
  return (uint32_t)(((uint64_t)a * b) >> 32);
#endif
}

ABM_HD uint64_t mulhi64(uint64_t a, uint64_t b) {
#if defined(__CUDA_ARCH__)
  return __umul64hi(a, b);
#else
  return (uint64_t)(((unsigned __int128)a * b) >> 64);
#endif
}

/* Philox4x32-10 (multipliers/Weyl constants as in curand_philox4x32_x.h:88-91) */
ABM_HOSTDEV void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1,
                          uint32_t out[4]) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t h0 = mulhi32(0xD2511F53u, c0), l0 = 0xD2511F53u * c0;
    const uint32_t h1 = mulhi32(0xCD9E8D57u, c2), l1 = 0xCD9E8D57u * c2;
    c0 = h1 ^ c1 ^ k0;
    c1 = l1;
    c2 = h0 ^ c3 ^ k1;
    c3 = l0;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

/* The positional per-(tick, agent) stream: the n-th `rand(rng, UInt64)` of this agent's step. */
struct AgentStream {
  uint32_t k0, k1, tick, id_lo, id_hi;
  uint32_t n;      /* next draw index */
  uint32_t blk;    /* block cached in w[], 0xffffffff = none */
  uint32_t w[4];

  ABM_HD void init(uint64_t seed, uint64_t tick_, uint64_t id) {
    k0 = (uint32_t)seed; k1 = (uint32_t)(seed >> 32);
    tick = (uint32_t)tick_;
    id_lo = (uint32_t)id; id_hi = (uint32_t)(id >> 32);
    n = 0; blk = 0xffffffffu;
  }
  ABM_HD uint64_t next_u64() {
    const uint32_t b = n >> 1, h = n & 1;
    if (b != blk) { philox4x32_10(b, tick, id_lo, id_hi, k0, k1, w); blk = b; }
    ++n;
    return h ? ((uint64_t)w[2] | ((uint64_t)w[3] << 32)) : ((uint64_t)w[0] | ((uint64_t)w[1] << 32));
  }
  /* rand(rng) :: Float64 in [0,1): low 52 bits as the mantissa of [1,2), minus 1 */
  ABM_HD double next_f64() {
    const uint64_t bits = 0x3ff0000000000000ull | (next_u64() & 0x000fffffffffffffull);
#if defined(__CUDA_ARCH__)
    return __longlong_as_double((long long)bits) - 1.0;
#else
    double d; memcpy(&d, &bits, 8); return d - 1.0;
#endif
  }
  /* rand(rng, 1:s): SamplerRangeNDL (Lemire); returns 0-based value in [0, s) */
  ABM_HD uint32_t next_below(uint64_t s) {
    uint64_t x = next_u64();
    uint64_t l = x * s;
    if (l < s) {
      const uint64_t t = (0 - s) % s;
      while (l < t) { x = next_u64(); l = x * s; }
    }
    return (uint32_t)mulhi64(x, s);
  }
};

/*
 * Schedulers.randomly (the scheduler every reference script selects, scripts/v0.1.jl:63) under the counter-based
 * contract of include/abm.h: in tick t the agents are stepped in ascending ORDER KEY, where key = P_t(id) and P_t is a
 * keyed permutation of [0, 2^b) — a 4-round balanced Feistel network over b = 2 hb bits, b the smallest even number
 * >= 2 with 2^b >= nextid(model) at the start of the tick (every id is below it), round keys = the four words of the
 * Philox block (seed; counter 0, tick, reserved agent id 2^64 - 1).  Schedulers.by_id: key = id.  Keys are unique and
 * below 2^b, so everything the engine indexes by id for ORDER (bitmaps, ranks, the min/max summaries) works on keys.
 */
struct SchedKey {
  int32_t on, hb;   /* on = 0: by_id */
  uint32_t k[4];
};
ABM_HOSTDEV int sched_half_bits(uint64_t next_id) { /* hb: 2^(2 hb) >= next_id */
  int hb = 1;
  while (hb < 16 && (1ull << (2 * hb)) < next_id) ++hb;
  return hb;
}
ABM_HOSTDEV void sched_round_keys(uint64_t seed, uint64_t tick, uint32_t k[4]) {
  philox4x32_10(0u, (uint32_t)tick, 0xFFFFFFFFu, 0xFFFFFFFFu, (uint32_t)seed, (uint32_t)(seed >> 32), k);
}
ABM_HOSTDEV uint32_t sched_mix(uint32_t z) {
  z *= 0x9E3779B1u; z ^= z >> 15; z *= 0x85EBCA77u; z ^= z >> 13;
  return z;
}
ABM_HOSTDEV uint32_t order_key(const SchedKey& sk, uint32_t id) {
  if (!sk.on) return id;
  const uint32_t m = (1u << sk.hb) - 1u;
  uint32_t L = id >> sk.hb, R = id & m;
#pragma unroll
  for (int i = 0; i < 4; ++i) { const uint32_t t = L ^ (sched_mix(R ^ sk.k[i]) & m); L = R; R = t; }
  return (L << sk.hb) | R;
}

}  // namespace abm
#endif /* ABM_RNG_H_ */
