/*
 * abm_p2p.h — ring exchange and tiny collectives of the row strips as plain kernels over NVLink peer memory.
 *
 * No reference counterpart (the reference is one process).  Each rank owns a "mailbox" in device memory that the
 * other ranks of the node map through CUDA IPC.  A sender's kernel stores its message straight into the receiver's
 * mailbox and then publishes a sequence number there; the receiver's next kernel spins on that number (with a
 * time-out) and reads the data locally.  Every channel is used symmetrically once per round (both sides send, then
 * wait), so two buffers per channel, alternating with the sequence number, are enough: a sender can only reach
 * round s + 2 after it has seen the receiver's round s + 1, which the receiver issued behind its reads of round s.
 * Compared with NCCL send/recv + all-reduce this removes the host-side cost of ~5 NCCL groups per tick.  The pack and
 * merge CTAs of abm_tile.h store into / wait on the mailboxes themselves; this file holds the layout and the two
 * all-to-all collectives (counters, offspring).
 */
#ifndef ABM_P2P_H_
#define ABM_P2P_H_
#ifndef ABM_EMU

#include "abm_kernels.h"

namespace abm {

enum { P2P_MAXP = 8, P2P_STATS_WORDS = 32, P2P_FLAG_STRIDE = 32 /* u32: one 128-byte line per flag */ };
enum { P2P_CH_AGENTS = 0, P2P_CH_STATE = 1, P2P_CH_STATS = 2, P2P_CH_BIRTHS = 3 };

struct P2PLayout {
  uint64_t agents_bytes, state_bytes, births_bytes; /* capacity of one buffer of each kind */
  uint64_t flag_off(int ch, int slot, int parity) const { return (uint64_t)(((ch * P2P_MAXP + slot) * 2 + parity) * P2P_FLAG_STRIDE) * 4; }
  uint64_t flags_bytes() const { return (uint64_t)(4 * P2P_MAXP * 2 * P2P_FLAG_STRIDE) * 4; }
  uint64_t agents_off(int dir, int parity) const { return flags_bytes() + (uint64_t)(dir * 2 + parity) * agents_bytes; }
  uint64_t state_off(int dir, int parity) const { return agents_off(2, 0) + (uint64_t)(dir * 2 + parity) * state_bytes; }
  uint64_t stats_off(int src, int parity) const { return state_off(2, 0) + (uint64_t)(src * 2 + parity) * P2P_STATS_WORDS * 8; }
  uint64_t births_off(int src, int parity) const { return stats_off(P2P_MAXP, 0) + (uint64_t)(src * 2 + parity) * births_bytes; }
  uint64_t total_bytes() const { return births_off(P2P_MAXP, 0); }
};

ABM_D uint32_t ld_flag(const uint32_t* p) { return *(const volatile uint32_t*)p; }

/* The per-round counters (StripStatsBody's vector) as an all-to-all in two launches.  Send: one CTA computes this rank's
 * vector, stores it into every rank's mailbox (its own included) and publishes.  Receive: one CTA waits for every rank's
 * vector, sums them into `out`, and (optionally) evaluates the commit guard from the sums. */
struct P2PStatsSendCta {
  const uint32_t* pending; const uint32_t* birth_flags; const uint32_t* tot_up; const uint32_t* tot_down;
  int32_t n_ranks, rank;
  unsigned long long* dst[P2P_MAXP]; uint32_t* flag[P2P_MAXP]; uint32_t seq;
  ABM_HD void operator()(uint32_t, uint32_t* smem) const {
    const int P = n_ranks, nv = 1 + 3 * P;
    unsigned long long* vec = (unsigned long long*)smem;
    ABM_CTA_FOR(t, 64) {
      if (t < nv) vec[t] = 0;
    }
    ABM_CTA_SYNC();
    ABM_CTA_FOR(t, 64) {
      if (t == 0) {
        vec[0] = *pending; vec[1 + rank] = *pending; vec[1 + P + rank] = *birth_flags;
        const uint32_t a = *tot_up, b = *tot_down;
        vec[1 + 2 * P + rank] = a > b ? a : b;
      }
    }
    ABM_CTA_SYNC();
    ABM_CTA_FOR(t, 64) { /* thread r serves rank r: its own stores, its own fence, its own flag */
      if (t < P) {
        for (int w = 0; w < nv; ++w) dst[t][w] = vec[w];
        flag_publish(flag[t], seq);
      }
    }
  }
};
struct P2PStatsRecvCta {
  const unsigned long long* src[P2P_MAXP]; const uint32_t* flag[P2P_MAXP]; uint32_t seq;
  int32_t n_ranks; unsigned long long* out; uint32_t* err;
  uint32_t* go; uint32_t birth_cap; /* optional guard (StripGuardBody) */
  ABM_HD void operator()(uint32_t, uint32_t*) const {
    const int P = n_ranks, nv = 1 + 3 * P;
    ABM_CTA_FOR(t, 64) { if (t < P) flag_wait(flag[t], seq, err, ERR_P2P_TIMEOUT); }
    ABM_CTA_SYNC();
    ABM_CTA_FOR(t, 64) {
      if (t < nv) { unsigned long long s = 0; for (int q = 0; q < P; ++q) s += src[q][t]; out[t] = s; }
    }
    ABM_CTA_SYNC();
    ABM_CTA_FOR(t, 64) {
      if (t == 0 && go) {
        bool ok = out[0] == 0;
        for (int r = 0; r < P; ++r) ok = ok && out[1 + P + r] <= birth_cap;
        *go = ok ? 1u : 0u;
      }
    }
  }
};

/* a few u64 per rank, all-to-all over the counters channel (abm_reduce on strips): thread r serves rank r */
struct P2PSmallSendCta {
  const unsigned long long* src; int32_t n, n_ranks;
  unsigned long long* dst[P2P_MAXP]; uint32_t* flag[P2P_MAXP]; uint32_t seq;
  ABM_HD void operator()(uint32_t, uint32_t*) const {
    ABM_CTA_FOR(t, 64) {
      if (t < n_ranks) { for (int w = 0; w < n; ++w) dst[t][w] = src[w]; flag_publish(flag[t], seq); }
    }
  }
};
struct P2PSmallRecvCta {
  const unsigned long long* src[P2P_MAXP]; const uint32_t* flag[P2P_MAXP]; uint32_t seq;
  int32_t n, n_ranks; unsigned long long* out; uint32_t* err;
  ABM_HD void operator()(uint32_t, uint32_t*) const {
    ABM_CTA_FOR(t, 64) { if (t < n_ranks) flag_wait(flag[t], seq, err, ERR_P2P_TIMEOUT); }
    ABM_CTA_SYNC();
    ABM_CTA_FOR(t, 64) { if (t < n_ranks) for (int w = 0; w < n; ++w) out[t * n + w] = src[t][w]; }
  }
};

/* offspring all-gather in two launches of n_ranks x BIRTH_CTAS CTAs: the CTAs (r, *) store [count, parent ids ...] of
 * this rank into rank r's mailbox (coalesced 128-byte rows over NVLink) and the one that finishes last publishes; then
 * the CTAs (r, *) wait for rank r's message and lay it out contiguously for the rank kernel */
enum { BIRTH_CTAS = 32 };
struct P2PBirthSendCta {
  const uint32_t* birth_parent; const uint32_t* birth_count; uint32_t cap;
  uint32_t* dst[P2P_MAXP]; uint32_t* flag[P2P_MAXP]; uint32_t seq;
  uint32_t* done; /* one counter per destination, left at zero */
  ABM_HD void operator()(uint32_t cta, uint32_t*) const {
    const uint32_t r = cta / BIRTH_CTAS, part = cta % BIRTH_CTAS;
    const uint32_t n = *birth_count < cap ? *birth_count : cap;
    ABM_CTA_FOR(t, 256) {
      for (uint32_t w = part * 256u + (uint32_t)t; w < 1 + n; w += 256u * BIRTH_CTAS) dst[r][w] = w == 0 ? *birth_count : birth_parent[w - 1];
    }
    ABM_CTA_SYNC();
    ABM_CTA_FOR(t, 256) {
      if (t == 0) {
        flag_fence();
        if (atomic_add(&done[r], 1u) == BIRTH_CTAS - 1) { done[r] = 0; flag_publish(flag[r], seq); }
      }
    }
  }
};
struct P2PBirthRecvCta {
  const uint32_t* src[P2P_MAXP]; const uint32_t* flag[P2P_MAXP]; uint32_t seq, cap; uint32_t* all; uint32_t* err;
  ABM_HD void operator()(uint32_t cta, uint32_t*) const {
    const uint32_t r = cta / BIRTH_CTAS, part = cta % BIRTH_CTAS;
    ABM_CTA_FOR(t, 256) { if (t == 0) flag_wait(flag[r], seq, err, ERR_P2P_TIMEOUT); }
    ABM_CTA_SYNC();
    ABM_CTA_FOR(t, 256) {
      const uint32_t n = src[r][0] < cap ? src[r][0] : cap;
      for (uint32_t w = part * 256u + (uint32_t)t; w < 1 + n; w += 256u * BIRTH_CTAS) all[(size_t)r * (1 + cap) + w] = src[r][w];
    }
  }
};

}  // namespace abm
#endif /* ABM_EMU */
#endif /* ABM_P2P_H_ */
