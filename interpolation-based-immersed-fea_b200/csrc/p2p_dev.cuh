// p2p_dev.cuh — device-side helpers of the NVLink peer-memory reductions (see p2p.cu).
#pragma once
#include "common.cuh"

namespace iife {

constexpr long long P2P_SPIN_LIMIT = 4000000000LL;

struct PeerTable {
  double *xbuf[P2P_MAX_RANKS];
  Mailbox *mbox[P2P_MAX_RANKS];
  long long dst_start[P2P_MAX_RANKS];
};

// what a fused reduction inside a compute kernel needs
struct P2PRed {
  int enabled;
  int me, nranks;
  Mailbox *mbox;                 // my mailbox
  Mailbox *peer[P2P_MAX_RANKS];  // everybody's mailbox (peer mappings)
  const unsigned long long *iter;  // device iteration counter: reductions 2*it+1 (delta) and 2*it+2 (z.r, z.z), it = *iter + k_off
  int *err;
  int k_off;                     // iteration index inside a graph-captured chunk (the counter moves once per chunk)
  int ll;                        // 1: low-latency packed words (Mailbox::it_ll) instead of values + fence + flag
  unsigned long long *trace;     // IIFE_CG_TRACE: [CG_TRACE_ITERS][CG_TRACE_SLOTS] globaltimer stamps (nullptr: off)
};

// development aid (IIFE_CG_TRACE=1): where the time of a row-partitioned CG iteration goes, in globaltimer nanoseconds
constexpr int CG_TRACE_ITERS = 1024, CG_TRACE_SLOTS = 12;
enum { TR_P_IN = 0, TR_P_WAITED, TR_P_OUT, TR_S_IN, TR_S_WAITED, TR_S_OUT, TR_U_IN, TR_U_WAITED, TR_U_OUT };
__device__ __forceinline__ void cg_trace(const P2PRed &r, int slot) {
  if (r.trace) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    const unsigned long long it_rel = *r.iter + (unsigned long long)r.k_off - r.iter[1];
    r.trace[(it_rel % CG_TRACE_ITERS) * CG_TRACE_SLOTS + slot] = t;
  }
}

// SpMV prologue of the three-kernel CG iteration: wait until the ghost entries of exchange *seq_base + k_off + 1 arrived
struct HaloWait {
  const unsigned long long *flags;     // my mailbox's halo_flag[] (nullptr: no wait)
  const unsigned long long *seq_base;  // device halo sequence counter (moves once per chunk)
  int k_off, nranks;
  unsigned int recv_mask;
  int *err;
  // interior first: slices [int_lo, int_lo + n_int) reference no ghost column and are multiplied before the wait, the
  // others after it, so the NVLink latency of the exchange and the skew between the ranks hide behind the bulk
  int interior_first;
  long long int_lo, n_int;
};

// boundary rows of p and where they go (Halo::brow...)
struct RowPush {
  int n_brow;
  const int *brow, *bptr;
  const unsigned char *bpeer;
  const long long *bdst;
  const unsigned int *bmask;
  PeerTable pt;
  int me, nranks;
  unsigned int send_mask;
  const unsigned long long *seq_base;  // halo sequence counter
  unsigned int *counter;               // CTAs that finished pushing
};

__device__ __forceinline__ void st_flag(unsigned long long *p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_flag(const unsigned long long *p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ bool spin_until(const unsigned long long *p, unsigned long long seq, int *err) {
  long long t0 = clock64();
  while (ld_flag(p) < seq) {
    if (clock64() - t0 > P2P_SPIN_LIMIT) {
      atomicExch(err, 1);
      return false;
    }
  }
  return true;
}

// called by threads 0..nranks-1 of ONE block: store my n partial sums into every rank's mailbox, raise flags
__device__ __forceinline__ void p2p_push(const P2PRed &r, unsigned long long seq, const double *vals, int n, int tid) {
  if (tid < r.nranks) {
    const int par = (int)(seq & 1ull);
    if (r.ll) {
      const unsigned long long tag = (seq & 0xffffffffull) << 32;
      unsigned long long *dst = r.peer[tid]->it_ll[par][r.me];
      for (int i = 0; i < n; ++i) {
        const unsigned long long u = (unsigned long long)__double_as_longlong(vals[i]);
        asm volatile("st.volatile.global.v2.u64 [%0], {%1, %2};" ::"l"(dst + 2 * i), "l"((u & 0xffffffffull) | tag), "l"((u >> 32) | tag)
                     : "memory");
      }
      return;
    }
    for (int i = 0; i < n; ++i) r.peer[tid]->it_vals[par][r.me][i] = vals[i];
    __threadfence_system();
    st_flag(&r.peer[tid]->it_flag[par][r.me], seq);
  }
}

// called by ONE WARP (all 32 lanes): lane q waits for rank q's partials of reduction `seq` and reads them
// (the waits and the remote-written loads proceed in parallel instead of one after the other); lane 0
// returns the sums added in rank order, so every rank gets bit-identical results.
__device__ __forceinline__ void p2p_wait_sum(const P2PRed &r, unsigned long long seq, double *out, int n) {
  const int lane = threadIdx.x & 31;
  const int par = (int)(seq & 1ull);
  double v[4] = {0.0, 0.0, 0.0, 0.0};
  if (lane < r.nranks) {
    if (r.ll) {
      const unsigned long long tag = seq & 0xffffffffull;
      const unsigned long long *src = r.mbox->it_ll[par][lane];
      const long long t0 = clock64();
      for (int i = 0; i < n; ++i) {
        unsigned long long w0, w1;
        for (;;) {
          asm volatile("ld.volatile.global.v2.u64 {%0, %1}, [%2];" : "=l"(w0), "=l"(w1) : "l"(src + 2 * i) : "memory");
          if ((w0 >> 32) == tag && (w1 >> 32) == tag) break;
          if (clock64() - t0 > P2P_SPIN_LIMIT) {
            atomicExch(r.err, 1);
            break;
          }
        }
        v[i] = __longlong_as_double((long long)((w0 & 0xffffffffull) | (w1 << 32)));
      }
    } else {
      spin_until(&r.mbox->it_flag[par][lane], seq, r.err);
      for (int i = 0; i < n; ++i) v[i] = ((volatile double *)r.mbox->it_vals[par][lane])[i];
    }
  }
  for (int i = 0; i < n; ++i) {
    double s = 0.0;
    for (int q = 0; q < r.nranks; ++q) s += __shfl_sync(0xffffffffu, v[i], q);
    out[i] = s;
  }
}

}  // namespace iife
