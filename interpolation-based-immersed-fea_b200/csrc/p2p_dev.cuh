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
};

// SpMV prologue of the three-kernel CG iteration: wait until the ghost entries of exchange *seq_base + k_off + 1 arrived
struct HaloWait {
  const unsigned long long *flags;     // my mailbox's halo_flag[] (nullptr: no wait)
  const unsigned long long *seq_base;  // device halo sequence counter (moves once per chunk)
  int k_off, nranks;
  unsigned int recv_mask;
  int *err;
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
    spin_until(&r.mbox->it_flag[par][lane], seq, r.err);
    for (int i = 0; i < n; ++i) v[i] = ((volatile double *)r.mbox->it_vals[par][lane])[i];
  }
  for (int i = 0; i < n; ++i) {
    double s = 0.0;
    for (int q = 0; q < r.nranks; ++q) s += __shfl_sync(0xffffffffu, v[i], q);
    out[i] = s;
  }
}

}  // namespace iife
