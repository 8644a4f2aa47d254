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
  const unsigned long long *iter;  // device iteration counter: reductions 2*iter+1 (delta) and 2*iter+2 (z.r, z.z)
  int *err;
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

// called by ONE thread: wait for all ranks' partials of reduction `seq`, add them in rank order
__device__ __forceinline__ void p2p_wait_sum(const P2PRed &r, unsigned long long seq, double *out, int n) {
  const int par = (int)(seq & 1ull);
  for (int q = 0; q < r.nranks; ++q) spin_until(&r.mbox->it_flag[par][q], seq, r.err);
  __threadfence_system();
  for (int i = 0; i < n; ++i) {
    double s = 0.0;
    for (int q = 0; q < r.nranks; ++q) s += ((volatile double *)r.mbox->it_vals[par][q])[i];
    out[i] = s;
  }
}

}  // namespace iife
