// ksp_persist.cuh — EXPERIMENTAL persistent cooperative CG kernel, included by ksp.cu; selected with
// IIFE_KSP_PERSIST=1 and OFF by default: written at the end of round 1 from the timeline in ROUND_NOTES.md
// (8 GPUs: 109 us per iteration against 51 us of HBM time — the rest is six kernel boundaries, two mailbox
// reductions and the halo), compiled and inspected but not yet run on a GPU.
//
// One cooperative launch runs up to n_iters CG iterations.  The arithmetic and the scalar steps are those of
// k_cg_p / k_spmv_sell / k_vec_dot / k_cg_update (same PETSc semantics); what changes is the plumbing:
//   phase A   p = z + (beta/beta_old) p                      (thread-private ranges: no barrier before it)
//   barrier 1
//   halo      boundary entries of p stored into the neighbours' ghost sections, last CTA raises their flags
//   phase B   w = A p on the SELL copy, interior slices first, then wait for the neighbours' flags, then the
//             slices that read ghost entries (those through L2); per-CTA partial of (p, w)
//   barrier 2
//   CTA 0 adds the partials in block order, exchanges the sum with the other ranks through the mailboxes
//   (rank-order sum: bit-identical on every rank) and publishes delta to its own grid through a flag
//   phase C   x += alpha p, r -= alpha w, partials of (z, r), (z, z)
//   barrier 3
//   CTA 0 reduces/exchanges/publishes; EVERY thread then performs the scalar step (beta, dp, convergence test)
//   on a private copy of the scalars, so the whole grid takes the same branch without another barrier.
// Only CTA 0 ever waits for another GPU (bounded spin; on a timeout it publishes NaN, which ends the loop
// uniformly with DIVERGED_NANORINF instead of desynchronising the grid).  Every CTA that waits on a flag is
// co-resident by construction (cooperative launch); no two kernels ever spin on each other.
#pragma once
#include <cooperative_groups.h>

namespace iife {

namespace cg = cooperative_groups;

struct GridBcast {  // lives in device memory, zeroed once per solve
  unsigned long long flag;
  double vals[2][2];
};

struct CgPersist {
  // SELL-32 operator (spmv.cu)
  const int *sell_ptr, *sell_cptr, *sell_col;
  const double *sell_val;
  long long n_rows, n_slices;
  const int *order;  // interior slices first (nullptr: natural order, everything interior)
  long long n_interior;
  // vectors: p has n_rows + n_ghost entries
  double *x, *r, *p, *w;
  const double *dinv;
  double *sc;
  int *fl;
  double *hist;
  long long hist_len;
  double *partials;  // [3 * MAX_PARTIALS]
  GridBcast *bc;
  int n_iters;
  // row-partitioned solver over peer memory (p2p.cu); dist == 0 on one GPU
  int dist;
  P2PRed pr;
  PeerTable pt;
  Mailbox *mbox;
  const int *send_idx;
  const unsigned char *send_peer;
  const int *send_off;
  long long total_send;
  unsigned int send_mask, recv_mask;
  unsigned long long *halo_seq, *iter_ptr;
  unsigned int *push_counter;
};

__device__ __forceinline__ void st_flag_gpu(unsigned long long *p, unsigned long long v) {
  asm volatile("st.release.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_flag_gpu(const unsigned long long *p) {
  unsigned long long v;
  asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}

// Polling with acquire loads makes the SM drop its L1 on every probe (CCTL.IVALL in the SASS), which also hits the
// x gathers of CTAs still multiplying on the same SM: poll relaxed, fence once when the flag has arrived.
__device__ __forceinline__ bool wait_flag_sys(const unsigned long long *p, unsigned long long seq, int *err) {
  unsigned long long v;
  const long long t0 = clock64();
  for (;;) {
    asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    if (v >= seq) break;
    if (clock64() - t0 > P2P_SPIN_LIMIT) {
      atomicExch(err, 1);
      return false;
    }
  }
  asm volatile("fence.acq_rel.sys;" ::: "memory");
  return true;
}
__device__ __forceinline__ void wait_flag_gpu(const unsigned long long *p, unsigned long long seq) {
  unsigned long long v;
  do {
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  } while (v < seq);
  asm volatile("fence.acq_rel.gpu;" ::: "memory");
}

// one slice of the SELL operator; x is written inside the same kernel, so it is read with ordinary cached
// loads (valid after a grid barrier) or, for slices that touch ghost entries, through L2 (VIA_L2)
template <int U, bool VIA_L2>
__device__ __forceinline__ double sell_slice_live(const int *__restrict__ sell_ptr, const int *__restrict__ sell_cptr,
                                                  const int *__restrict__ sell_col, const double *__restrict__ sell_val,
                                                  const double *x, long long s, int lane, long long n_rows) {
  const int sb = __ldg(sell_ptr + s), se = __ldg(sell_ptr + s + 1);
  const int cb = __ldg(sell_cptr + s), ce = __ldg(sell_cptr + s + 1);
  const double *vp = sell_val + sb + lane;
  const int width = (se - sb) >> 5;
  const bool uniform = (ce - cb) == width;
  double a[U];
#pragma unroll
  for (int u = 0; u < U; ++u) a[u] = 0.0;
  int k = 0;
  if (uniform) {
    const long long i = s * 32 + lane;
    const int row = (i < n_rows) ? (int)i : 0;
    const int live = (i < n_rows) ? 1 : 0;
    const int *op = sell_col + cb;
    for (; k + U <= width; k += U) {
      int c[U];
      double v[U];
#pragma unroll
      for (int u = 0; u < U; ++u) c[u] = row + live * __ldg(op + k + u);
#pragma unroll
      for (int u = 0; u < U; ++u) v[u] = __ldcs(vp + (k + u) * 32);
#pragma unroll
      for (int u = 0; u < U; ++u) a[u] = fma(v[u], VIA_L2 ? __ldcg(x + c[u]) : x[c[u]], a[u]);
    }
    for (; k < width; ++k) {
      const int c1 = row + live * __ldg(op + k);
      a[0] = fma(__ldcs(vp + k * 32), VIA_L2 ? __ldcg(x + c1) : x[c1], a[0]);
    }
  } else {
    const int *cp = sell_col + cb + lane;
    for (; k + U <= width; k += U) {
      int c[U];
      double v[U];
#pragma unroll
      for (int u = 0; u < U; ++u) c[u] = __ldcs(cp + (k + u) * 32);
#pragma unroll
      for (int u = 0; u < U; ++u) v[u] = __ldcs(vp + (k + u) * 32);
#pragma unroll
      for (int u = 0; u < U; ++u) a[u] = fma(v[u], VIA_L2 ? __ldcg(x + c[u]) : x[c[u]], a[u]);
    }
    for (; k < width; ++k) {
      const int c1 = __ldcs(cp + k * 32);
      a[0] = fma(__ldcs(vp + k * 32), VIA_L2 ? __ldcg(x + c1) : x[c1], a[0]);
    }
  }
  double acc = a[0];
#pragma unroll
  for (int u = 1; u < U; ++u) acc += a[u];
  return acc;
}

// CTA 0: add `nv` rows of per-CTA partials in block order, exchange with the other ranks, publish to the grid.
// Other CTAs: wait for the publication.  Returns the nv sums in out[] in every thread of every CTA.
__device__ __forceinline__ void persist_reduce(const CgPersist &a, const double *partials, int nv, unsigned long long bseq,
                                               unsigned long long pseq, double *red, double *s_val, double *out) {
  const int tid = threadIdx.x;
  const int par = (int)(bseq & 1ull);
  if (blockIdx.x == 0) {
    for (int v = 0; v < nv; ++v) {
      double s = 0.0;
      for (int k = tid; k < (int)gridDim.x; k += blockDim.x) s += __ldcg(partials + (size_t)v * MAX_PARTIALS + k);
      s = block_sum(s, red);
      if (tid == 0) s_val[v] = s;
    }
    __syncthreads();
    if (a.dist) {
      p2p_push(a.pr, pseq, s_val, nv, tid);  // threads 0..nranks-1
      __syncthreads();
      if (tid < 32) {
        double t[4];
        p2p_wait_sum(a.pr, pseq, t, nv);  // bounded spin; sets *pr.err on a timeout
        if (tid == 0) {
          const bool bad = (*(volatile int *)a.pr.err) != 0;
          for (int v = 0; v < nv; ++v) s_val[v] = bad ? __longlong_as_double(0x7ff8000000000000LL) : t[v];
        }
      }
      __syncthreads();
    }
    if (tid == 0) {
      for (int v = 0; v < nv; ++v) a.bc->vals[par][v] = s_val[v];
      __threadfence();
      st_flag_gpu(&a.bc->flag, bseq);
    }
  } else {
    if (tid == 0) {
      wait_flag_gpu(&a.bc->flag, bseq);
      for (int v = 0; v < nv; ++v) s_val[v] = ((volatile double *)a.bc->vals[par])[v];
    }
    __syncthreads();
  }
  for (int v = 0; v < nv; ++v) out[v] = s_val[v];
  __syncthreads();  // s_val is reused by the next reduction
}

// minBlocks = 4 (<= 64 registers): with ptxas' own choice (48 registers, 36 bytes of spills) the slice loop keeps
// one load in flight instead of U (checked in the SASS; the same thing happened to the dot-fused SELL kernel)
template <int U>
__global__ void __launch_bounds__(VEC_THREADS, 4) k_cg_persist(CgPersist a) {
  cg::grid_group grid = cg::this_grid();
  __shared__ double red[32];
  __shared__ double s_val[4];
  __shared__ bool s_last;
  const int tid = threadIdx.x, lane = tid & 31;
  const long long gtid = (long long)blockIdx.x * blockDim.x + tid;
  const long long gstride = (long long)gridDim.x * blockDim.x;
  const long long w0 = gtid >> 5, nw = gstride >> 5;
  const long long n = a.n_rows;
  const bool lead = (blockIdx.x == 0 && tid == 0);
  // private copy of the scalar state: identical in every thread of the grid at every step
  double lsc[S_COUNT];
  int lfl[F_COUNT];
#pragma unroll
  for (int k = 0; k < S_COUNT; ++k) lsc[k] = a.sc[k];
#pragma unroll
  for (int k = 0; k < F_COUNT; ++k) lfl[k] = a.fl[k];
  unsigned long long iter = a.dist ? *a.iter_ptr : 0ull;
  unsigned long long hseq = a.dist ? *a.halo_seq : 0ull;
  unsigned long long bseq = *(volatile unsigned long long *)&a.bc->flag;  // last publication of the previous launch

  for (int it = 0; it < a.n_iters && lfl[F_REASON] == 0; ++it) {
    // ---- phase A: p = z + (beta/beta_old) p, z = D^-1 r
    {
      const bool first = (lfl[F_ITS] == 0);
      const double bb = first ? 0.0 : lsc[S_BETA] / lsc[S_BETA_OLD];
      for (long long i = gtid; i < n; i += gstride) {
        const double z = (a.dinv ? __ldg(a.dinv + i) : 1.0) * a.r[i];
        a.p[i] = first ? z : fma(bb, a.p[i], z);
      }
    }
    grid.sync();
    // ---- halo: my boundary entries into the neighbours' ghost sections
    if (a.dist) {
      for (long long k = gtid; k < a.total_send; k += gstride) {
        const int q = a.send_peer[k];
        a.pt.xbuf[q][a.pt.dst_start[q] + (k - a.send_off[q])] = a.p[a.send_idx[k]];
      }
      __threadfence_system();
      __syncthreads();
      if (tid == 0) s_last = (atomicAdd(a.push_counter, 1u) == gridDim.x - 1);
      __syncthreads();
      hseq += 1ull;
      if (s_last) {
        __threadfence_system();
        if (tid < a.pr.nranks && ((a.send_mask >> tid) & 1u)) st_flag(&a.pt.mbox[tid]->halo_flag[a.pr.me], hseq);
        if (tid == 0) *a.push_counter = 0u;  // next use is behind the next barrier 1
      }
    }
    // ---- phase B: w = A p, partial (p, w)
    double dsum = 0.0;
    for (long long idx = w0; idx < a.n_interior; idx += nw) {
      const long long s = a.order ? (long long)a.order[idx] : idx;
      const double acc = sell_slice_live<U, false>(a.sell_ptr, a.sell_cptr, a.sell_col, a.sell_val, a.p, s, lane, n);
      const long long i = s * 32 + lane;
      if (i < n) {
        a.w[i] = acc;
        dsum = fma(acc, a.p[i], dsum);
      }
    }
    if (a.dist) {
      if (tid < a.pr.nranks && ((a.recv_mask >> tid) & 1u)) wait_flag_sys(&a.mbox->halo_flag[tid], hseq, a.pr.err);
      __syncthreads();
      for (long long idx = a.n_interior + w0; idx < a.n_slices; idx += nw) {
        const long long s = a.order[idx];
        const double acc = sell_slice_live<U, true>(a.sell_ptr, a.sell_cptr, a.sell_col, a.sell_val, a.p, s, lane, n);
        const long long i = s * 32 + lane;
        if (i < n) {
          a.w[i] = acc;
          dsum = fma(acc, a.p[i], dsum);
        }
      }
    }
    {
      const double bs = block_sum(dsum, red);
      if (tid == 0) a.partials[blockIdx.x] = bs;
    }
    grid.sync();
    double delta;
    bseq += 1ull;
    persist_reduce(a, a.partials, 1, bseq, 2ull * iter + 1ull, red, s_val, &delta);
    if (lead) a.sc[S_DELTA] = delta;
    lsc[S_DELTA] = delta;
    if (!(delta > 0.0)) {
      // (p, A p) <= 0 or NaN: no update (PETSc: its = i+1 at that point)
      lfl[F_ITS] += 1;
      lfl[F_REASON] = isnan(delta) ? IIFE_KSP_DIVERGED_NANORINF : IIFE_KSP_DIVERGED_INDEFINITE_MAT;
      if (lead) {
        a.fl[F_ITS] = lfl[F_ITS];
        __threadfence();
        a.fl[F_REASON] = lfl[F_REASON];
      }
      break;
    }
    // ---- phase C: x += alpha p, r -= alpha w, partials of (z, r), (z, z)
    {
      const double alpha = lsc[S_BETA] / delta;
      double acc0 = 0.0, acc1 = 0.0;
      for (long long i = gtid; i < n; i += gstride) {
        const double pi = a.p[i], wi = a.w[i];
        a.x[i] = fma(alpha, pi, a.x[i]);
        const double ri = fma(-alpha, wi, a.r[i]);
        a.r[i] = ri;
        const double z = (a.dinv ? __ldg(a.dinv + i) : 1.0) * ri;
        acc0 = fma(z, ri, acc0);
        acc1 = fma(z, z, acc1);
      }
      const double b0 = block_sum(acc0, red);
      if (tid == 0) a.partials[MAX_PARTIALS + blockIdx.x] = b0;
      const double b1 = block_sum(acc1, red);
      if (tid == 0) a.partials[2 * MAX_PARTIALS + blockIdx.x] = b1;
    }
    grid.sync();
    double zrzz[2];
    bseq += 1ull;
    persist_reduce(a, a.partials + MAX_PARTIALS, 2, bseq, 2ull * iter + 2ull, red, s_val, zrzz);
    cg_update_scalars(lsc, lfl, zrzz[0], zrzz[1], nullptr, 0);  // every thread, on its private copy
    if (lead) cg_update_scalars(a.sc, a.fl, zrzz[0], zrzz[1], a.hist, a.hist_len);
    iter += 1ull;
  }
  if (lead && a.dist) {
    *a.iter_ptr = iter;
    *a.halo_seq = hseq;
  }
}

}  // namespace iife
