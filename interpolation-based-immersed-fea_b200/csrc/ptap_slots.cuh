// ptap_slots.cuh — hash-free numeric PtAP for ordinary rows, included by ptap.cu.
//
// The pattern never changes between numeric calls (every Newton iteration / time step of the reference
// re-runs AT_R_A on the same sparsity, common.py:432-435), so the symbolic phase records once, for every
// product term of a row, the DENSE INDEX of its destination:
//   slot1[t] = rank of column k in the sorted intermediate row (M^T A_f)[i,:]   (one byte per term)
//   slot2[t] = rank of column l in the sorted output row A_b[i,:]               (one byte per term)
// plus the sorted column list of the intermediate rows (pattern only — its values still never leave
// shared memory).  The numeric kernel then needs no column indices of A_f or M, no hashing, no key
// comparison and no atomics: a step is  load value, load slot byte, LDS / DFMA / STS  into the lane
// group's private copy of a dense accumulator.  The ncu profile of the hashing kernel
// (profiles/r01_ptap_numeric.md) showed ~3 700 warp instructions per output row, mostly probe loops,
// convergence barriers and shuffles; this kernel executes a fraction of that.
//
// Layout per warp in shared memory: h1v[NG1][cap1] then h2v[NG2][cap2] (fp64), NG = 32 / lanes-per-row.
// Rows qualify when both the intermediate and the output row have <= 256 entries (slot bytes);
// everything else stays on the hashing kernels of ptap.cu / ptap_warp.cuh.
#pragma once

namespace iife {

// Tuning knob (compile time): `make EXTRA_NVCCFLAGS=-DIIFE_SLOT_MINBLOCKS=4` lets ptxas use 64 registers for the
// slot-plan kernels (its own choice for <4,2> is 48 registers + 24 bytes of spills); unmeasured so far.
#ifndef IIFE_SLOT_MINBLOCKS
#define IIFE_SLOT_MINBLOCKS 0
#endif
#if IIFE_SLOT_MINBLOCKS > 0
#define IIFE_SLOT_BOUNDS __launch_bounds__(256, IIFE_SLOT_MINBLOCKS)
#else
#define IIFE_SLOT_BOUNDS __launch_bounds__(256)
#endif

constexpr int PS_BATCH = 4;
constexpr int SLOT_TAIL_BYTES = 32;  // one lane number per long item of a 32-item chunk

// items: one per lane (beg/len/w/off in registers); entries e of item `it` add w * x_val[beg+e] into
// hv[group][slots[off+e]].
//
// Operand rows longer than the lane group (M rows of 8 entries with 4 lanes per row: one row in eight of the
// trilinear operator) need a second pass.  CTAIL = false walks all items again and lets the short ones idle;
// CTAIL = true first compacts the lane numbers of the long items into `tail_src` (32 bytes of shared memory per
// warp), so the second pass costs steps only for the items that need it.
template <int LG, bool CTAIL>
__device__ __forceinline__ void slot_stage(int cnt, int my_beg, int my_len, double my_w, int my_off,
                                           const double *__restrict__ x_val, const unsigned char *__restrict__ slots,
                                           double *hv, int cap, int lane, unsigned char *tail_src) {
  constexpr int G = 1 << LG, NG = 32 >> LG;
  const int g = lane >> LG, lg = lane & (G - 1);
  double *hv_g = hv + (size_t)g * cap;
  const int nsteps = (cnt + NG - 1) / NG;
  for (int s0 = 0; s0 < nsteps; s0 += PS_BATCH) {
    int sl[PS_BATCH];
    double v[PS_BATCH];
#pragma unroll
    for (int b = 0; b < PS_BATCH; ++b) {
      int it = (s0 + b) * NG + g;
      int src = it & 31;
      int beg = __shfl_sync(0xffffffffu, my_beg, src);
      int len = __shfl_sync(0xffffffffu, my_len, src);
      int off = __shfl_sync(0xffffffffu, my_off, src);
      double w = __shfl_sync(0xffffffffu, my_w, src);
      bool ok = (it < cnt) && (lg < len);
      sl[b] = -1;
      v[b] = 0.0;
      if (ok) {
        sl[b] = (int)__ldg(slots + off + lg);
        v[b] = w * __ldg(x_val + beg + lg);
      }
    }
#pragma unroll
    for (int b = 0; b < PS_BATCH; ++b) {
      __syncwarp();
      if (sl[b] >= 0) hv_g[sl[b]] += v[b];
    }
  }
  const unsigned long_mask = __ballot_sync(0xffffffffu, my_len > G);  // my_len is 0 beyond cnt
  if (CTAIL && long_mask) {
    const int n_long = __popc(long_mask);
    if (my_len > G) tail_src[__popc(long_mask & ((1u << lane) - 1u))] = (unsigned char)lane;
    __syncwarp();
    for (int t0 = 0; t0 < n_long; t0 += NG) {
      const int idx = t0 + g;
      const int src = idx < n_long ? (int)tail_src[idx] : 0;
      int beg = __shfl_sync(0xffffffffu, my_beg, src);
      int len = __shfl_sync(0xffffffffu, my_len, src);
      int off = __shfl_sync(0xffffffffu, my_off, src);
      double w = __shfl_sync(0xffffffffu, my_w, src);
      if (idx >= n_long) len = 0;
      for (int e = G + lg; __any_sync(0xffffffffu, e < len); e += G) {
        int s1 = -1;
        double v = 0.0;
        if (e < len) {
          s1 = (int)__ldg(slots + off + e);
          v = w * __ldg(x_val + beg + e);
        }
        __syncwarp();
        if (s1 >= 0) hv_g[s1] += v;
      }
    }
  } else if (long_mask) {  // operand rows longer than G: remaining entries
    for (int s = 0; s < nsteps; ++s) {
      int it = s * NG + g;
      int src = it & 31;
      int beg = __shfl_sync(0xffffffffu, my_beg, src);
      int len = __shfl_sync(0xffffffffu, my_len, src);
      int off = __shfl_sync(0xffffffffu, my_off, src);
      double w = __shfl_sync(0xffffffffu, my_w, src);
      if (it >= cnt) len = 0;
      for (int e = G + lg; __any_sync(0xffffffffu, e < len); e += G) {
        int s1 = -1;
        double v = 0.0;
        if (e < len) {
          s1 = (int)__ldg(slots + off + e);
          v = w * __ldg(x_val + beg + e);
        }
        __syncwarp();
        if (s1 >= 0) hv_g[s1] += v;
      }
    }
  }
  __syncwarp();
}

template <int LG1, int LG2, bool CTAIL>
__global__ void IIFE_SLOT_BOUNDS k_ptap_numeric_slots(PtapArgs a, int cap1, int cap2) {
  constexpr int NG1 = 32 >> LG1, NG2 = 32 >> LG2;
  extern __shared__ __align__(16) unsigned char smem[];
  const int lane = threadIdx.x & 31, wic = threadIdx.x >> 5;
  const int wpc = blockDim.x >> 5;
  // per warp: h1v[NG1][cap1], h2v[NG2][cap2] (fp64) and SLOT_TAIL_BYTES for the compacted second pass
  const size_t per_warp = ((size_t)NG1 * cap1 + (size_t)NG2 * cap2) * 8 + SLOT_TAIL_BYTES;
  double *h1v = (double *)(smem + per_warp * wic);
  double *h2v = h1v + (size_t)NG1 * cap1;
  unsigned char *tail_src = (unsigned char *)(h2v + (size_t)NG2 * cap2);
  const int64_t warp_global = (int64_t)blockIdx.x * wpc + wic;
  const int64_t n_warps = (int64_t)gridDim.x * wpc;

  for (int64_t wi = warp_global; wi < a.n_rows; wi += n_warps) {
    const int i = a.rows[wi];
    const int mt_b = __ldg(a.mt_rowptr + i), mt_n = __ldg(a.mt_rowptr + i + 1) - mt_b;
    const int cb = __ldg(a.c_rowptr + i), n2 = __ldg(a.c_rowptr + i + 1) - cb;
    const int ib = __ldg(a.inter_rowptr + i), n1 = __ldg(a.inter_rowptr + i + 1) - ib;
    const unsigned char *s1 = a.slot1 + a.s1_off[i];
    const unsigned char *s2 = a.slot2 + a.s2_off[i];
    // ---- clear the accumulators (only the used prefix of every private copy)
    {
      const int n1r = (n1 + 1) & ~1, n2r = (n2 + 1) & ~1;
      double2 z2 = make_double2(0.0, 0.0);
#pragma unroll
      for (int gg = 0; gg < NG1; ++gg)
        for (int s = lane * 2; s < n1r; s += 64) *(double2 *)(h1v + (size_t)gg * cap1 + s) = z2;
#pragma unroll
      for (int gg = 0; gg < NG2; ++gg)
        for (int s = lane * 2; s < n2r; s += 64) *(double2 *)(h2v + (size_t)gg * cap2 + s) = z2;
    }
    __syncwarp();
    // ---- stage 1: H1[slot] += Mt[i,j] * A[j,e]
    {
      int base_off = 0;
      for (int base = 0; base < mt_n; base += 32) {
        int q = base + lane;
        int my_beg = 0, my_len = 0;
        double my_w = 0.0;
        if (q < mt_n) {
          my_w = __ldg(a.mt_val + mt_b + q);
          if (a.mt_abeg) {  // packed metadata: coalesced, no dependent gather
            my_beg = __ldg(a.mt_abeg + mt_b + q);
            my_len = (int)__ldg(a.mt_alen + mt_b + q);
          } else {
            int j = __ldg(a.mt_col + mt_b + q);
            my_beg = __ldg(a.a_rowptr + j);
            my_len = __ldg(a.a_rowptr + j + 1) - my_beg;
          }
        }
        int total;
        int my_off = warp_excl_scan(my_len, lane, &total);
        slot_stage<LG1, CTAIL>(min(32, mt_n - base), my_beg, my_len, my_w, my_off, a.a_val, s1 + base_off, h1v, cap1,
                               lane, tail_src);
        base_off += total;
      }
    }
    // ---- merge the private copies of the intermediate row (fixed order)
    for (int q = lane; q < n1; q += 32) {
      double v = h1v[q];
#pragma unroll
      for (int gg = 1; gg < NG1; ++gg) v += h1v[(size_t)gg * cap1 + q];
      h1v[q] = v;
    }
    __syncwarp();
    // ---- stage 2: H2[slot] += H1[q] * M[k_q, e]
    {
      int base_off = 0;
      for (int base = 0; base < n1; base += 32) {
        int q = base + lane;
        int my_beg = 0, my_len = 0;
        double my_w = 0.0;
        if (q < n1) {
          my_w = h1v[q];
          if (a.inter_mbeg) {
            my_beg = __ldg(a.inter_mbeg + ib + q);
            my_len = (int)__ldg(a.inter_mlen + ib + q);
          } else {
            int k = __ldg(a.inter_col + ib + q);
            my_beg = __ldg(a.m_rowptr + k);
            my_len = __ldg(a.m_rowptr + k + 1) - my_beg;
          }
        }
        int total;
        int my_off = warp_excl_scan(my_len, lane, &total);
        slot_stage<LG2, CTAIL>(min(32, n1 - base), my_beg, my_len, my_w, my_off, a.m_val, s2 + base_off, h2v, cap2,
                               lane, tail_src);
        base_off += total;
      }
    }
    // ---- write the row (slot = position in the sorted output row)
    for (int s = lane; s < n2; s += 32) {
      double v = h2v[s];
#pragma unroll
      for (int gg = 1; gg < NG2; ++gg) v += h2v[(size_t)gg * cap2 + s];
      a.c_val[cb + s] = v;
    }
    __syncwarp();
  }
}

typedef void (*slot_kernel_t)(PtapArgs, int, int);
static slot_kernel_t pick_slot_kernel(int lg1, int lg2, bool ctail) {
#define PSK(a_, b_) \
  if (lg1 == a_ && lg2 == b_) return ctail ? k_ptap_numeric_slots<a_, b_, true> : k_ptap_numeric_slots<a_, b_, false>;
  PSK(3, 2) PSK(3, 3) PSK(3, 4) PSK(3, 5)
  PSK(4, 2) PSK(4, 3) PSK(4, 4) PSK(4, 5)
  PSK(5, 2) PSK(5, 3) PSK(5, 4) PSK(5, 5)
#undef PSK
  return nullptr;
}

}  // namespace iife
