// ptap_warp.cuh — the numeric PtAP kernel for ordinary rows (one warp per output row), included by
// ptap.cu.  Same two-stage traversal as the generic kernel there, restructured around what the ncu
// profile of the first version showed (profiles/r01_ptap_numeric_v0.md): the warp sat on the long
// scoreboard behind one dependent global load per step, and the fp64 shared-memory atomics compile to
// CAS spin loops (ATOMS.CAST.SPIN) that serialise conflicting lanes.
//
//  * value tables are PRIVATISED per lane group: the 32/G lanes-groups that walk different operand
//    rows at the same time each add into their own copy of the value table, so a step is a plain
//    LDS / DFMA / STS (entries of one CSR row have distinct columns => distinct slots inside a
//    group); the copies are merged in a fixed order when the table is compacted / written out.
//    Only the KEY table is shared and uses the native 32-bit ATOMS.CAS.  No fp64 atomics, and the
//    summation order is fixed => bit-reproducible.
//  * operand entries are fetched in batches of BATCH steps before any hashing, so BATCH independent
//    (colind, val) loads per lane are in flight instead of one.
//  * per-item metadata (row start, length, weight) lives in registers and is broadcast with shuffles.
#pragma once

namespace iife {

constexpr int PW_BATCH = 4;

__device__ __forceinline__ int sh_insert(int *hk, unsigned mask, int shift, int key) {
  unsigned h = ((unsigned)key * HASH_MUL) >> shift;
  for (unsigned probes = 0; probes <= mask; ++probes) {
    int old = hk[h];
    if (old == key) return (int)h;
    if (old == EMPTY) {
      old = atomicCAS(&hk[h], EMPTY, key);
      if (old == EMPTY || old == key) return (int)h;
    }
    h = (h + 1) & mask;
  }
  return -1;
}

__device__ __forceinline__ int sh_find(const int *hk, unsigned mask, int shift, int key) {
  unsigned h = ((unsigned)key * HASH_MUL) >> shift;
  for (unsigned probes = 0; probes <= mask; ++probes) {
    int cur = hk[h];
    if (cur == key) return (int)h;
    if (cur == EMPTY) return -1;
    h = (h + 1) & mask;
  }
  return -1;
}

// One stage: for the `cnt` items held one per lane (my_beg/my_len/my_w), accumulate
//   hv[group][slot(col)] += w * X.val   over all entries of the items' CSR rows of X.
// INSERT: create keys (stage 1) or look them up (stage 2).
template <int LG, bool INSERT>
__device__ __forceinline__ void warp_stage(int cnt, int my_beg, int my_len, double my_w,
                                           const int *__restrict__ x_col, const double *__restrict__ x_val, int *hk,
                                           double *hv, int cap, unsigned mask, int shift, int lane, int *fail) {
  constexpr int G = 1 << LG, NG = 32 >> LG;
  const int g = lane >> LG, lg = lane & (G - 1);
  double *hv_g = hv + (size_t)g * cap;
  const int nsteps = (cnt + NG - 1) / NG;
  for (int s0 = 0; s0 < nsteps; s0 += PW_BATCH) {
    int c[PW_BATCH];
    double v[PW_BATCH];
#pragma unroll
    for (int b = 0; b < PW_BATCH; ++b) {
      int it = (s0 + b) * NG + g;
      int src = it & 31;
      int beg = __shfl_sync(0xffffffffu, my_beg, src);
      int len = __shfl_sync(0xffffffffu, my_len, src);
      double w = __shfl_sync(0xffffffffu, my_w, src);
      bool ok = (it < cnt) && (lg < len);
      c[b] = EMPTY;
      v[b] = 0.0;
      if (ok) {
        c[b] = __ldg(x_col + beg + lg);
        v[b] = w * __ldg(x_val + beg + lg);
      }
    }
#pragma unroll
    for (int b = 0; b < PW_BATCH; ++b) {
      int slot = -1;
      if (c[b] != EMPTY) {
        slot = INSERT ? sh_insert(hk, mask, shift, c[b]) : sh_find(hk, mask, shift, c[b]);
        if (slot < 0) *fail = 1;
      }
      __syncwarp();
      if (slot >= 0) hv_g[slot] += v[b];
    }
  }
  // rows longer than G (uncommon): remaining entries, one step at a time
  if (__any_sync(0xffffffffu, my_len > G)) {
    for (int s = 0; s < nsteps; ++s) {
      int it = s * NG + g;
      int src = it & 31;
      int beg = __shfl_sync(0xffffffffu, my_beg, src);
      int len = __shfl_sync(0xffffffffu, my_len, src);
      double w = __shfl_sync(0xffffffffu, my_w, src);
      if (it >= cnt) len = 0;
      for (int e = G + lg; __any_sync(0xffffffffu, e < len); e += G) {
        int slot = -1;
        double v = 0.0;
        if (e < len) {
          int col = __ldg(x_col + beg + e);
          v = w * __ldg(x_val + beg + e);
          slot = INSERT ? sh_insert(hk, mask, shift, col) : sh_find(hk, mask, shift, col);
          if (slot < 0) *fail = 1;
        }
        __syncwarp();
        if (slot >= 0) hv_g[slot] += v;
      }
    }
  }
  __syncwarp();
}

template <int LG1, int LG2>
__global__ void __launch_bounds__(256) k_ptap_numeric_warp(PtapArgs a) {
  constexpr int NG1 = 32 >> LG1, NG2 = 32 >> LG2;
  extern __shared__ __align__(16) unsigned char smem[];
  const int lane = threadIdx.x & 31, wic = threadIdx.x >> 5;
  const int wpc = blockDim.x >> 5;
  const int cap1 = 1 << a.log_cap1, cap2 = 1 << a.log_cap2;
  const unsigned mask1 = cap1 - 1, mask2 = cap2 - 1;
  const int shift1 = 32 - a.log_cap1, shift2 = 32 - a.log_cap2;
  const size_t per_warp = ((size_t)NG1 * cap1 + (size_t)NG2 * cap2) * 8 + ((size_t)cap1 + cap2) * 4;
  unsigned char *wbase = smem + per_warp * wic;
  double *h1v = (double *)wbase;
  double *h2v = h1v + (size_t)NG1 * cap1;
  int *h1k = (int *)(h2v + (size_t)NG2 * cap2);
  int *h2k = h1k + cap1;
  const int64_t warp_global = (int64_t)blockIdx.x * wpc + wic;
  const int64_t n_warps = (int64_t)gridDim.x * wpc;
  int fail = 0;

  for (int64_t wi = warp_global; wi < a.n_rows; wi += n_warps) {
    const int i = a.rows[wi];
    const int mt_b = __ldg(a.mt_rowptr + i), mt_n = __ldg(a.mt_rowptr + i + 1) - mt_b;
    const int cb = __ldg(a.c_rowptr + i), n2 = __ldg(a.c_rowptr + i + 1) - cb;
    // ---- clear the tables (values with 16-byte stores)
    {
      double2 z2 = make_double2(0.0, 0.0);
      double2 *v2 = (double2 *)h1v;
      const int nv2 = (NG1 * cap1 + NG2 * cap2) >> 1;
      for (int s = lane; s < nv2; s += 32) v2[s] = z2;
      int4 e4 = make_int4(EMPTY, EMPTY, EMPTY, EMPTY);
      int4 *k4 = (int4 *)h1k;
      const int nk4 = (cap1 + cap2) >> 2;
      for (int s = lane; s < nk4; s += 32) k4[s] = e4;
    }
    __syncwarp();
    // ---- output keys are known from the symbolic phase
    for (int s = lane; s < n2; s += 32)
      if (sh_insert(h2k, mask2, shift2, __ldg(a.c_col + cb + s)) < 0) fail = 1;
    // ---- stage 1: H1 = sum_j Mt[i,j] * A[j,:]
    for (int base = 0; base < mt_n; base += 32) {
      int q = base + lane;
      int my_beg = 0, my_len = 0;
      double my_w = 0.0;
      if (q < mt_n) {
        int j = __ldg(a.mt_col + mt_b + q);
        my_w = __ldg(a.mt_val + mt_b + q);
        my_beg = __ldg(a.a_rowptr + j);
        my_len = __ldg(a.a_rowptr + j + 1) - my_beg;
      }
      warp_stage<LG1, true>(min(32, mt_n - base), my_beg, my_len, my_w, a.a_col, a.a_val, h1k, h1v, cap1, mask1, shift1,
                            lane, &fail);
    }
    // ---- compact H1 in place, merging the NG1 private copies in a fixed order
    int n1 = 0;
    for (int sb = 0; sb < cap1; sb += 32) {
      int k = h1k[sb + lane];
      double v = h1v[sb + lane];
#pragma unroll
      for (int gg = 1; gg < NG1; ++gg) v += h1v[(size_t)gg * cap1 + sb + lane];
      unsigned m = __ballot_sync(0xffffffffu, k != EMPTY);
      __syncwarp();
      if (k != EMPTY) {
        int pos = n1 + __popc(m & ((1u << lane) - 1u));
        h1k[pos] = k;
        h1v[pos] = v;
      }
      n1 += __popc(m);
      __syncwarp();
    }
    // ---- stage 2: H2 = sum_k H1[k] * M[k,:]   (keys pre-filled: lookups only)
    {
      int nk = 0, nbeg = 0, nlen = 0;
      double nv = 0.0;
      if (lane < n1) {
        nk = h1k[lane];
        nv = h1v[lane];
        nbeg = __ldg(a.m_rowptr + nk);
        nlen = __ldg(a.m_rowptr + nk + 1) - nbeg;
      }
      for (int base = 0; base < n1; base += 32) {
        int my_beg = nbeg, my_len = nlen;
        double my_w = nv;
        // prefetch the next chunk's row pointers while this chunk is processed
        int qn = base + 32 + lane;
        nlen = 0;
        if (qn < n1) {
          nk = h1k[qn];
          nv = h1v[qn];
          nbeg = __ldg(a.m_rowptr + nk);
          nlen = __ldg(a.m_rowptr + nk + 1) - nbeg;
        }
        warp_stage<LG2, false>(min(32, n1 - base), my_beg, my_len, my_w, a.m_col, a.m_val, h2k, h2v, cap2, mask2, shift2,
                               lane, &fail);
      }
    }
    // ---- write the row: columns ascending as stored by the symbolic phase
    for (int s = lane; s < n2; s += 32) {
      int slot = sh_find(h2k, mask2, shift2, __ldg(a.c_col + cb + s));
      double v = 0.0;
      if (slot >= 0) {
#pragma unroll
        for (int gg = 0; gg < NG2; ++gg) v += h2v[(size_t)gg * cap2 + slot];
      } else {
        fail = 1;
      }
      a.c_val[cb + s] = v;
    }
    __syncwarp();
  }
  if (fail) atomicExch(a.err_flag, 1);
}

static size_t warp_kernel_smem_per_warp(int lg1, int lg2, int log_cap1, int log_cap2) {
  size_t ng1 = 32 >> lg1, ng2 = 32 >> lg2, c1 = (size_t)1 << log_cap1, c2 = (size_t)1 << log_cap2;
  return (ng1 * c1 + ng2 * c2) * 8 + (c1 + c2) * 4;
}

typedef void (*warp_kernel_t)(PtapArgs);
static warp_kernel_t pick_warp_kernel(int lg1, int lg2) {
#define PWK(a_, b_) \
  if (lg1 == a_ && lg2 == b_) return k_ptap_numeric_warp<a_, b_>;
  PWK(3, 2) PWK(3, 3) PWK(3, 4) PWK(3, 5)
  PWK(4, 2) PWK(4, 3) PWK(4, 4) PWK(4, 5)
  PWK(5, 2) PWK(5, 3) PWK(5, 4) PWK(5, 5)
#undef PWK
  return nullptr;
}

}  // namespace iife
