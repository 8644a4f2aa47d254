// ptap_tpl.cuh — TEMPLATE numeric PtAP (device side), included by ptap.cu.  See ptap_tpl_host.h for the idea and
// the program format.
//
// Build (once per plan and per version of the values of M / R; tpl_ensure in ptap.cu):
//   k_tpl_hash      one warp per slot-plan row: two independent 64-bit hashes of everything the numeric result of the
//                   row depends on EXCEPT the values and positions of its A_f rows: operand row lengths, the
//                   destination byte of every product term (slot plan), the values of R[i,:] and of the M rows
//   (radix sort of the first hash, cub)       rows with equal structure become neighbours
//   k_tpl_heads / k_tpl_run_starts / k_tpl_select   runs of equal hashes with at least `min_rows` members
//   k_tpl_extract   the raw description of one representative row per run -> host, compiled there (ptap_tpl_host.h)
//   k_tpl_assign    members whose SECOND hash equals the representative's join the template (a 128-bit match);
//                   everything else stays on the per-row kernels
// Numeric (every call):
//   k_ptap_numeric_tpl   one warp per output row, chunks of rows of one template per warp; per row it reads the
//                   operand starts (4 B per operand row), the A_f values, and writes the output values: no slot
//                   bytes, no column indices, no M / R values from HBM.  The program is shared by thousands of
//                   rows and stays in L1/L2.
#pragma once
#include <cub/device/device_radix_sort.cuh>
#include "ptap_tpl_host.h"

namespace iife {

constexpr int TPL_CHUNK = 16;  // rows of one template handed to a warp at a time

// fixed-stride raw record written by k_tpl_extract (parsed by tpl_parse_raw in ptap.cu)
constexpr int TPLR_INTS = 0;                                     // n0, n1, n2, T1, T2, ok, pad, pad
constexpr int TPLR_W = 32;                                       // f64 [MAX_N0]
constexpr int TPLR_MVAL = TPLR_W + 8 * tpl::MAX_N0;              // f64 [MAX_T2 + 1]
constexpr int TPLR_LEN1 = TPLR_MVAL + 8 * (tpl::MAX_T2 + 1);     // u8  [MAX_N0]
constexpr int TPLR_SLOT1 = TPLR_LEN1 + tpl::MAX_N0;              // u8  [MAX_T1 + 1]
constexpr int TPLR_LEN2 = TPLR_SLOT1 + tpl::MAX_T1 + 1;          // u8  [MAX_N1]
constexpr int TPLR_SLOT2 = TPLR_LEN2 + tpl::MAX_N1;              // u8  [MAX_T2 + 1]
constexpr int TPLR_STRIDE = TPLR_SLOT2 + tpl::MAX_T2 + 1;
static_assert(TPLR_STRIDE % 8 == 0, "raw record stride keeps the f64 sections aligned");

__device__ __forceinline__ unsigned long long tpl_mix(unsigned long long x) {
  x ^= x >> 30;
  x *= 0xbf58476d1ce4e5b9ull;
  x ^= x >> 27;
  x *= 0x94d049bb133111ebull;
  x ^= x >> 31;
  return x;
}
// order-dependent through `pos`, summed over items => lanes can hash their items independently
__device__ __forceinline__ void tpl_item(unsigned long long &ha, unsigned long long &hb, unsigned tag, unsigned pos,
                                         unsigned long long v) {
  const unsigned long long k = tpl_mix(((unsigned long long)tag << 40) ^ (unsigned long long)pos ^ 0x51ed270b7a3c9f15ull);
  ha += tpl_mix(v ^ k);
  hb += tpl_mix((v + 0x9e3779b97f4a7c15ull) * 0xd6e8feb86659fd93ull ^ (k >> 1) ^ (k << 63));
}

// rows[0..n): rows of the slot-plan bins.  Ineligible rows get a key nobody shares.
__global__ void __launch_bounds__(256) k_tpl_hash(PtapArgs a, unsigned long long *__restrict__ h1,
                                                  unsigned long long *__restrict__ h2) {
  const int lane = threadIdx.x & 31;
  const long long wpc = blockDim.x >> 5;
  const long long n_warps = (long long)gridDim.x * wpc;
  for (long long wi = (long long)blockIdx.x * wpc + (threadIdx.x >> 5); wi < a.n_rows; wi += n_warps) {
    const int i = a.rows[wi];
    const int mtb = a.mt_rowptr[i], n0 = a.mt_rowptr[i + 1] - mtb;
    const int ib = a.inter_rowptr[i], n1 = a.inter_rowptr[i + 1] - ib;
    const int n2 = a.c_rowptr[i + 1] - a.c_rowptr[i];
    const unsigned char *s1 = a.slot1 + a.s1_off[i];
    const unsigned char *s2 = a.slot2 + a.s2_off[i];
    unsigned long long ha = 0, hb = 0;
    if (lane == 0) {
      tpl_item(ha, hb, 0, 0, (unsigned long long)n0);
      tpl_item(ha, hb, 0, 1, (unsigned long long)n1);
      tpl_item(ha, hb, 0, 2, (unsigned long long)n2);
    }
    int T1 = 0, T2 = 0;
    for (int q = lane; q < n0; q += 32) {
      const int len = (int)a.mt_alen[mtb + q];
      T1 += len;
      tpl_item(ha, hb, 1, (unsigned)q, (unsigned long long)len);
      tpl_item(ha, hb, 2, (unsigned)q, (unsigned long long)__double_as_longlong(a.mt_val[mtb + q]));
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) T1 += __shfl_xor_sync(0xffffffffu, T1, o);
    for (int t = lane; t < T1; t += 32) tpl_item(ha, hb, 3, (unsigned)t, (unsigned long long)s1[t]);
    int base_off = 0;
    for (int base = 0; base < n1; base += 32) {
      const int q = base + lane;
      int len = 0, beg = 0;
      if (q < n1) {
        len = (int)a.inter_mlen[ib + q];
        beg = a.inter_mbeg[ib + q];
        tpl_item(ha, hb, 4, (unsigned)q, (unsigned long long)len);
      }
      int total;
      const int off = warp_excl_scan(len, lane, &total);
      for (int e = 0; e < len; ++e) {
        const unsigned t = (unsigned)(base_off + off + e);
        tpl_item(ha, hb, 5, t, (unsigned long long)__double_as_longlong(a.m_val[beg + e]));
        tpl_item(ha, hb, 6, t, (unsigned long long)s2[t]);
      }
      base_off += total;
    }
    T2 = base_off;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      ha += __shfl_xor_sync(0xffffffffu, ha, o);
      hb += __shfl_xor_sync(0xffffffffu, hb, o);
    }
    const bool ok = n0 >= 1 && n0 <= tpl::MAX_N0 && T1 >= 1 && T1 <= tpl::MAX_T1 && n1 >= 1 && n1 <= tpl::MAX_N1 && n2 >= 1 &&
                    n2 <= tpl::MAX_N2 && T2 >= 1 && T2 <= tpl::MAX_T2;
    if (lane == 0) {
      // eligible keys have the top bit clear; ineligible rows get (1 << 63) | position: a run of one
      h1[wi] = ok ? (ha >> 1) : ((1ull << 63) | (unsigned long long)wi);
      h2[wi] = hb;
    }
  }
}

__global__ void k_tpl_iota(int *__restrict__ v, long long n) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (; i < n; i += stride) v[i] = (int)i;
}

__global__ void k_tpl_heads(const unsigned long long *__restrict__ key, long long n, int *__restrict__ head) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (; i < n; i += stride) head[i] = (i == 0 || key[i] != key[i - 1]) ? 1 : 0;
}

// run_start[r] = first sorted position of run r; run_start[n_runs] = n is written by the host
__global__ void k_tpl_run_starts(const int *__restrict__ head, const int *__restrict__ run_of, long long n,
                                 int *__restrict__ run_start) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (; i < n; i += stride)
    if (head[i]) run_start[run_of[i]] = (int)i;
}

// runs with at least min_rows members: (start, count) pairs, at most `cap` of them (n_sel counts all candidates)
__global__ void k_tpl_select(const int *__restrict__ run_start, long long n_runs, long long n, int min_rows, int cap,
                             int *__restrict__ sel, int *__restrict__ n_sel) {
  long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (; r < n_runs; r += stride) {
    const int b = run_start[r];
    const int e = (r + 1 < n_runs) ? run_start[r + 1] : (int)n;
    if (e - b >= min_rows) {
      const int k = atomicAdd(n_sel, 1);
      if (k < cap) {
        sel[2 * k] = b;
        sel[2 * k + 1] = e - b;
      }
    }
  }
}

// one warp per template: raw description of its representative row (the first member of the run)
__global__ void __launch_bounds__(256) k_tpl_extract(PtapArgs a, const int *__restrict__ sel,
                                                     const int *__restrict__ sorted_pos, int n_tpl,
                                                     unsigned char *__restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int wpc = blockDim.x >> 5;
  for (int t = blockIdx.x * wpc + (threadIdx.x >> 5); t < n_tpl; t += gridDim.x * wpc) {
    unsigned char *rec = out + (size_t)t * TPLR_STRIDE;
    const int i = a.rows[sorted_pos[sel[2 * t]]];
    const int mtb = a.mt_rowptr[i], n0 = a.mt_rowptr[i + 1] - mtb;
    const int ib = a.inter_rowptr[i], n1 = a.inter_rowptr[i + 1] - ib;
    const int n2 = a.c_rowptr[i + 1] - a.c_rowptr[i];
    const unsigned char *s1 = a.slot1 + a.s1_off[i];
    const unsigned char *s2 = a.slot2 + a.s2_off[i];
    int T1 = 0;
    for (int q = lane; q < n0; q += 32) {
      const int len = (int)a.mt_alen[mtb + q];
      T1 += len;
      if (q < tpl::MAX_N0) {
        rec[TPLR_LEN1 + q] = (unsigned char)len;
        ((double *)(rec + TPLR_W))[q] = a.mt_val[mtb + q];
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) T1 += __shfl_xor_sync(0xffffffffu, T1, o);
    for (int p = lane; p < T1 && p <= tpl::MAX_T1; p += 32) rec[TPLR_SLOT1 + p] = s1[p];
    int base_off = 0;
    for (int base = 0; base < n1; base += 32) {
      const int q = base + lane;
      int len = 0, beg = 0;
      if (q < n1) {
        len = (int)a.inter_mlen[ib + q];
        beg = a.inter_mbeg[ib + q];
        if (q < tpl::MAX_N1) rec[TPLR_LEN2 + q] = (unsigned char)len;
      }
      int total;
      const int off = warp_excl_scan(len, lane, &total);
      for (int e = 0; e < len; ++e) {
        const int p = base_off + off + e;
        if (p <= tpl::MAX_T2) {
          rec[TPLR_SLOT2 + p] = s2[p];
          ((double *)(rec + TPLR_MVAL))[p] = a.m_val[beg + e];
        }
      }
      base_off += total;
    }
    if (lane == 0) {
      int *hd = (int *)(rec + TPLR_INTS);
      hd[0] = n0;
      hd[1] = n1;
      hd[2] = n2;
      hd[3] = T1;
      hd[4] = base_off;
      hd[5] = 1;
    }
  }
}

// tpl_of[pos] = template of the row at list position pos (-1: none).  sel = (start, count) in sorted order;
// valid[t] = 0 for templates the host compiler rejected.
__global__ void k_tpl_assign(const int *__restrict__ sel, const int *__restrict__ valid, int n_tpl,
                             const int *__restrict__ sorted_pos, const unsigned long long *__restrict__ h2,
                             int *__restrict__ tpl_of) {
  for (int t = blockIdx.y; t < n_tpl; t += gridDim.y) {
    if (!valid[t]) continue;
    const int b = sel[2 * t], cnt = sel[2 * t + 1];
    const unsigned long long ref = h2[sorted_pos[b]];
    for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < cnt; k += gridDim.x * blockDim.x) {
      const int pos = sorted_pos[b + k];
      if (h2[pos] == ref) tpl_of[pos] = t;
    }
  }
}

// flags over the SORTED order (templated rows) and over the LIST order (remaining rows of each slot bin)
__global__ void k_tpl_flags(const int *__restrict__ tpl_of, const int *__restrict__ sorted_pos, long long n,
                            int *__restrict__ f_sorted, int *__restrict__ f_rest) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (; i < n; i += stride) {
    f_sorted[i] = tpl_of[sorted_pos[i]] >= 0 ? 1 : 0;
    f_rest[i] = tpl_of[i] >= 0 ? 0 : 1;
  }
}

__global__ void k_tpl_scatter(const int *__restrict__ rows, const int *__restrict__ tpl_of,
                              const int *__restrict__ sorted_pos, const int *__restrict__ off_sorted,
                              const int *__restrict__ off_rest, long long n, int *__restrict__ tpl_rows,
                              int *__restrict__ rest_rows) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (; i < n; i += stride) {
    const int sp = sorted_pos[i];
    if (tpl_of[sp] >= 0) tpl_rows[off_sorted[i]] = rows[sp];
    if (tpl_of[i] < 0) rest_rows[off_rest[i]] = rows[i];
  }
}

// out[2t], out[2t+1] = range of template t inside tpl_rows
__global__ void k_tpl_ranges(const int *__restrict__ sel, int n_tpl, const int *__restrict__ off_sorted,
                             int *__restrict__ out) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t < n_tpl) {
    out[2 * t] = off_sorted[sel[2 * t]];
    out[2 * t + 1] = off_sorted[sel[2 * t] + sel[2 * t + 1]];
  }
}

// ------------------------------------------------------------------------------------------------
// numeric kernel
// ------------------------------------------------------------------------------------------------
struct TplArgs {
  const unsigned char *blobs;
  const long long *blob_off;  // [n_tpl]
  const int *chunks;          // [3 * n_chunks]: template, first row (index into rows), row count
  int n_chunks;
  const int *rows;            // templated rows, grouped by template, ascending inside a template
  int s_cap, o1_cap, o2_cap;  // per-row buffer sizes (entries)
  int n0_cap;                 // operand starts per row (multiple of 32)
};

__device__ __forceinline__ double tpl_lds(unsigned addr) {
  double v;
  asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(addr));
  return v;
}
__device__ __forceinline__ void tpl_sts(unsigned addr, double v) { asm volatile("st.shared.f64 [%0], %1;" ::"r"(addr), "d"(v) : "memory"); }

// One gather stage for the warp's TWO rows: O[flush] = sum over the lane's run of (COEF ? coef * SRC[src] : SRC[src]).
// Every program word and coefficient is loaded and decoded once for both rows; row B's buffers sit at the constant
// byte offset `delta` from row A's, so both rows use one address register.
template <bool COEF>
__device__ __forceinline__ void tpl_gather2(int S, const unsigned *__restrict__ prog, const double *__restrict__ coef,
                                            unsigned src_base, unsigned out_base, unsigned delta, int lane) {
  double accA = 0.0, accB = 0.0;
  prog += lane;
  coef += lane;
  int s = 0;
  for (; s + 4 <= S; s += 4) {
    unsigned u[4];
    double c[4];
#pragma unroll
    for (int b = 0; b < 4; ++b) {
      u[b] = __ldg(prog + (s + b) * 32);
      if (COEF) c[b] = __ldg(coef + (s + b) * 32);
    }
    double vA[4], vB[4];
#pragma unroll
    for (int b = 0; b < 4; ++b) {
      const unsigned ad = src_base + (u[b] & 0xFFFFu);
      vA[b] = tpl_lds(ad);
      vB[b] = tpl_lds(ad + delta);
    }
#pragma unroll
    for (int b = 0; b < 4; ++b) {
      accA = COEF ? fma(c[b], vA[b], accA) : accA + vA[b];
      accB = COEF ? fma(c[b], vB[b], accB) : accB + vB[b];
      const unsigned f = u[b] >> 16;
      if (f != tpl::NO_FLUSH) {
        tpl_sts(out_base + f, accA);
        tpl_sts(out_base + f + delta, accB);
        accA = 0.0;
        accB = 0.0;
      }
    }
  }
  for (; s < S; ++s) {
    const unsigned u = __ldg(prog + s * 32);
    const unsigned ad = src_base + (u & 0xFFFFu);
    const double vA = tpl_lds(ad), vB = tpl_lds(ad + delta);
    if (COEF) {
      const double c = __ldg(coef + s * 32);
      accA = fma(c, vA, accA);
      accB = fma(c, vB, accB);
    } else {
      accA += vA;
      accB += vB;
    }
    const unsigned f = u >> 16;
    if (f != tpl::NO_FLUSH) {
      tpl_sts(out_base + f, accA);
      tpl_sts(out_base + f + delta, accB);
      accA = 0.0;
      accB = 0.0;
    }
  }
}

// pieces of split destinations: round j adds the (j+1)-th piece of every split destination into it (distinct
// destinations within a round: one lane each); g = ptr[nr + 1] (padded to an even count) then (destination, extra)
// index pairs.  Rare: stage 1 is packed without splits, stage 2 adds its extras while the row is written.
__device__ __forceinline__ void tpl_combine(int nr, const unsigned short *__restrict__ g, double *O, int lane) {
  const unsigned short *pairs = g + ((nr + 2) & ~1);
  for (int rd = 0; rd < nr; ++rd) {
    const int e = (int)__ldg(g + rd + 1);
    for (int k = (int)__ldg(g + rd) + lane; k < e; k += 32) {
      const unsigned pr = __ldg((const unsigned *)pairs + k);  // two 16-bit indices
      O[pr & 0xFFFFu] += O[pr >> 16];
    }
    __syncwarp();
  }
}

// Two rows per warp.  A warp runs the SAME program for two rows of its chunk at once: the two rows' independent load ->
// shared-memory -> accumulate chains interleave (the one-row version of this kernel issued on 55 % of the cycles with
// long-scoreboard stalls on top: 9.8 ms at N_b=184 against 8.2 ms), and the instruction count per row halves (528
// against 976 warp instructions per row, ncu).
#ifndef IIFE_TPL_MINBLOCKS
#define IIFE_TPL_MINBLOCKS 3
#endif
__global__ void __launch_bounds__(256, IIFE_TPL_MINBLOCKS) k_ptap_numeric_tpl(PtapArgs a, TplArgs t) {
  extern __shared__ __align__(16) unsigned char smem[];
  const int lane = threadIdx.x & 31, wic = threadIdx.x >> 5;
  const int wpc = blockDim.x >> 5;
  // per warp: sbegA[n0_cap], sbegB[n0_cap] ints, then for row A and again for row B: S[s_cap], O1[o1_cap], O2[o2_cap]
  const unsigned delta = (unsigned)(((size_t)t.s_cap + t.o1_cap + t.o2_cap) * 8);
  const size_t per_warp = (size_t)t.n0_cap * 8 + 2 * (size_t)delta;
  unsigned char *base = smem + per_warp * wic;
  int *sbegA = (int *)base, *sbegB = sbegA + t.n0_cap;
  double *S = (double *)(base + (size_t)t.n0_cap * 8);
  double *O1 = S + t.s_cap;
  double *O2 = O1 + t.o1_cap;
  double *SB = (double *)((unsigned char *)S + delta), *O1B = (double *)((unsigned char *)O1 + delta),
         *O2B = (double *)((unsigned char *)O2 + delta);
  const unsigned S_sh = (unsigned)__cvta_generic_to_shared(S);
  const unsigned O1_sh = (unsigned)__cvta_generic_to_shared(O1);
  const unsigned O2_sh = (unsigned)__cvta_generic_to_shared(O2);
  if (lane == 0) {  // the zero slot of the write-time combine
    O2[0] = 0.0;
    O2B[0] = 0.0;
  }
  __syncwarp();
  const double *__restrict__ a_val = a.a_val;
  const long long n_warps = (long long)gridDim.x * wpc;
  for (long long ch = (long long)blockIdx.x * wpc + wic; ch < t.n_chunks; ch += n_warps) {
    const int tp = __ldg(t.chunks + 3 * ch), r0 = __ldg(t.chunks + 3 * ch + 1), cnt = __ldg(t.chunks + 3 * ch + 2);
    const unsigned char *blob = t.blobs + __ldg(t.blob_off + tp);
    const tpl::Header *h = (const tpl::Header *)blob;
    const int n0 = h->n0, stg_steps = h->stg_steps, n2 = h->n2, S1 = h->S1, S2 = h->S2;
    const int ng1 = h->ng1, ng2 = h->ng2, wc2 = h->wc2;
    const unsigned *stg = (const unsigned *)(blob + h->off_stg) + lane;
    const double *wt = (const double *)(blob + h->off_w);
    const unsigned *p1 = (const unsigned *)(blob + h->off_p1);
    const unsigned short *g1 = (const unsigned short *)(blob + h->off_g1);
    const double *c2 = (const double *)(blob + h->off_c2);
    const unsigned *p2 = (const unsigned *)(blob + h->off_p2);
    const unsigned short *g2 = (const unsigned short *)(blob + h->off_g2);
    const unsigned *xw = (const unsigned *)(blob + h->off_xw);
    for (int r = 0; r < cnt; r += 2) {
      const int iA = __ldg(t.rows + r0 + r);
      const int iB = __ldg(t.rows + r0 + (r + 1 < cnt ? r + 1 : r));  // odd tail: the last row twice (same values twice)
      const int mtbA = __ldg(a.mt_rowptr + iA), mtbB = __ldg(a.mt_rowptr + iB);
      const int cbA = __ldg(a.c_rowptr + iA), cbB = __ldg(a.c_rowptr + iB);
      for (int q = lane; q < n0; q += 32) {
        sbegA[q] = __ldg(a.mt_abeg + mtbA + q);
        sbegB[q] = __ldg(a.mt_abeg + mtbB + q);
      }
      __syncwarp();
      // ---- staging of both rows: S[slot] = w[q] * A.val[beg[q] + e]; a staging word is slot << 16 | q << 8 | e.  The
      // slots of a phase of 16 lanes lie in 16 different banks, as do the reads of stage 1 (edge colouring of the
      // template compiler).  4 steps x 2 rows of loads in flight per lane.
      {
        int s = 0;
        for (; s + 4 <= stg_steps; s += 4) {
          unsigned m[4];
          double vA[4], vB[4], w[4];
#pragma unroll
          for (int b = 0; b < 4; ++b) m[b] = __ldg(stg + (s + b) * 32);
#pragma unroll
          for (int b = 0; b < 4; ++b) {
            vA[b] = 0.0;
            vB[b] = 0.0;
            w[b] = 0.0;
            if (m[b] != tpl::STG_PAD) {
              const unsigned q = (m[b] >> 8) & 255u;
              const int e = (int)(m[b] & 255u);
              vA[b] = __ldg(a_val + sbegA[q] + e);
              vB[b] = __ldg(a_val + sbegB[q] + e);
              w[b] = __ldg(wt + q);
            }
          }
#pragma unroll
          for (int b = 0; b < 4; ++b)
            if (m[b] != tpl::STG_PAD) {
              S[m[b] >> 16] = w[b] * vA[b];
              SB[m[b] >> 16] = w[b] * vB[b];
            }
        }
        for (; s < stg_steps; ++s) {
          const unsigned m = __ldg(stg + s * 32);
          if (m != tpl::STG_PAD) {
            const unsigned q = (m >> 8) & 255u;
            const int e = (int)(m & 255u);
            const double w = __ldg(wt + q);
            S[m >> 16] = w * __ldg(a_val + sbegA[q] + e);
            SB[m >> 16] = w * __ldg(a_val + sbegB[q] + e);
          }
        }
      }
      __syncwarp();
      // ---- stage 1: intermediate rows
      tpl_gather2<false>(S1, p1, nullptr, S_sh, O1_sh, delta, lane);
      __syncwarp();
      if (ng1) {
        tpl_combine(ng1, g1, O1, lane);
        tpl_combine(ng1, g1, O1B, lane);
      }
      // ---- stage 2: output rows
      tpl_gather2<true>(S2, p2, c2, O1_sh, O2_sh, delta, lane);
      __syncwarp();
      if (ng2 && !wc2) {
        tpl_combine(ng2, g2, O2, lane);
        tpl_combine(ng2, g2, O2B, lane);
      }
      for (int o = lane; o < n2; o += 32) {
        double vA = O2[1 + o], vB = O2B[1 + o];
        if (wc2) {  // pieces of a split destination: (piece 0 + piece 1) + piece 2; index 0 is the zero slot
          const unsigned x = __ldg(xw + o);
          vA = (vA + O2[x & 0xFFFFu]) + O2[x >> 16];
          vB = (vB + O2B[x & 0xFFFFu]) + O2B[x >> 16];
        }
        a.c_val[cbA + o] = vA;
        a.c_val[cbB + o] = vB;
      }
      __syncwarp();
    }
  }
}

}  // namespace iife
