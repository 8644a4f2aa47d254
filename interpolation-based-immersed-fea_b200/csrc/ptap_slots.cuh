// ptap_slots.cuh — hash-free numeric PtAP for ordinary rows (per-row kernel: every row has its own slot plan), included
// by ptap.cu.  Rows that share their structure with many others take the template kernel of ptap_tpl.cuh instead.
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
// Rows qualify when both the intermediate and the output row have <= 256 entries (slot bytes), or, with two bytes per
// term, <= 2040 / <= 512 entries (the wide rows of quadratic and 3-D unfitted backgrounds); everything else stays on the
// hashing kernels of ptap.cu / ptap_warp.cuh.
//
// Kernel notes (second version, measured in round 2: 24.0 -> 18.9 ms at N_b=184 against the first one, which is gone):
//   * the 32 item descriptors of a chunk {w, beg, off, len} are staged once in shared memory (16 B each) and
//     every step reads its item's descriptor with one broadcast LDS.128 instead of 5 shuffles;
//   * the privatised accumulator copies are padded by one double, so that the same slot in different copies
//     falls into different banks;
//   * operand rows longer than the lane group are finished in a compacted second pass;
//   * accumulators are cleared as one contiguous range.
// Summation order per output entry: main pass in item order per lane group, then the compacted second pass, then
// the copies in ascending order — deterministic.
#pragma once


namespace iife {

// Tuning knob (compile time): minBlocks of the slot-plan kernels (4 = 64 registers: ptxas' own choice is 48 + spills)
#ifndef IIFE_SLOT_MINBLOCKS
#define IIFE_SLOT_MINBLOCKS 0
#endif
#if IIFE_SLOT_MINBLOCKS > 0
#define IIFE_SLOT_BOUNDS __launch_bounds__(256, IIFE_SLOT_MINBLOCKS)
#else
#define IIFE_SLOT_BOUNDS __launch_bounds__(256)
#endif

constexpr int SLOT_TAIL_BYTES = 32;  // one lane number per long item of a 32-item chunk
typedef void (*slot_kernel_t)(PtapArgs, int, int);

constexpr int PS2_BATCH = 4;
constexpr int PS2_DESC_BYTES = 32 * 16;

struct __align__(16) SlotDesc {
  double w;
  int beg;
  unsigned offlen;  // off (low 16 bits: <= 32 * 256) | len << 16 (<= 256)
};

template <int LG, class ST>
__device__ __forceinline__ void slot_stage2(int my_beg, int my_len, double my_w, int my_off,
                                            const double *__restrict__ x_val, const ST *__restrict__ slots,
                                            double *hv, int stride, int lane, SlotDesc *desc, unsigned char *tail_src) {
  constexpr int G = 1 << LG, NG = 32 >> LG;
  const int g = lane >> LG, lg = lane & (G - 1);
  double *hv_g = hv + (size_t)g * stride;
  // every lane publishes its item (len == 0 beyond the end of the list)
  {
    SlotDesc d;
    d.w = my_w;
    d.beg = my_beg;
    d.offlen = (unsigned)my_off | ((unsigned)my_len << 16);
    desc[lane] = d;
  }
  __syncwarp();  // descriptor stores visible to the whole warp (the vote below is no memory barrier)
  const unsigned live_mask = __ballot_sync(0xffffffffu, my_len > 0);
  const unsigned long_mask = __ballot_sync(0xffffffffu, my_len > G);
  if (live_mask == 0u) return;
  const int n_items = 32 - __clz(live_mask);  // items are contiguous from lane 0, empty operand rows may sit between
  const int nsteps = (n_items + NG - 1) / NG;
  for (int s0 = 0; s0 < nsteps; s0 += PS2_BATCH) {
    int sl[PS2_BATCH];
    double v[PS2_BATCH];
#pragma unroll
    for (int b = 0; b < PS2_BATCH; ++b) {
      const int it = (s0 + b) * NG + g;
      sl[b] = -1;
      v[b] = 0.0;
      if (it < 32) {
        const SlotDesc d = desc[it];
        const int len = (int)(d.offlen >> 16), off = (int)(d.offlen & 0xffffu);
        if (lg < len) {
          sl[b] = (int)__ldg(slots + (unsigned)(off + lg));
          v[b] = d.w * __ldg(x_val + (unsigned)(d.beg + lg));
        }
      }
    }
#pragma unroll
    for (int b = 0; b < PS2_BATCH; ++b) {
      __syncwarp();
      if (sl[b] >= 0) hv_g[sl[b]] += v[b];
    }
  }
  if (long_mask) {  // operand rows longer than the lane group: compacted second pass
    const int n_long = __popc(long_mask);
    if (my_len > G) tail_src[__popc(long_mask & ((1u << lane) - 1u))] = (unsigned char)lane;
    __syncwarp();
    for (int t0 = 0; t0 < n_long; t0 += NG) {
      const int idx = t0 + g;
      int len = 0, off = 0, beg = 0;
      double w = 0.0;
      if (idx < n_long) {
        const SlotDesc d = desc[tail_src[idx]];
        len = (int)(d.offlen >> 16);
        off = (int)(d.offlen & 0xffffu);
        beg = d.beg;
        w = d.w;
      }
      for (int e = G + lg; __any_sync(0xffffffffu, e < len); e += G) {
        int s1 = -1;
        double vv = 0.0;
        if (e < len) {
          s1 = (int)__ldg(slots + (unsigned)(off + e));
          vv = w * __ldg(x_val + (unsigned)(beg + e));
        }
        __syncwarp();
        if (s1 >= 0) hv_g[s1] += vv;
      }
    }
  }
  __syncwarp();  // descriptors and tail_src are rewritten by the next chunk
}

// ST = unsigned char: rows of bins 5 / 6 (both rows <= 256 entries); unsigned short: the wide rows of bin 7
template <int LG1, int LG2, class ST>
__global__ void IIFE_SLOT_BOUNDS k_ptap_numeric_slots(PtapArgs a, int cap1, int cap2) {
  constexpr int NG1 = 32 >> LG1, NG2 = 32 >> LG2;
  extern __shared__ __align__(16) unsigned char smem[];
  const int lane = threadIdx.x & 31, wic = threadIdx.x >> 5;
  const int wpc = blockDim.x >> 5;
  const int st1 = cap1 + 1, st2 = cap2 + 1;  // padded copy strides (doubles)
  // per warp (16-byte aligned): desc[32], h1v[NG1][st1], h2v[NG2][st2], tail_src[32]
  const size_t acc_doubles = ((size_t)NG1 * st1 + (size_t)NG2 * st2 + 1) & ~(size_t)1;
  const size_t per_warp = PS2_DESC_BYTES + acc_doubles * 8 + SLOT_TAIL_BYTES;
  unsigned char *base = smem + per_warp * wic;
  SlotDesc *desc = (SlotDesc *)base;
  double *h1v = (double *)(base + PS2_DESC_BYTES);
  double *h2v = h1v + (size_t)NG1 * st1;
  unsigned char *tail_src = base + PS2_DESC_BYTES + acc_doubles * 8;
  const int64_t warp_global = (int64_t)blockIdx.x * wpc + wic;
  const int64_t n_warps = (int64_t)gridDim.x * wpc;
  const double *__restrict__ a_val = a.a_val;
  const double *__restrict__ m_val = a.m_val;

  for (int64_t wi = warp_global; wi < a.n_rows; wi += n_warps) {
    const int i = a.rows[wi];
    const int mt_b = __ldg(a.mt_rowptr + i), mt_n = __ldg(a.mt_rowptr + i + 1) - mt_b;
    const int cb = __ldg(a.c_rowptr + i), n2 = __ldg(a.c_rowptr + i + 1) - cb;
    const int ib = __ldg(a.inter_rowptr + i), n1 = __ldg(a.inter_rowptr + i + 1) - ib;
    const ST *s1 = (const ST *)(a.slot1 + a.s1_off[i]);
    const ST *s2 = (const ST *)(a.slot2 + a.s2_off[i]);
    // ---- clear the accumulators: one contiguous range (h1v and h2v are adjacent), 16-byte stores
    {
      double2 *z = (double2 *)h1v;
      const int n2x = (int)(acc_doubles >> 1);
      const double2 z2 = make_double2(0.0, 0.0);
      for (int s = lane; s < n2x; s += 32) z[s] = z2;
    }
    __syncwarp();
    // ---- stage 1: H1[slot] += Mt[i,j] * A[j,e]
    {
      int base_off = 0;
      for (int cbase = 0; cbase < mt_n; cbase += 32) {
        const int q = cbase + lane;
        int my_beg = 0, my_len = 0;
        double my_w = 0.0;
        if (q < mt_n) {
          my_w = __ldg(a.mt_val + mt_b + q);
          if (a.mt_abeg) {
            my_beg = __ldg(a.mt_abeg + mt_b + q);
            my_len = (int)__ldg(a.mt_alen + mt_b + q);
          } else {
            const int j = __ldg(a.mt_col + mt_b + q);
            my_beg = __ldg(a.a_rowptr + j);
            my_len = __ldg(a.a_rowptr + j + 1) - my_beg;
          }
        }
        int total;
        const int my_off = warp_excl_scan(my_len, lane, &total);
        slot_stage2<LG1, ST>(my_beg, my_len, my_w, my_off, a_val, s1 + base_off, h1v, st1, lane, desc, tail_src);
        base_off += total;
      }
    }
    // ---- merge the private copies of the intermediate row (fixed order)
    for (int q = lane; q < n1; q += 32) {
      double v = h1v[q];
#pragma unroll
      for (int gg = 1; gg < NG1; ++gg) v += h1v[(size_t)gg * st1 + q];
      h1v[q] = v;
    }
    __syncwarp();
    // ---- stage 2: H2[slot] += H1[q] * M[k_q, e]
    {
      int base_off = 0;
      for (int cbase = 0; cbase < n1; cbase += 32) {
        const int q = cbase + lane;
        int my_beg = 0, my_len = 0;
        double my_w = 0.0;
        if (q < n1) {
          my_w = h1v[q];
          if (a.inter_mbeg) {
            my_beg = __ldg(a.inter_mbeg + ib + q);
            my_len = (int)__ldg(a.inter_mlen + ib + q);
          } else {
            const int k = __ldg(a.inter_col + ib + q);
            my_beg = __ldg(a.m_rowptr + k);
            my_len = __ldg(a.m_rowptr + k + 1) - my_beg;
          }
        }
        int total;
        const int my_off = warp_excl_scan(my_len, lane, &total);
        slot_stage2<LG2, ST>(my_beg, my_len, my_w, my_off, m_val, s2 + base_off, h2v, st2, lane, desc, tail_src);
        base_off += total;
      }
    }
    // ---- write the row (slot = position in the sorted output row)
    for (int s = lane; s < n2; s += 32) {
      double v = h2v[s];
#pragma unroll
      for (int gg = 1; gg < NG2; ++gg) v += h2v[(size_t)gg * st2 + s];
      a.c_val[cb + s] = v;
    }
    __syncwarp();
  }
}

static size_t slot_per_warp_bytes(int lg1, int lg2, int cap1, int cap2) {
  size_t acc = ((size_t)(32 >> lg1) * (cap1 + 1) + (size_t)(32 >> lg2) * (cap2 + 1) + 1) & ~(size_t)1;
  return PS2_DESC_BYTES + acc * 8 + SLOT_TAIL_BYTES;
}

static slot_kernel_t pick_slot_kernel(int lg1, int lg2, bool wide = false) {
  if (wide) {  // at most two copies of the 2040-entry intermediate row and four of the 512-entry output row
#define PSKW(a_, b_) \
  if (lg1 == a_ && lg2 == b_) return k_ptap_numeric_slots<a_, b_, unsigned short>;
    PSKW(4, 3) PSKW(4, 4) PSKW(4, 5) PSKW(5, 3) PSKW(5, 4) PSKW(5, 5)
#undef PSKW
    return nullptr;
  }
#define PSK2(a_, b_) \
  if (lg1 == a_ && lg2 == b_) return k_ptap_numeric_slots<a_, b_, unsigned char>;
  PSK2(3, 2) PSK2(3, 3) PSK2(3, 4) PSK2(3, 5)
  PSK2(4, 2) PSK2(4, 3) PSK2(4, 4) PSK2(4, 5)
  PSK2(5, 2) PSK2(5, 3) PSK2(5, 4) PSK2(5, 5)
#undef PSK2
  return nullptr;
}

}  // namespace iife
