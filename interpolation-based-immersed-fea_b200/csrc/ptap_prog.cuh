// ptap_prog.cuh — numeric PtAP for rows of the small bin (intermediate row <= 128, output row <= 32 entries) that
// did NOT find a template (ptap_tpl.cuh): stage 2 runs a per-row gather program built once per (plan, values of M).
// Measured in round 2 at N_b=184 with templates disabled: 11.7 ms against 18.9 ms for the slot kernel alone
// (ptap_slots.cuh); it costs 9 bytes of program per padded stage-2 term (33 GB for the whole cube), so it is used
// only while the program fits half the free device memory (IIFE_PTAP_PROG=0 disables it).
//
// Stage 1 (H1 = Mt[i,:] * A) is the slot-plan code (ptap_slots.cuh).  Stage 2 (A_b[i,:] = H1 * M[K_i,:]) no longer
// scatters: M does not change between numeric calls in the reference's workflow (only A_f does), so the stage is
// compiled once per (plan, values of M) into a GATHER PROGRAM per output row:
//   * the terms of the row are grouped by output entry l; entry l owns g_l = ceil(cnt_l / S) consecutive lanes,
//     S = the smallest step count for which all groups fit the 32 lanes;
//   * per step and lane the program stores the M value itself (8 B: the loads are coalesced) and the index of the
//     intermediate entry it multiplies (1 B; 255 = no term);
//   * the kernel runs  acc += val * H1[idx]  in a register for S steps, adds the lanes of a group with shuffles
//     (fixed order) and the group's first lane writes the entry: no shared-memory read-modify-write, no private
//     copies, no merge.
// Term order inside a group: items in ascending order of the intermediate column, 32 at a time, entry by entry
// (ranked with match.any) — fixed by the plan, hence deterministic; not the rounding of v1/v2.
#pragma once

namespace iife {

constexpr int PROG_IDLE = 255;

struct ProgArgs {
  int *steps;                // [n_rows]  S per row (count pass output)
  const long long *off;      // [n_rows+1] first step of the row in the program
  unsigned char *lane_out;   // [n_rows*32] output entry fed by each lane (255: idle)
  unsigned char *maxg;       // [n_rows]  largest group of the row
  unsigned char *slot;       // [total_steps*32]
  double *val;               // [total_steps*32]
};

// One warp per row.  FILL = false: steps[] only.  FILL = true: lane map + program.
template <bool FILL>
__global__ void __launch_bounds__(256) k_prog_build(PtapArgs a, ProgArgs pg) {
  __shared__ int s_cnt[8][32];
  __shared__ int s_ls[8][32];
  __shared__ int s_cur[8][32];
  const int lane = threadIdx.x & 31, wic = threadIdx.x >> 5;
  const int wpc = blockDim.x >> 5;
  int *cnt = s_cnt[wic], *ls = s_ls[wic], *cur = s_cur[wic];
  const long long n_warps = (long long)gridDim.x * wpc;
  for (long long wi = (long long)blockIdx.x * wpc + wic; wi < a.n_rows; wi += n_warps) {
    const int i = a.rows[wi];
    const int ib = __ldg(a.inter_rowptr + i), n1 = __ldg(a.inter_rowptr + i + 1) - ib;
    const int n2 = __ldg(a.c_rowptr + i + 1) - __ldg(a.c_rowptr + i);
    const unsigned char *s2 = a.slot2 + a.s2_off[i];
    cnt[lane] = 0;
    cur[lane] = 0;
    __syncwarp();
    // ---- terms per output entry
    {
      int base_off = 0;
      for (int cbase = 0; cbase < n1; cbase += 32) {
        const int q = cbase + lane;
        int my_len = 0;
        if (q < n1) {
          if (a.inter_mlen) my_len = (int)__ldg(a.inter_mlen + ib + q);
          else {
            const int k = __ldg(a.inter_col + ib + q);
            my_len = __ldg(a.m_rowptr + k + 1) - __ldg(a.m_rowptr + k);
          }
        }
        int total;
        const int my_off = warp_excl_scan(my_len, lane, &total);
        for (int e = 0; e < my_len; ++e) atomicAdd(&cnt[s2[base_off + my_off + e]], 1);
        base_off += total;
      }
    }
    __syncwarp();
    const int c_l = (lane < n2) ? cnt[lane] : 0;
    int total_terms = c_l;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) total_terms += __shfl_xor_sync(0xffffffffu, total_terms, o);
    int S;
    if (!FILL) {
      S = (total_terms + 31) >> 5;
      if (S < 1) S = 1;
      for (;;) {
        int need = (c_l + S - 1) / S;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) need += __shfl_xor_sync(0xffffffffu, need, o);
        if (need <= 32) break;  // terminates: S = max count gives one lane per entry, n2 <= 32
        ++S;
      }
      if (lane == 0) pg.steps[wi] = S;
      __syncwarp();
      continue;
    }
    S = pg.steps[wi];
    const long long off = pg.off[wi];
    const int g_l = (c_l + S - 1) / S;
    int total_lanes;
    const int ls_l = warp_excl_scan(g_l, lane, &total_lanes);
    ls[lane] = ls_l;
    int mg = g_l;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mg = max(mg, __shfl_xor_sync(0xffffffffu, mg, o));
    // lane map: lane L feeds the entry whose group contains it
    {
      unsigned char *lo = pg.lane_out + wi * 32;
      lo[lane] = (unsigned char)PROG_IDLE;
      __syncwarp();
      for (int t = 0; t < g_l; ++t) lo[ls_l + t] = (unsigned char)lane;
      if (lane == 0) pg.maxg[wi] = (unsigned char)mg;
    }
    __syncwarp();
    // ---- place every term: entry-by-entry over chunks of 32 items, ranks from match.any (deterministic)
    {
      int base_off = 0;
      for (int cbase = 0; cbase < n1; cbase += 32) {
        const int q = cbase + lane;
        int my_len = 0, my_beg = 0;
        if (q < n1) {
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
        int max_len = my_len;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) max_len = max(max_len, __shfl_xor_sync(0xffffffffu, max_len, o));
        for (int e = 0; e < max_len; ++e) {
          const bool has = e < my_len;
          const unsigned active = __ballot_sync(0xffffffffu, has);
          if (has) {
            const int l = (int)s2[base_off + my_off + e];
            const unsigned same = __match_any_sync(active, l);
            const int rank = __popc(same & ((1u << lane) - 1u));
            const int pos = cur[l] + rank;
            __syncwarp(active);
            if (rank == 0) cur[l] = pos + __popc(same);
            const int tl = ls[l] + pos / S, st = pos - (pos / S) * S;
            const long long at = (off + st) * 32 + tl;
            pg.slot[at] = (unsigned char)q;
            pg.val[at] = __ldg(a.m_val + my_beg + e);
          }
          __syncwarp();
        }
        base_off += total;
      }
    }
    __syncwarp();
  }
}

template <int LG1>
__global__ void IIFE_SLOT_BOUNDS k_ptap_numeric_prog(PtapArgs a, ProgArgs pg, int cap1) {
  constexpr int NG1 = 32 >> LG1;
  extern __shared__ __align__(16) unsigned char smem[];
  const int lane = threadIdx.x & 31, wic = threadIdx.x >> 5;
  const int wpc = blockDim.x >> 5;
  const int st1 = cap1 + 1;
  // per warp (16-byte aligned): desc[32], h1v[NG1][st1], tail_src[32]
  const size_t acc_doubles = ((size_t)NG1 * st1 + 1) & ~(size_t)1;
  const size_t per_warp = PS2_DESC_BYTES + acc_doubles * 8 + SLOT_TAIL_BYTES;
  unsigned char *base = smem + per_warp * wic;
  SlotDesc *desc = (SlotDesc *)base;
  double *h1v = (double *)(base + PS2_DESC_BYTES);
  unsigned char *tail_src = base + PS2_DESC_BYTES + acc_doubles * 8;
  const long long n_warps = (long long)gridDim.x * wpc;
  const double *__restrict__ a_val = a.a_val;

  for (long long wi = (long long)blockIdx.x * wpc + wic; wi < a.n_rows; wi += n_warps) {
    const int i = a.rows[wi];
    const int mt_b = __ldg(a.mt_rowptr + i), mt_n = __ldg(a.mt_rowptr + i + 1) - mt_b;
    const int cb = __ldg(a.c_rowptr + i);
    const int n1 = __ldg(a.inter_rowptr + i + 1) - __ldg(a.inter_rowptr + i);
    const unsigned char *s1 = a.slot1 + a.s1_off[i];
    {
      double2 *z = (double2 *)h1v;
      const int n2x = (int)(acc_doubles >> 1);
      const double2 z2 = make_double2(0.0, 0.0);
      for (int s = lane; s < n2x; s += 32) z[s] = z2;
    }
    __syncwarp();
    // ---- stage 1 (as in ptap_slots2.cuh)
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
        slot_stage2<LG1>(my_beg, my_len, my_w, my_off, a_val, s1 + base_off, h1v, st1, lane, desc, tail_src);
        base_off += total;
      }
    }
    for (int q = lane; q < n1; q += 32) {
      double v = h1v[q];
#pragma unroll
      for (int gg = 1; gg < NG1; ++gg) v += h1v[(size_t)gg * st1 + q];
      h1v[q] = v;
    }
    __syncwarp();
    // ---- stage 2: the gather program of this row
    {
      const int S = pg.steps[wi];
      const long long at0 = pg.off[wi] * 32 + lane;
      const int my_out = (int)pg.lane_out[wi * 32 + lane];
      const int mg = (int)pg.maxg[wi];
      double acc = 0.0;
      for (int t = 0; t < S; ++t) {
        const int s = (int)__ldcs(pg.slot + at0 + (long long)t * 32);
        const double v = __ldcs(pg.val + at0 + (long long)t * 32);
        if (s != PROG_IDLE) acc = fma(v, h1v[s], acc);
      }
      double sum = acc;
      for (int d = 1; d < mg; ++d) {  // lanes of a group are consecutive: the first one adds its followers in order
        const double o = __shfl_down_sync(0xffffffffu, acc, d);
        const int oo = __shfl_down_sync(0xffffffffu, my_out, d);
        if (lane + d < 32 && oo == my_out) sum += o;
      }
      const int prev = __shfl_up_sync(0xffffffffu, my_out, 1);
      if (my_out != PROG_IDLE && (lane == 0 || prev != my_out)) a.c_val[cb + my_out] = sum;
    }
    __syncwarp();
  }
}

typedef void (*prog_kernel_t)(PtapArgs, ProgArgs, int);
static prog_kernel_t pick_prog_kernel(int lg1) {
  if (lg1 == 3) return k_ptap_numeric_prog<3>;
  if (lg1 == 4) return k_ptap_numeric_prog<4>;
  if (lg1 == 5) return k_ptap_numeric_prog<5>;
  return nullptr;
}
static size_t prog_per_warp_bytes(int lg1, int cap1) {
  size_t acc = ((size_t)(32 >> lg1) * (cap1 + 1) + 1) & ~(size_t)1;
  return PS2_DESC_BYTES + acc * 8 + SLOT_TAIL_BYTES;
}

}  // namespace iife
