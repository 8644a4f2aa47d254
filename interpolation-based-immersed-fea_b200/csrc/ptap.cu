// ptap.cu — two-phase sparse Galerkin triple product A_b = M^T A_f M.
//
// Replaces the reference's AT_R_A (la_utils.py:165-182): two in-place MatTranspose calls and two
// MatMatMult(MAT_INITIAL_MATRIX) calls, i.e. ((M^T) A_f) M with the intermediate M^T A_f materialised
// and both symbolic phases redone on every call.
//
// Here: one output row i of A_b is owned by one *team* (a warp for ordinary rows, a whole CTA for
// fat rows).  The team runs both products back to back and keeps the intermediate row
// (M^T A_f)[i,:] in a hash table in shared memory, so the intermediate never touches HBM:
//
//   stage 1   H1[k] += Mt[i,j] * A_f[j,k]     j in Mt row i (ascending), k in A_f row j
//   stage 2   H2[l] += H1[k]   * M[k,l]       k in H1 (table order),     l in M row k
//   write     C.val[row i] = H2[C.col[row i]] (columns ascending, as PETSc stores them)
//
// Within a team, G lanes cooperate on one operand row (G = power of two near the mean row length of
// that operand), so a warp reads 32/G operand rows per step with coalesced (colind, val) runs.
// The symbolic phase runs the same traversal on keys only (structural product: stored zeros count,
// numeric cancellation never removes an entry — SURVEY A.2), first counting, then filling sorted
// column lists.  Table sizes come from a ladder of levels; a row whose table overflows at one level
// is retried at the next, the last level keeps its tables in global memory, so any input works.
#include "common.cuh"
#include <algorithm>
#include <list>
#include <stdlib.h>
#include <chrono>

namespace iife {

// IIFE_PLAN_DEBUG=1: wall time of the stages of the symbolic phase and of the template build (stderr)
struct PlanTimer {
  bool on;
  std::chrono::steady_clock::time_point t0;
  PlanTimer() : on(getenv("IIFE_PLAN_DEBUG") != nullptr), t0(std::chrono::steady_clock::now()) {}
  void lap(const char *what) {
    if (!on) return;
    cudaStreamSynchronize(ctx().stream);
    auto t1 = std::chrono::steady_clock::now();
    fprintf(stderr, "[plan] %-28s %8.3f ms\n", what, std::chrono::duration<double, std::milli>(t1 - t0).count());
    t0 = t1;
  }
};

constexpr int EMPTY = 0x7fffffff;
constexpr unsigned HASH_MUL = 2654435761u;

struct Level {
  int warp_team;      // 1: one warp per row, 0: one CTA per row
  int threads;        // CTA size
  int log_cap1, log_cap2;  // table sizes (0 = tables in global memory, sized per plan)
};

// symbolic ladder (keys only: 4 B per slot)
static const Level SYM_LEVELS[4] = {{1, 256, 10, 8}, {1, 128, 12, 10}, {0, 256, 15, 14}, {0, 256, 0, 0}};
// numeric ladder (key + fp64 value: 12 B per slot); load factor <= 0.5 from the exact counts
static const Level NUM_LEVELS[5] = {{1, 256, 8, 6}, {1, 128, 10, 8}, {0, 128, 11, 10}, {0, 256, 13, 12}, {0, 256, 0, 0}};
constexpr int N_SYM_LEVELS = 4;
constexpr int N_NUM_LEVELS = 5;
constexpr int N_BINS = 8;  // numeric bins: 0..4 hashing levels, 5..7 slot-plan rows (ptap_slots.cuh); 7 = wide rows, 2-byte slots
constexpr int N_SLOT_BINS = 3;
constexpr int SLOT_CAP1[N_SLOT_BINS] = {128, 256, 2040}, SLOT_CAP2[N_SLOT_BINS] = {32, 256, 512};

struct Plan {
  uint64_t fpM = 0, fpA = 0, fpR = 0;
  int64_t n_f = 0, n_b = 0, nnzM = 0, nnzA = 0;  // n_b = rows of the result; n_f = rows of A
  int64_t n_k = 0, n_ccols = 0, nnzR = 0;         // columns of A (= rows of P), columns of the result
  bool general_rap = false;  // true: C = R A P with an explicit restriction R given by the caller
  Mat *MT = nullptr;       // PtAP: explicit transpose of M (pattern + values refreshed per numeric call)
  int *mt_perm = nullptr;  // MT.val[p] = M.val[mt_perm[p]]
  uint64_t mt_vals_uid = 0, mt_vals_version = 0;
  int *c_rowptr = nullptr, *c_colind = nullptr;
  int64_t nnz_c = 0, nnz_inter = 0;
  int *n1 = nullptr;          // exact size of the intermediate row (M^T A_f)[i,:]
  int *bin_rows = nullptr;    // rows grouped by numeric level, ascending inside a level
  int64_t bin_off[N_BINS + 1] = {0};
  // slot plan
  unsigned char *slot1 = nullptr, *slot2 = nullptr;
  long long *s1_off = nullptr, *s2_off = nullptr;
  int *inter_rowptr = nullptr, *inter_col = nullptr;
  int64_t s1_total = 0, s2_total = 0, inter_total = 0;
  int *mt_abeg = nullptr, *inter_mbeg = nullptr;
  unsigned char *mt_alen = nullptr, *inter_mlen = nullptr;
  bool packed_meta = false;
  int logG1 = 4, logG2 = 2;
  // global-memory tables of the last numeric level
  int g_log_cap1 = 0, g_log_cap2 = 0, g_ctas = 0;
  int wide_cap1 = 0, wide_cap2 = 0;  // accumulator sizes of bin 7 (longest intermediate / output row in it)
  int *g_keys = nullptr;
  double *g_vals = nullptr;
  size_t g_keys_n = 0, g_vals_n = 0;
  int *err_flag = nullptr;
  // v3 stage-2 gather program of the small-row bin (ptap_prog.cuh, experimental)
  int pg_state = 0;  // 0 not built, 1 built, -1 not possible (memory)
  int64_t pg_rows = 0, pg_total_steps = 0;
  int *pg_steps = nullptr;
  long long *pg_off = nullptr;
  unsigned char *pg_lane_out = nullptr, *pg_maxg = nullptr, *pg_slot = nullptr;
  double *pg_val = nullptr;
  uint64_t pg_m_uid = 0, pg_m_version = 0;
  // template plan (ptap_tpl.cuh): rows of the slot-plan bins whose structure AND M / R values are shared by many
  // rows run a precompiled gather program; the others ("rest") stay on the per-row kernels
  int tp_state = 0;  // 0 not built, 1 built (possibly with no template)
  uint64_t tp_m_uid = 0, tp_m_version = 0, tp_r_uid = 0, tp_r_version = 0;
  int tp_n_tpl = 0, tp_n_chunks = 0, tp_chunk_rows = 0;
  int64_t tp_n_rows = 0, tp_list_n = 0, tp_rest5 = 0, tp_rest6 = 0;
  unsigned char *tp_blobs = nullptr;
  size_t tp_blob_bytes = 0;
  long long *tp_blob_off = nullptr;
  int *tp_chunks = nullptr, *tp_rows = nullptr, *tp_rest_rows = nullptr;
  int tp_s_cap = 0, tp_o1_cap = 0, tp_o2_cap = 0, tp_n0_cap = 32;
  double tp_use1 = 0.0, tp_use2 = 0.0;  // mean lane use of the two gather stages, weighted by rows
};

// ------------------------------------------------------------------------------------------------
// device helpers
// ------------------------------------------------------------------------------------------------
template <bool WARP>
__device__ __forceinline__ void team_sync() {
  if (WARP) __syncwarp();
  else __syncthreads();
}

__device__ __forceinline__ int h_insert(int *hk, unsigned mask, int shift, int key) {
  unsigned h = ((unsigned)key * HASH_MUL) >> shift;
  for (unsigned probes = 0; probes <= mask; ++probes) {
    int old = ((volatile int *)hk)[h];
    if (old == key) return (int)h;
    if (old == EMPTY) {
      old = atomicCAS(&hk[h], EMPTY, key);
      if (old == EMPTY || old == key) return (int)h;
    }
    h = (h + 1) & mask;
  }
  return -1;
}

__device__ __forceinline__ int h_find(const int *hk, unsigned mask, int shift, int key) {
  unsigned h = ((unsigned)key * HASH_MUL) >> shift;
  for (unsigned probes = 0; probes <= mask; ++probes) {
    int cur = hk[h];
    if (cur == key) return (int)h;
    if (cur == EMPTY) return -1;
    h = (h + 1) & mask;
  }
  return -1;
}

// Accumulate  H[c] += w_q * X[r_q, c]  over the items q = 0..n_items-1, item q = (ik[q], iw[q]).
// Items with key EMPTY are skipped (table walked without compaction).  T threads, thread index t.
// G = 1 << logG lanes share one item.  INSERT: keys are created on demand (CAS); otherwise the key
// must already be in the table.  NUMERIC: values are accumulated, otherwise keys only.
// Returns (in *fail) nonzero if an insert/find could not be served (overflow or pattern mismatch).
template <bool WARP, bool NUMERIC, bool INSERT>
__device__ __forceinline__ void accumulate_items(int T, int t, int n_items, const int *ik, const double *iw,
                                                 const int *__restrict__ x_rowptr, const int *__restrict__ x_col,
                                                 const double *__restrict__ x_val, int logG, int *hk, double *hv,
                                                 unsigned mask, int shift, int *st_beg, int *st_len, double *st_w,
                                                 int *fail) {
  const int G = 1 << logG;
  const int g = t >> logG, lg = t & (G - 1), NG = T >> logG;
  for (int base = 0; base < n_items; base += T) {
    int q = base + t;
    int len = 0;
    if (q < n_items) {
      int r = ik[q];
      if (r != EMPTY) {
        int b = __ldg(x_rowptr + r);
        len = __ldg(x_rowptr + r + 1) - b;
        st_beg[t] = b;
        if (NUMERIC) st_w[t] = iw[q];
      }
    }
    st_len[t] = len;
    team_sync<WARP>();
    int cnt = min(T, n_items - base);
    for (int it = g; it < cnt; it += NG) {
      int l = st_len[it];
      if (l == 0) continue;
      int b = st_beg[it];
      double w = NUMERIC ? st_w[it] : 0.0;
      for (int e = lg; e < l; e += G) {
        int c = __ldg(x_col + b + e);
        int slot = INSERT ? h_insert(hk, mask, shift, c) : h_find(hk, mask, shift, c);
        if (slot < 0) {
          *fail = 1;
        } else if (NUMERIC) {
          atomicAdd(&hv[slot], w * __ldg(x_val + b + e));
        }
      }
    }
    team_sync<WARP>();
  }
}

// ordered in-place compaction of a table by one warp; returns the number of occupied slots
__device__ __forceinline__ int warp_compact(int *hk, double *hv, int cap, int lane) {
  int n = 0;
  for (int base = 0; base < cap; base += 32) {
    int k = hk[base + lane];
    double v = hv ? hv[base + lane] : 0.0;
    unsigned m = __ballot_sync(0xffffffffu, k != EMPTY);
    __syncwarp();
    if (k != EMPTY) {
      int pos = n + __popc(m & ((1u << lane) - 1u));
      hk[pos] = k;
      if (hv) hv[pos] = v;
    }
    n += __popc(m);
    __syncwarp();
  }
  return n;
}

// number of occupied slots, CTA team (result in every thread)
__device__ __forceinline__ int cta_count(const int *hk, int cap, int *scratch) {
  int c = 0;
  for (int s = threadIdx.x; s < cap; s += blockDim.x) c += (hk[s] != EMPTY);
  __syncthreads();
  if (threadIdx.x == 0) *scratch = 0;
  __syncthreads();
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) c += __shfl_down_sync(0xffffffffu, c, o);
  if ((threadIdx.x & 31) == 0 && c) atomicAdd(scratch, c);
  __syncthreads();
  int r = *scratch;
  __syncthreads();
  return r;
}

// ascending bitonic sort of a[0..P), P a power of two
template <bool WARP>
__device__ __forceinline__ void team_bitonic(int *a, int P, int T, int t) {
  for (int size = 2; size <= P; size <<= 1) {
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      for (int x = t; x < (P >> 1); x += T) {
        int lo = 2 * x - (x & (stride - 1));
        int hi = lo + stride;
        bool asc = ((lo & size) == 0);
        int kl = a[lo], kh = a[hi];
        if ((kl > kh) == asc) {
          a[lo] = kh;
          a[hi] = kl;
        }
      }
      team_sync<WARP>();
    }
  }
}

struct RowTables {
  int *h1k, *h2k;
  double *h1v, *h2v;
  int *st_beg, *st_len;
  double *st_w;
  int *scratch;
};

// carve the team's tables out of dynamic shared memory (or the global workspace)
template <bool WARP, bool NUMERIC>
__device__ __forceinline__ RowTables carve(unsigned char *smem, int cap1, int cap2, int T, int team_in_cta,
                                           int *g_keys, double *g_vals, int team_global) {
  RowTables rt;
  // per-team shared layout: [st_w T doubles][h1v cap1][h2v cap2][h1k cap1][h2k cap2][st_beg T][st_len T][scratch 2]
  bool global_tables = (g_keys != nullptr);
  size_t c1 = global_tables ? 0 : (size_t)cap1, c2 = global_tables ? 0 : (size_t)cap2;
  size_t per_team = (size_t)T * 8 + (NUMERIC ? (c1 + c2) * 8 : 0) + (c1 + c2) * 4 + (size_t)T * 8 + 8;
  per_team = (per_team + 15) & ~(size_t)15;
  unsigned char *p = smem + per_team * team_in_cta;
  rt.st_w = (double *)p;
  p += (size_t)T * 8;
  if (!global_tables) {
    if (NUMERIC) {
      rt.h1v = (double *)p;
      p += c1 * 8;
      rt.h2v = (double *)p;
      p += c2 * 8;
    } else {
      rt.h1v = rt.h2v = nullptr;
    }
    rt.h1k = (int *)p;
    p += c1 * 4;
    rt.h2k = (int *)p;
    p += c2 * 4;
  } else {
    rt.h1k = g_keys + (size_t)team_global * ((size_t)cap1 + cap2);
    rt.h2k = rt.h1k + cap1;
    rt.h1v = NUMERIC ? g_vals + (size_t)team_global * ((size_t)cap1 + cap2) : nullptr;
    rt.h2v = NUMERIC ? rt.h1v + cap1 : nullptr;
  }
  rt.st_beg = (int *)p;
  p += (size_t)T * 4;
  rt.st_len = (int *)p;
  p += (size_t)T * 4;
  rt.scratch = (int *)p;
  return rt;
}

static size_t team_smem_bytes(bool numeric, bool global_tables, int cap1, int cap2, int T) {
  size_t c1 = global_tables ? 0 : (size_t)cap1, c2 = global_tables ? 0 : (size_t)cap2;
  size_t per_team = (size_t)T * 8 + (numeric ? (c1 + c2) * 8 : 0) + (c1 + c2) * 4 + (size_t)T * 8 + 8;
  return (per_team + 15) & ~(size_t)15;
}

struct PtapArgs {
  // operands
  const int *mt_rowptr, *mt_col;
  const double *mt_val;
  const int *a_rowptr, *a_col;
  const double *a_val;
  const int *m_rowptr, *m_col;
  const double *m_val;
  // output pattern / values
  const int *c_rowptr;
  int *c_col;
  double *c_val;
  // row selection
  const int *rows;      // list of rows (nullptr: all rows 0..n_rows)
  const int *n_list;    // device count of the list (nullptr: use n_rows)
  int64_t n_rows;
  // symbolic outputs
  int *n1, *n2;         // per-row counts
  signed char *level;   // level at which the row was resolved (symbolic)
  int this_level;
  int *ovf_rows, *n_ovf;  // rows to retry at the next level
  // geometry
  int log_cap1, log_cap2, logG1, logG2;
  int *g_keys;
  double *g_vals;
  int *err_flag;
  // slot plan (ptap_slots.cuh): the symbolic phase records, for every product term of a "slot row",
  // the dense index of its destination (rank of the key in the sorted intermediate / output row)
  int *e1, *e2;                     // count pass: number of stage-1 / stage-2 product terms of the row
  const unsigned char *slot_row;    // fill pass: 1 if the row gets a slot plan
  const long long *s1_off, *s2_off; // per-row offsets into slot1 / slot2
  unsigned char *slot1, *slot2;
  const int *inter_rowptr;          // sorted pattern of the intermediate rows (slot rows only)
  int *inter_col;
  // packed operand-row metadata (start, length) per Mt entry / per intermediate entry: replaces the
  // dependent rowptr gathers of the numeric kernel by coalesced loads (nullptr: gather from rowptr)
  const int *mt_abeg, *inter_mbeg;
  const unsigned char *mt_alen, *inter_mlen;
};

__device__ __forceinline__ int warp_excl_scan(int v, int lane, int *total) {
  int incl = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    int t = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += t;
  }
  *total = __shfl_sync(0xffffffffu, incl, 31);
  return incl - v;
}

__device__ __forceinline__ int rank_sorted(const int *a, int n, int key) {
  int lo = 0, hi = n;
  while (lo < hi) {
    int mid = (lo + hi) >> 1;
    if (a[mid] < key) lo = mid + 1;
    else hi = mid;
  }
  return lo;
}

// For the items (row ids in `ik`, n_items of them) of operand X, write for every entry the rank of its
// column in the sorted key list `keys[0..nk)` as one byte (two for wide rows) at out[running offset].  One warp.
__device__ __forceinline__ void warp_emit_slots(int n_items, const int *ik, const int *__restrict__ x_rowptr,
                                                const int *__restrict__ x_col, int logG, const int *keys, int nk,
                                                unsigned char *out, int lane, bool wide) {
  const int G = 1 << logG, NG = 32 >> logG;
  const int g = lane >> logG, lg = lane & (G - 1);
  long long base_off = 0;
  for (int base = 0; base < n_items; base += 32) {
    int q = base + lane;
    int beg = 0, len = 0;
    if (q < n_items) {
      int r = ik[q];
      beg = __ldg(x_rowptr + r);
      len = __ldg(x_rowptr + r + 1) - beg;
    }
    int total;
    int off = warp_excl_scan(len, lane, &total);
    int cnt = min(32, n_items - base);
    for (int it0 = 0; it0 < cnt; it0 += NG) {
      int it = it0 + g;
      int src = it & 31;
      int b = __shfl_sync(0xffffffffu, beg, src);
      int l = __shfl_sync(0xffffffffu, len, src);
      int o = __shfl_sync(0xffffffffu, off, src);
      if (it < cnt)
        for (int e = lg; e < l; e += G) {
          const int rk = rank_sorted(keys, nk, __ldg(x_col + b + e));
          if (wide) ((unsigned short *)out)[base_off + o + e] = (unsigned short)rk;  // rows of bin 7: two bytes per term
          else out[base_off + o + e] = (unsigned char)rk;
        }
    }
    base_off += total;
  }
}

// ------------------------------------------------------------------------------------------------
// symbolic kernels.  MODE 0: count (n1, n2, overflow list);  MODE 1: fill sorted C.col
// ------------------------------------------------------------------------------------------------
template <bool WARP, int MODE>
__global__ void k_ptap_symbolic(PtapArgs a) {
  extern __shared__ __align__(16) unsigned char smem[];
  const int T = WARP ? 32 : blockDim.x;
  const int t = WARP ? (threadIdx.x & 31) : threadIdx.x;
  const int team_in_cta = WARP ? (threadIdx.x >> 5) : 0;
  const int teams_per_cta = WARP ? (blockDim.x >> 5) : 1;
  const int64_t team_global = (int64_t)blockIdx.x * teams_per_cta + team_in_cta;
  const int64_t n_teams = (int64_t)gridDim.x * teams_per_cta;
  const int cap1 = 1 << a.log_cap1, cap2 = 1 << a.log_cap2;
  const unsigned mask1 = cap1 - 1, mask2 = cap2 - 1;
  const int shift1 = 32 - a.log_cap1, shift2 = 32 - a.log_cap2;
  RowTables rt = carve<WARP, false>(smem, cap1, cap2, T, team_in_cta, a.g_keys, nullptr, (int)team_global);
  const int64_t n_work = a.n_list ? (int64_t)*a.n_list : a.n_rows;

  for (int64_t wi = team_global; wi < n_work; wi += n_teams) {
    const int i = a.rows ? a.rows[wi] : (int)wi;
    if (MODE == 1 && a.level[i] != a.this_level) continue;
    if (MODE == 0 && a.rows == nullptr && a.level[i] != -1) continue;
    const int mt_b = a.mt_rowptr[i], mt_n = a.mt_rowptr[i + 1] - mt_b;
    if (mt_n == 0) {
      if (MODE == 0 && t == 0) {
        a.n1[i] = 0;
        a.n2[i] = 0;
        a.level[i] = (signed char)a.this_level;
      }
      continue;
    }
    for (int s = t; s < cap1; s += T) rt.h1k[s] = EMPTY;
    for (int s = t; s < cap2; s += T) rt.h2k[s] = EMPTY;
    if (t == 0) rt.scratch[1] = 0;
    team_sync<WARP>();
    // stage 1: keys of (M^T A)[i,:]
    accumulate_items<WARP, false, true>(T, t, mt_n, a.mt_col + mt_b, nullptr, a.a_rowptr, a.a_col, nullptr, a.logG1,
                                        rt.h1k, nullptr, mask1, shift1, rt.st_beg, rt.st_len, rt.st_w, &rt.scratch[1]);
    int n1;
    if (WARP) n1 = warp_compact(rt.h1k, nullptr, cap1, t);
    else n1 = cta_count(rt.h1k, cap1, rt.scratch);
    bool ovf = (rt.scratch[1] != 0) || (n1 > (cap1 / 4) * 3);
    int n2 = 0;
    const bool slot_fill = WARP && MODE == 1 && a.slot_row && a.slot_row[i] && !ovf;
    if (WARP && MODE == 0 && a.e1 && !ovf) {
      // product-term counts of the two stages (sizes of the slot plan)
      int s1 = 0, s2 = 0;
      for (int q = t; q < mt_n; q += 32) {
        int j = a.mt_col[mt_b + q];
        s1 += a.a_rowptr[j + 1] - a.a_rowptr[j];
      }
      for (int q = t; q < n1; q += 32) {
        int k = rt.h1k[q];
        s2 += a.m_rowptr[k + 1] - a.m_rowptr[k];
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        s1 += __shfl_xor_sync(0xffffffffu, s1, o);
        s2 += __shfl_xor_sync(0xffffffffu, s2, o);
      }
      if (t == 0) {
        a.e1[i] = s1;
        a.e2[i] = s2;
      }
    }
    if (slot_fill) {
      // sorted intermediate pattern + destination index of every stage-1 product term
      int P1 = 32;
      while (P1 < n1) P1 <<= 1;
      for (int s = n1 + t; s < P1; s += 32) rt.h1k[s] = EMPTY;
      __syncwarp();
      team_bitonic<true>(rt.h1k, P1, 32, t);
      const int ib = a.inter_rowptr[i];
      for (int s = t; s < n1; s += 32) a.inter_col[ib + s] = rt.h1k[s];
      warp_emit_slots(mt_n, a.mt_col + mt_b, a.a_rowptr, a.a_col, a.logG1, rt.h1k, n1, a.slot1 + a.s1_off[i], t, a.slot_row[i] == 2);
      __syncwarp();
    }
    if (!ovf) {
      // stage 2: keys of ((M^T A) M)[i,:]
      accumulate_items<WARP, false, true>(T, t, WARP ? n1 : cap1, rt.h1k, nullptr, a.m_rowptr, a.m_col, nullptr,
                                          a.logG2, rt.h2k, nullptr, mask2, shift2, rt.st_beg, rt.st_len, rt.st_w,
                                          &rt.scratch[1]);
      if (WARP) n2 = warp_compact(rt.h2k, nullptr, cap2, t);
      else n2 = cta_count(rt.h2k, cap2, rt.scratch);
      ovf = (rt.scratch[1] != 0) || (n2 > (cap2 / 4) * 3);
    }
    if (MODE == 0) {
      if (t == 0) {
        if (ovf) {
          a.ovf_rows[atomicAdd(a.n_ovf, 1)] = i;
        } else {
          a.n1[i] = n1;
          a.n2[i] = n2;
          a.level[i] = (signed char)a.this_level;
        }
      }
    } else {
      // sorted column list of row i
      if (ovf) {
        if (t == 0) atomicExch(a.err_flag, 2);
      } else {
        int P;
        if (WARP) {
          P = 32;
          while (P < n2) P <<= 1;
          for (int s = n2 + t; s < P; s += T) rt.h2k[s] = EMPTY;
          team_sync<WARP>();
        } else {
          P = cap2;  // uncompacted table: EMPTY (= INT_MAX) sorts to the end
        }
        team_bitonic<WARP>(rt.h2k, P, T, t);
        const int cb = a.c_rowptr[i];
        if (a.c_rowptr[i + 1] - cb != n2) {
          if (t == 0) atomicExch(a.err_flag, 3);
        } else {
          for (int s = t; s < n2; s += T) a.c_col[cb + s] = rt.h2k[s];
          if (slot_fill) {
            team_sync<WARP>();
            warp_emit_slots(n1, rt.h1k, a.m_rowptr, a.m_col, a.logG2, rt.h2k, n2, a.slot2 + a.s2_off[i], t, a.slot_row[i] == 2);
          }
        }
      }
    }
    team_sync<WARP>();
  }
}

// ------------------------------------------------------------------------------------------------
// numeric kernel
// ------------------------------------------------------------------------------------------------
template <bool WARP>
__global__ void k_ptap_numeric(PtapArgs a) {
  extern __shared__ __align__(16) unsigned char smem[];
  const int T = WARP ? 32 : blockDim.x;
  const int t = WARP ? (threadIdx.x & 31) : threadIdx.x;
  const int team_in_cta = WARP ? (threadIdx.x >> 5) : 0;
  const int teams_per_cta = WARP ? (blockDim.x >> 5) : 1;
  const int64_t team_global = (int64_t)blockIdx.x * teams_per_cta + team_in_cta;
  const int64_t n_teams = (int64_t)gridDim.x * teams_per_cta;
  const int cap1 = 1 << a.log_cap1, cap2 = 1 << a.log_cap2;
  const unsigned mask1 = cap1 - 1, mask2 = cap2 - 1;
  const int shift1 = 32 - a.log_cap1, shift2 = 32 - a.log_cap2;
  RowTables rt = carve<WARP, true>(smem, cap1, cap2, T, team_in_cta, a.g_keys, a.g_vals, (int)team_global);

  for (int64_t wi = team_global; wi < a.n_rows; wi += n_teams) {
    const int i = a.rows[wi];
    const int mt_b = a.mt_rowptr[i], mt_n = a.mt_rowptr[i + 1] - mt_b;
    const int cb = a.c_rowptr[i], n2 = a.c_rowptr[i + 1] - cb;
    for (int s = t; s < cap1; s += T) {
      rt.h1k[s] = EMPTY;
      rt.h1v[s] = 0.0;
    }
    for (int s = t; s < cap2; s += T) {
      rt.h2k[s] = EMPTY;
      rt.h2v[s] = 0.0;
    }
    if (t == 0) rt.scratch[1] = 0;
    team_sync<WARP>();
    // output keys are known from the symbolic phase: stage 2 only looks them up
    for (int s = t; s < n2; s += T)
      if (h_insert(rt.h2k, mask2, shift2, a.c_col[cb + s]) < 0) rt.scratch[1] = 1;
    // stage 1: H1 = sum_j Mt[i,j] A[j,:]
    accumulate_items<WARP, true, true>(T, t, mt_n, a.mt_col + mt_b, a.mt_val + mt_b, a.a_rowptr, a.a_col, a.a_val,
                                       a.logG1, rt.h1k, rt.h1v, mask1, shift1, rt.st_beg, rt.st_len, rt.st_w,
                                       &rt.scratch[1]);
    int n_items = cap1;
    if (WARP) n_items = warp_compact(rt.h1k, rt.h1v, cap1, t);
    // stage 2: H2 = sum_k H1[k] M[k,:]
    accumulate_items<WARP, true, false>(T, t, n_items, rt.h1k, rt.h1v, a.m_rowptr, a.m_col, a.m_val, a.logG2,
                                        rt.h2k, rt.h2v, mask2, shift2, rt.st_beg, rt.st_len, rt.st_w, &rt.scratch[1]);
    for (int s = t; s < n2; s += T) {
      int slot = h_find(rt.h2k, mask2, shift2, a.c_col[cb + s]);
      a.c_val[cb + s] = slot >= 0 ? rt.h2v[slot] : 0.0;
    }
    if (t == 0 && rt.scratch[1]) atomicExch(a.err_flag, 1);
    team_sync<WARP>();
  }
}

}  // namespace iife
#include "ptap_warp.cuh"
#include "ptap_slots.cuh"
#include "ptap_prog.cuh"
#include "ptap_tpl.cuh"
namespace iife {

// ------------------------------------------------------------------------------------------------
// small helper kernels
// ------------------------------------------------------------------------------------------------
__global__ void k_fill_schar(signed char *p, int64_t n, signed char v) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) p[i] = v;
}

// rows that get a slot plan: resolved by a warp-team symbolic level; 1 = both rows <= 256 entries (one byte per
// product term), 2 = wide rows (intermediate row <= 2040, output row <= 512 entries: two bytes per term).  The byte
// counts e1 / e2 of a row are rounded up to even so that every row's plan starts 2-byte aligned.  Non-slot rows get zero
// sizes so that the plan arrays only hold slot rows.
__global__ void k_slot_rows(const int *__restrict__ n1, const int *__restrict__ n2, const signed char *__restrict__ level,
                            int64_t n, int enable, unsigned char *__restrict__ slot_row, const int *__restrict__ t1,
                            const int *__restrict__ t2, int *__restrict__ e1, int *__restrict__ e2, int *__restrict__ n1m) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) {  // t1 / t2: product terms of the two stages (count pass), e1 / e2: bytes of the plan
    const bool warp_row = level[i] >= 0 && level[i] <= 1 && n2[i] > 0;
    const bool s = (enable & 1) && warp_row && n1[i] <= 256 && n2[i] <= 256;
    const bool wide = (enable & 2) && warp_row && !s && n1[i] <= 2040 && n2[i] <= 512 && t1[i] < (1 << 29) && t2[i] < (1 << 29);
    slot_row[i] = s ? 1 : (wide ? 2 : 0);
    e1[i] = s ? ((t1[i] + 1) & ~1) : (wide ? 2 * t1[i] : 0);
    e2[i] = s ? ((t2[i] + 1) & ~1) : (wide ? 2 * t2[i] : 0);
    n1m[i] = (s || wide) ? n1[i] : 0;
  }
}

// numeric level of each row from the exact counts; -1 for empty output rows
__global__ void k_numeric_level(const int *__restrict__ n1, const int *__restrict__ n2, int64_t n,
                                signed char *__restrict__ lvl, int c10, int c20, int c11, int c21, int c12, int c22,
                                int c13, int c23, const unsigned char *__restrict__ slot_row) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) {
    int a = n1[i], b = n2[i];
    signed char l;
    if (b == 0) l = -1;
    else if (slot_row && slot_row[i] == 2) l = 7;
    else if (slot_row && slot_row[i]) l = (a <= 128 && b <= 32) ? 5 : 6;
    else if (2 * a <= c10 && 2 * b <= c20) l = 0;
    else if (2 * a <= c11 && 2 * b <= c21) l = 1;
    else if (2 * a <= c12 && 2 * b <= c22) l = 2;
    else if (2 * a <= c13 && 2 * b <= c23) l = 3;
    else l = 4;
    lvl[i] = l;
  }
}

// (start, length) of the operand row behind every entry of `ids` (Mt columns -> A rows; intermediate
// columns -> M rows); lengths above 255 raise `bad` (the numeric kernel then gathers from rowptr)
__global__ void k_row_meta(const int *__restrict__ ids, int64_t n, const int *__restrict__ x_rowptr,
                           int *__restrict__ beg, unsigned char *__restrict__ len, int *__restrict__ bad) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) {
    int r = ids[i];
    int b = x_rowptr[r], l = x_rowptr[r + 1] - b;
    beg[i] = b;
    len[i] = (unsigned char)(l > 255 ? 255 : l);
    if (l > 255) atomicOr(bad, 1);
  }
}

__global__ void k_level_flag(const signed char *__restrict__ lvl, int64_t n, int which, int *__restrict__ flag) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) flag[i] = (lvl[i] == which);
}

__global__ void k_level_scatter(const signed char *__restrict__ lvl, const int *__restrict__ off, int64_t n, int which,
                                int *__restrict__ out) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride)
    if (lvl[i] == which) out[off[i]] = (int)i;
}

// max and sum of selected counts
__global__ void k_list_max(const int *__restrict__ v, const int *__restrict__ rows, int64_t n, int *__restrict__ out) {
  int m = 0;
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) m = max(m, v[rows ? rows[i] : i]);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = max(m, __shfl_down_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0) atomicMax(out, m);
}

__global__ void k_sum_i32(const int *__restrict__ v, int64_t n, unsigned long long *__restrict__ out) {
  unsigned long long s = 0;
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) s += (unsigned long long)v[i];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0 && s) atomicAdd(out, s);
}

// upper bound of the intermediate row size for listed rows: sum of A row lengths over Mt row
__global__ void k_ub1_list(const int *__restrict__ mt_rowptr, const int *__restrict__ mt_col,
                           const int *__restrict__ a_rowptr, const int *__restrict__ rows, int n,
                           unsigned long long *__restrict__ out_max) {
  for (int r = blockIdx.x; r < n; r += gridDim.x) {
    int i = rows[r];
    unsigned long long s = 0;
    for (int q = mt_rowptr[i] + threadIdx.x; q < mt_rowptr[i + 1]; q += blockDim.x) {
      int j = mt_col[q];
      s += (unsigned long long)(a_rowptr[j + 1] - a_rowptr[j]);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
    __shared__ unsigned long long acc;
    if (threadIdx.x == 0) acc = 0;
    __syncthreads();
    if ((threadIdx.x & 31) == 0) atomicAdd(&acc, s);
    __syncthreads();
    if (threadIdx.x == 0) atomicMax(out_max, acc);
    __syncthreads();
  }
}

static int grid_for(int64_t n, int threads = 256) {
  int64_t g = (n + threads - 1) / threads;
  int64_t cap = (int64_t)ctx().sm_count * 16;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return (int)g;
}

static int log2_ceil(int64_t v) {
  int l = 0;
  while (((int64_t)1 << l) < v) ++l;
  return l;
}

// mean length of the NON-EMPTY rows decides G (M has many empty rows in real data)
__global__ void k_count_nonempty(const int *__restrict__ rowptr, int64_t n, unsigned long long *__restrict__ out) {
  unsigned long long s = 0;
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) s += (rowptr[i + 1] > rowptr[i]);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0 && s) atomicAdd(out, s);
}

static int mean_nonempty_logG(const Mat *X, int *logG) {
  Tmp<unsigned long long> d;
  IIFE_TRY(d.alloc(1));
  IIFE_CUDA(cudaMemsetAsync(d.p, 0, 8, ctx().stream));
  if (X->n_rows) IIFE_LAUNCH(k_count_nonempty, grid_for(X->n_rows), 256, 0, X->rowptr, X->n_rows, d.p);
  IIFE_CHECK_LAUNCH();
  unsigned long long h = 0;
  IIFE_CUDA(cudaMemcpyAsync(&h, d.p, 8, cudaMemcpyDeviceToHost, ctx().stream));
  IIFE_CUDA(cudaStreamSynchronize(ctx().stream));
  double mean = h ? (double)X->nnz / (double)h : 1.0;
  int l = 1;
  while ((1 << l) < mean && l < 5) ++l;
  *logG = l;
  return IIFE_OK;
}

int transpose_build(const Mat *A, Mat **T_out, int **perm_out);  // mat.cu

static void pg_free(Plan *P) {
  if (P->pg_steps) dev_free_t(P->pg_steps, (size_t)P->pg_rows);
  if (P->pg_off) dev_free_t(P->pg_off, (size_t)P->pg_rows + 1);
  if (P->pg_lane_out) dev_free_t(P->pg_lane_out, (size_t)P->pg_rows * 32);
  if (P->pg_maxg) dev_free_t(P->pg_maxg, (size_t)P->pg_rows);
  if (P->pg_slot) dev_free_t(P->pg_slot, (size_t)P->pg_total_steps * 32);
  if (P->pg_val) dev_free_t(P->pg_val, (size_t)P->pg_total_steps * 32);
  P->pg_steps = nullptr;
  P->pg_off = nullptr;
  P->pg_lane_out = P->pg_maxg = P->pg_slot = nullptr;
  P->pg_val = nullptr;
  P->pg_rows = P->pg_total_steps = 0;
  P->pg_state = 0;
}

static void tpl_free(Plan *P) {
  if (P->tp_blobs) dev_free_t(P->tp_blobs, P->tp_blob_bytes);
  if (P->tp_blob_off) dev_free_t(P->tp_blob_off, (size_t)P->tp_n_tpl);
  if (P->tp_chunks) dev_free_t(P->tp_chunks, (size_t)P->tp_n_chunks * 3);
  if (P->tp_rows) dev_free_t(P->tp_rows, (size_t)P->tp_n_rows);
  if (P->tp_rest_rows) dev_free_t(P->tp_rest_rows, (size_t)P->tp_list_n);
  P->tp_blobs = nullptr;
  P->tp_blob_off = nullptr;
  P->tp_chunks = P->tp_rows = P->tp_rest_rows = nullptr;
  P->tp_blob_bytes = 0;
  P->tp_n_tpl = P->tp_n_chunks = 0;
  P->tp_n_rows = P->tp_list_n = P->tp_rest5 = P->tp_rest6 = 0;
  P->tp_state = 0;
}

static int plan_free(Plan *P) {
  if (!P) return IIFE_OK;
  if (P->MT) mat_free(P->MT);
  if (P->mt_perm) dev_free_t(P->mt_perm, (size_t)P->nnzM);
  if (P->c_rowptr) dev_free_t(P->c_rowptr, (size_t)P->n_b + 1);
  if (P->c_colind) dev_free_t(P->c_colind, (size_t)P->nnz_c);
  if (P->n1) dev_free_t(P->n1, (size_t)P->n_b);
  if (P->bin_rows) dev_free_t(P->bin_rows, (size_t)P->n_b);
  if (P->g_keys) dev_free_t(P->g_keys, P->g_keys_n);
  if (P->g_vals) dev_free_t(P->g_vals, P->g_vals_n);
  if (P->err_flag) dev_free_t(P->err_flag, 1);
  if (P->slot1) dev_free_t(P->slot1, (size_t)P->s1_total);
  if (P->slot2) dev_free_t(P->slot2, (size_t)P->s2_total);
  if (P->s1_off) dev_free_t(P->s1_off, (size_t)P->n_b + 1);
  if (P->s2_off) dev_free_t(P->s2_off, (size_t)P->n_b + 1);
  if (P->inter_rowptr) dev_free_t(P->inter_rowptr, (size_t)P->n_b + 1);
  if (P->inter_col) dev_free_t(P->inter_col, (size_t)P->inter_total);
  if (P->mt_abeg) dev_free_t(P->mt_abeg, (size_t)P->nnzR);
  if (P->mt_alen) dev_free_t(P->mt_alen, (size_t)P->nnzR);
  if (P->inter_mbeg) dev_free_t(P->inter_mbeg, (size_t)P->inter_total);
  if (P->inter_mlen) dev_free_t(P->inter_mlen, (size_t)P->inter_total);
  pg_free(P);
  tpl_free(P);
  delete P;
  return IIFE_OK;
}

template <class K>
static int set_smem(K kernel, size_t bytes) {
  if (bytes > 48 * 1024) IIFE_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
  return IIFE_OK;
}

static int read_int(const int *dev, int *host) {
  IIFE_CUDA(cudaMemcpyAsync(host, dev, sizeof(int), cudaMemcpyDeviceToHost, ctx().stream));
  IIFE_CUDA(cudaStreamSynchronize(ctx().stream));
  return IIFE_OK;
}

// launch one symbolic level.  mode 0 count / 1 fill
static int launch_symbolic(const Level &L, int mode, PtapArgs &a, int64_t n_work_host, size_t g_keys_per_team,
                           int *ctas_out) {
  Ctx &c = ctx();
  bool global_tables = (L.log_cap1 == 0);
  int teams_per_cta = L.warp_team ? L.threads / 32 : 1;
  int T = L.warp_team ? 32 : L.threads;
  size_t smem = team_smem_bytes(false, global_tables, 1 << a.log_cap1, 1 << a.log_cap2, T) * teams_per_cta;
  int max_ctas_sm = (int)((size_t)(c.max_smem_optin ? c.max_smem_optin : 227 * 1024) / (smem + 1024));
  if (max_ctas_sm < 1) return set_err(IIFE_ERR_UNSUPPORTED, "symbolic level needs %zu bytes of shared memory", smem);
  int by_threads = 2048 / L.threads;
  if (max_ctas_sm > by_threads) max_ctas_sm = by_threads;
  int64_t ctas = (n_work_host + teams_per_cta - 1) / teams_per_cta;
  int64_t cap = (int64_t)c.sm_count * max_ctas_sm;
  if (ctas > cap) ctas = cap;
  if (ctas < 1) ctas = 1;
  if (global_tables && ctas_out && *ctas_out > 0 && ctas > *ctas_out) ctas = *ctas_out;
  (void)g_keys_per_team;
  if (L.warp_team) {
    if (mode == 0) {
      IIFE_TRY(set_smem(k_ptap_symbolic<true, 0>, smem));
      IIFE_LAUNCH((k_ptap_symbolic<true, 0>), (int)ctas, L.threads, smem, a);
    } else {
      IIFE_TRY(set_smem(k_ptap_symbolic<true, 1>, smem));
      IIFE_LAUNCH((k_ptap_symbolic<true, 1>), (int)ctas, L.threads, smem, a);
    }
  } else {
    if (mode == 0) {
      IIFE_TRY(set_smem(k_ptap_symbolic<false, 0>, smem));
      IIFE_LAUNCH((k_ptap_symbolic<false, 0>), (int)ctas, L.threads, smem, a);
    } else {
      IIFE_TRY(set_smem(k_ptap_symbolic<false, 1>, smem));
      IIFE_LAUNCH((k_ptap_symbolic<false, 1>), (int)ctas, L.threads, smem, a);
    }
  }
  IIFE_CHECK_LAUNCH();
  if (ctas_out) *ctas_out = (int)ctas;
  return IIFE_OK;
}

// R == nullptr: PtAP (the restriction is the transpose of M, built here).  R != nullptr: general triple
// product C = R A M with R: n_out x nJ, A: nJ x nK, M: nK x n_ccols (the row-partitioned path, where a
// rank's block of M^T, the A_f rows it touches and the M rows those touch are gathered first).
static int ptap_symbolic_impl(Mat *R, Mat *M, Mat *A, Plan **out) {
  Ctx &c = ctx();
  if (!R) {
    if (M->n_rows != A->n_rows || A->n_rows != A->n_cols)
      return set_err(IIFE_ERR_ARG, "PtAP shape mismatch: M is %lld x %lld, A is %lld x %lld", (long long)M->n_rows,
                     (long long)M->n_cols, (long long)A->n_rows, (long long)A->n_cols);
  } else if (R->n_cols != A->n_rows || A->n_cols != M->n_rows) {
    return set_err(IIFE_ERR_ARG, "RAP shape mismatch: R %lld x %lld, A %lld x %lld, P %lld x %lld", (long long)R->n_rows,
                   (long long)R->n_cols, (long long)A->n_rows, (long long)A->n_cols, (long long)M->n_rows, (long long)M->n_cols);
  }
  Plan *P = new Plan();
  P->general_rap = (R != nullptr);
  P->n_f = A->n_rows;
  P->n_k = A->n_cols;
  P->n_b = R ? R->n_rows : M->n_cols;
  P->n_ccols = M->n_cols;
  P->nnzM = M->nnz;
  P->nnzA = A->nnz;
  P->nnzR = R ? R->nnz : M->nnz;
  int rc = IIFE_OK;
  const int64_t n_b = P->n_b;
  Tmp<int> n2, ovf_a, ovf_b, n_ovf, flag, off, e1, e2, n1m;
  Tmp<unsigned char> slot_row;
  Tmp<signed char> level, nlevel;
  Tmp<int> sym_keys;  // global tables of the last symbolic level
  Tmp<unsigned long long> u64;
  std::vector<int> level_list_n(N_SYM_LEVELS, 0);
  PlanTimer pt;
  do {
    if ((rc = mat_fingerprint(M, &P->fpM)) != IIFE_OK) break;
    if ((rc = mat_fingerprint(A, &P->fpA)) != IIFE_OK) break;
    pt.lap("fingerprints");
    if (!R) {
      if ((rc = transpose_build(M, &P->MT, &P->mt_perm)) != IIFE_OK) break;
    } else if ((rc = mat_fingerprint(R, &P->fpR)) != IIFE_OK) break;
    Mat *Rm = R ? R : P->MT;
    pt.lap("transpose of M");
    if ((rc = mean_nonempty_logG(A, &P->logG1)) != IIFE_OK) break;
    if ((rc = mean_nonempty_logG(M, &P->logG2)) != IIFE_OK) break;
    if ((rc = dev_alloc_t(&P->n1, (size_t)n_b)) != IIFE_OK) break;
    if ((rc = dev_alloc_t(&P->c_rowptr, (size_t)n_b + 1)) != IIFE_OK) break;
    if ((rc = dev_alloc_t(&P->err_flag, 1)) != IIFE_OK) break;
    if ((rc = n2.alloc((size_t)n_b + 1)) != IIFE_OK) break;
    if ((rc = level.alloc((size_t)n_b)) != IIFE_OK) break;
    if ((rc = ovf_a.alloc((size_t)n_b)) != IIFE_OK) break;
    if ((rc = ovf_b.alloc((size_t)n_b)) != IIFE_OK) break;
    if ((rc = n_ovf.alloc(2)) != IIFE_OK) break;
    if ((rc = u64.alloc(1)) != IIFE_OK) break;
    if ((rc = e1.alloc((size_t)n_b + 1)) != IIFE_OK) break;
    if ((rc = e2.alloc((size_t)n_b + 1)) != IIFE_OK) break;
    if ((rc = n1m.alloc((size_t)n_b + 1)) != IIFE_OK) break;
    if ((rc = slot_row.alloc((size_t)n_b + 1)) != IIFE_OK) break;
    cudaMemsetAsync(e1.p, 0, ((size_t)n_b + 1) * sizeof(int), c.stream);
    cudaMemsetAsync(e2.p, 0, ((size_t)n_b + 1) * sizeof(int), c.stream);
    cudaMemsetAsync(P->err_flag, 0, sizeof(int), c.stream);
    cudaMemsetAsync(n_ovf.p, 0, 2 * sizeof(int), c.stream);
    cudaMemsetAsync(n2.p, 0, ((size_t)n_b + 1) * sizeof(int), c.stream);
    cudaMemsetAsync(P->n1, 0, (size_t)(n_b ? n_b : 1) * sizeof(int), c.stream);
    if (n_b) IIFE_LAUNCH(k_fill_schar, grid_for(n_b), 256, 0, level.p, n_b, (signed char)-1);

    PtapArgs a{};
    a.mt_rowptr = Rm->rowptr;
    a.mt_col = Rm->colind;
    a.mt_val = nullptr;
    a.a_rowptr = A->rowptr;
    a.a_col = A->colind;
    a.a_val = nullptr;
    a.m_rowptr = M->rowptr;
    a.m_col = M->colind;
    a.m_val = nullptr;
    a.n1 = P->n1;
    a.n2 = n2.p;
    a.level = level.p;
    a.logG1 = P->logG1;
    a.logG2 = P->logG2;
    a.err_flag = P->err_flag;
    a.e1 = e1.p;
    a.e2 = e2.p;

    // ---- count pass down the ladder
    int *lists[2] = {ovf_a.p, ovf_b.p};
    int *level_lists[N_SYM_LEVELS] = {nullptr, nullptr, nullptr, nullptr};
    // the overflow list of level l is the work list of level l+1; lists of levels >= 1 must survive
    // until the fill pass, so each gets its own buffer (ovf_a for level 1; levels 2, 3 reuse ovf_b
    // halves — a level-2 list can never be longer than the level-1 list).
    int64_t work = n_b;
    int sym_g_log1 = 0, sym_g_log2 = 0, sym_g_ctas = 0;
    Tmp<int> list2, list3;
    for (int l = 0; l < N_SYM_LEVELS && work > 0; ++l) {
      const Level &L = SYM_LEVELS[l];
      a.this_level = l;
      a.rows = l == 0 ? nullptr : level_lists[l];
      a.n_list = nullptr;
      a.n_rows = work;
      a.g_keys = nullptr;
      a.log_cap1 = L.log_cap1;
      a.log_cap2 = L.log_cap2;
      int ctas = 0;
      if (L.log_cap1 == 0) {
        // global tables: size from an upper bound of the intermediate row and from n_b
        cudaMemsetAsync(u64.p, 0, 8, c.stream);
        IIFE_LAUNCH(k_ub1_list, (int)(work < 1024 ? work : 1024), 256, 0, Rm->rowptr, Rm->colind, A->rowptr, a.rows, (int)work, u64.p);
        unsigned long long ub = 0;
        cudaMemcpyAsync(&ub, u64.p, 8, cudaMemcpyDeviceToHost, c.stream);
        if (cudaStreamSynchronize(c.stream) != cudaSuccess) { rc = set_err(IIFE_ERR_CUDA, "ub1: %s", cudaGetErrorString(cudaGetLastError())); break; }
        int64_t b1 = (int64_t)ub < P->n_k ? (int64_t)ub : P->n_k;
        sym_g_log1 = log2_ceil(2 * b1 + 2);
        sym_g_log2 = log2_ceil(2 * P->n_ccols + 2);
        if (sym_g_log1 > 30 || sym_g_log2 > 30) { rc = set_err(IIFE_ERR_UNSUPPORTED, "PtAP row too large for the global hash level"); break; }
        size_t per_team = ((size_t)1 << sym_g_log1) + ((size_t)1 << sym_g_log2);
        size_t budget = (size_t)4 << 30;  // 4 GiB of table workspace at most
        int64_t teams = (int64_t)(budget / (per_team * 4));
        if (teams < 1) teams = 1;
        if (teams > work) teams = work;
        if (teams > c.sm_count * 2) teams = c.sm_count * 2;
        if ((rc = sym_keys.alloc(per_team * (size_t)teams)) != IIFE_OK) break;
        a.g_keys = sym_keys.p;
        a.log_cap1 = sym_g_log1;
        a.log_cap2 = sym_g_log2;
        sym_g_ctas = (int)teams;
        ctas = sym_g_ctas;
      }
      // overflow destination
      int *dst = nullptr;
      if (l + 1 < N_SYM_LEVELS) {
        if (l == 0) dst = ovf_a.p;
        else if (l == 1) { if ((rc = list2.alloc((size_t)work)) != IIFE_OK) break; dst = list2.p; }
        else { if ((rc = list3.alloc((size_t)work)) != IIFE_OK) break; dst = list3.p; }
      } else {
        dst = lists[1];  // last level cannot overflow by construction; any report is an error
      }
      a.ovf_rows = dst;
      a.n_ovf = n_ovf.p;
      cudaMemsetAsync(n_ovf.p, 0, sizeof(int), c.stream);
      if ((rc = launch_symbolic(L, 0, a, work, 0, &ctas)) != IIFE_OK) break;
      int h_ovf = 0;
      if ((rc = read_int(n_ovf.p, &h_ovf)) != IIFE_OK) break;
      level_list_n[l] = (int)work;
      if (l + 1 < N_SYM_LEVELS) level_lists[l + 1] = dst;
      else if (h_ovf) { rc = set_err(IIFE_ERR_STATE, "PtAP symbolic: %d rows overflowed the global level", h_ovf); break; }
      work = h_ovf;
    }
    if (rc != IIFE_OK) break;
    pt.lap("count pass");

    // ---- row pointers of C
    int64_t total = 0;
    if ((rc = exclusive_scan_i32(n2.p, P->c_rowptr, n_b, &total)) != IIFE_OK) break;
    P->nnz_c = total;
    if ((rc = dev_alloc_t(&P->c_colind, (size_t)total)) != IIFE_OK) break;
    cudaMemsetAsync(u64.p, 0, 8, c.stream);
    if (n_b) IIFE_LAUNCH(k_sum_i32, grid_for(n_b), 256, 0, P->n1, n_b, u64.p);
    {
      unsigned long long s = 0;
      cudaMemcpyAsync(&s, u64.p, 8, cudaMemcpyDeviceToHost, c.stream);
      cudaStreamSynchronize(c.stream);
      P->nnz_inter = (int64_t)s;
    }

    // ---- slot plan sizes (ptap_slots.cuh): which rows, offsets of their product terms
    {
      const bool slots_on = !(getenv("IIFE_PTAP_SLOTS") && atoi(getenv("IIFE_PTAP_SLOTS")) == 0);
      const bool wide_on = !(getenv("IIFE_PTAP_SLOTS_WIDE") && atoi(getenv("IIFE_PTAP_SLOTS_WIDE")) == 0);
      if ((rc = dev_alloc_t(&P->s1_off, (size_t)n_b + 1)) != IIFE_OK) break;
      if ((rc = dev_alloc_t(&P->s2_off, (size_t)n_b + 1)) != IIFE_OK) break;
      if ((rc = dev_alloc_t(&P->inter_rowptr, (size_t)n_b + 1)) != IIFE_OK) break;
      Tmp<int> b1, b2;  // bytes of the plan per row (e1 / e2 keep the term counts of the count pass)
      if ((rc = b1.alloc((size_t)n_b + 1)) != IIFE_OK) break;
      if ((rc = b2.alloc((size_t)n_b + 1)) != IIFE_OK) break;
      int64_t it_total = 0;
      // narrow + wide rows, then narrow rows only, then none (hashing kernels): whatever fits (int32 pattern offsets, memory)
      const int modes[3] = {slots_on ? (wide_on ? 3 : 1) : 0, slots_on ? 1 : 0, 0};
      for (int attempt = 0; attempt < 3; ++attempt) {
        if (attempt > 0 && modes[attempt] == modes[attempt - 1]) continue;
        if (n_b) IIFE_LAUNCH(k_slot_rows, grid_for(n_b), 256, 0, P->n1, n2.p, level.p, n_b, modes[attempt], slot_row.p, e1.p, e2.p, b1.p, b2.p, n1m.p);
        if ((rc = exclusive_scan_i32_i64(b1.p, P->s1_off, n_b, &P->s1_total)) != IIFE_OK) break;
        if ((rc = exclusive_scan_i32_i64(b2.p, P->s2_off, n_b, &P->s2_total)) != IIFE_OK) break;
        rc = exclusive_scan_i32(n1m.p, P->inter_rowptr, n_b, &it_total);
        size_t free_b = 0, total_b = 0;
        cudaMemGetInfo(&free_b, &total_b);
        const int64_t need = P->s1_total + P->s2_total + 4 * it_total;
        const bool fits = rc == IIFE_OK && need <= (int64_t)(free_b / 2);
        if (rc != IIFE_OK && rc != IIFE_ERR_UNSUPPORTED) break;
        if (fits || modes[attempt] == 0) break;
        rc = IIFE_OK;
      }
      if (rc != IIFE_OK) break;
      P->inter_total = it_total;
      if ((rc = dev_alloc_t(&P->slot1, (size_t)P->s1_total)) != IIFE_OK) break;
      if ((rc = dev_alloc_t(&P->slot2, (size_t)P->s2_total)) != IIFE_OK) break;
      if ((rc = dev_alloc_t(&P->inter_col, (size_t)P->inter_total)) != IIFE_OK) break;
      a.slot_row = slot_row.p;
      a.s1_off = P->s1_off;
      a.s2_off = P->s2_off;
      a.slot1 = P->slot1;
      a.slot2 = P->slot2;
      a.inter_rowptr = P->inter_rowptr;
      a.inter_col = P->inter_col;
    }
    pt.lap("scans + slot plan sizes");
    // ---- fill pass: same traversal, same level per row, sorted columns written
    a.c_rowptr = P->c_rowptr;
    a.c_col = P->c_colind;
    for (int l = 0; l < N_SYM_LEVELS; ++l) {
      if (level_list_n[l] == 0) continue;
      const Level &L = SYM_LEVELS[l];
      a.this_level = l;
      a.rows = l == 0 ? nullptr : level_lists[l];
      a.n_list = nullptr;
      a.n_rows = level_list_n[l];
      a.g_keys = nullptr;
      a.log_cap1 = L.log_cap1;
      a.log_cap2 = L.log_cap2;
      int ctas = 0;
      if (L.log_cap1 == 0) {
        a.g_keys = sym_keys.p;
        a.log_cap1 = sym_g_log1;
        a.log_cap2 = sym_g_log2;
        ctas = sym_g_ctas;
      }
      if ((rc = launch_symbolic(L, 1, a, a.n_rows, 0, &ctas)) != IIFE_OK) break;
    }
    if (rc != IIFE_OK) break;
    int h_err = 0;
    if ((rc = read_int(P->err_flag, &h_err)) != IIFE_OK) break;
    if (h_err) { rc = set_err(IIFE_ERR_STATE, "PtAP symbolic fill pass inconsistent with count pass (code %d)", h_err); break; }

    pt.lap("fill pass");
    // ---- packed operand-row metadata for the slot kernel
    if (P->inter_total > 0 && !getenv("IIFE_PTAP_NOMETA")) {
      Mat *Rm2 = R ? R : P->MT;
      if ((rc = dev_alloc_t(&P->mt_abeg, (size_t)P->nnzR)) != IIFE_OK) break;
      if ((rc = dev_alloc_t(&P->mt_alen, (size_t)P->nnzR)) != IIFE_OK) break;
      if ((rc = dev_alloc_t(&P->inter_mbeg, (size_t)P->inter_total)) != IIFE_OK) break;
      if ((rc = dev_alloc_t(&P->inter_mlen, (size_t)P->inter_total)) != IIFE_OK) break;
      cudaMemsetAsync(n_ovf.p, 0, sizeof(int), c.stream);
      IIFE_LAUNCH(k_row_meta, grid_for(P->nnzR), 256, 0, Rm2->colind, P->nnzR, A->rowptr, P->mt_abeg, P->mt_alen, n_ovf.p);
      IIFE_LAUNCH(k_row_meta, grid_for(P->inter_total), 256, 0, P->inter_col, P->inter_total, M->rowptr, P->inter_mbeg,
                  P->inter_mlen, n_ovf.p);
      int bad = 0;
      if ((rc = read_int(n_ovf.p, &bad)) != IIFE_OK) break;
      P->packed_meta = (bad == 0);
    }
    // ---- numeric row bins from the exact counts
    if ((rc = nlevel.alloc((size_t)n_b)) != IIFE_OK) break;
    if ((rc = flag.alloc((size_t)n_b + 1)) != IIFE_OK) break;
    if ((rc = off.alloc((size_t)n_b + 1)) != IIFE_OK) break;
    if ((rc = dev_alloc_t(&P->bin_rows, (size_t)n_b)) != IIFE_OK) break;
    if (n_b)
      IIFE_LAUNCH(k_numeric_level, grid_for(n_b), 256, 0, P->n1, n2.p, n_b, nlevel.p, 1 << NUM_LEVELS[0].log_cap1,
                  1 << NUM_LEVELS[0].log_cap2, 1 << NUM_LEVELS[1].log_cap1, 1 << NUM_LEVELS[1].log_cap2,
                  1 << NUM_LEVELS[2].log_cap1, 1 << NUM_LEVELS[2].log_cap2, 1 << NUM_LEVELS[3].log_cap1,
                  1 << NUM_LEVELS[3].log_cap2, (const unsigned char *)slot_row.p);
    P->bin_off[0] = 0;
    for (int l = 0; l < N_BINS; ++l) {
      int64_t cnt = 0;
      if (n_b) {
        IIFE_LAUNCH(k_level_flag, grid_for(n_b), 256, 0, nlevel.p, n_b, l, flag.p);
        if ((rc = exclusive_scan_i32(flag.p, off.p, n_b, &cnt)) != IIFE_OK) break;
        if (cnt) IIFE_LAUNCH(k_level_scatter, grid_for(n_b), 256, 0, nlevel.p, off.p, n_b, l, P->bin_rows + P->bin_off[l]);
      }
      P->bin_off[l + 1] = P->bin_off[l] + cnt;
    }
    if (rc != IIFE_OK) break;
    // accumulator sizes of the wide slot-plan rows: the longest rows of the bin, not the bin's limits
    {
      const int64_t n_wide = P->bin_off[N_NUM_LEVELS + 3] - P->bin_off[N_NUM_LEVELS + 2];
      P->wide_cap1 = SLOT_CAP1[2];
      P->wide_cap2 = SLOT_CAP2[2];
      if (n_wide > 0) {
        const int *rows = P->bin_rows + P->bin_off[N_NUM_LEVELS + 2];
        int m1 = 0, m2 = 0;
        cudaMemsetAsync(n_ovf.p, 0, 2 * sizeof(int), c.stream);
        IIFE_LAUNCH(k_list_max, grid_for(n_wide), 256, 0, P->n1, rows, n_wide, n_ovf.p);
        IIFE_LAUNCH(k_list_max, grid_for(n_wide), 256, 0, n2.p, rows, n_wide, n_ovf.p + 1);
        if ((rc = read_int(n_ovf.p, &m1)) != IIFE_OK) break;
        if ((rc = read_int(n_ovf.p + 1, &m2)) != IIFE_OK) break;
        P->wide_cap1 = std::min(SLOT_CAP1[2], (m1 + 7) & ~7);
        P->wide_cap2 = std::min(SLOT_CAP2[2], (m2 + 7) & ~7);
      }
    }
    // global tables of the last numeric level
    int64_t n_last = P->bin_off[N_NUM_LEVELS] - P->bin_off[N_NUM_LEVELS - 1];  // bin 4 = global-memory tables
    if (n_last > 0) {
      const int *rows = P->bin_rows + P->bin_off[N_NUM_LEVELS - 1];
      int m1 = 0, m2 = 0;
      cudaMemsetAsync(n_ovf.p, 0, 2 * sizeof(int), c.stream);
      IIFE_LAUNCH(k_list_max, grid_for(n_last), 256, 0, P->n1, rows, n_last, n_ovf.p);
      IIFE_LAUNCH(k_list_max, grid_for(n_last), 256, 0, n2.p, rows, n_last, n_ovf.p + 1);
      if ((rc = read_int(n_ovf.p, &m1)) != IIFE_OK) break;
      if ((rc = read_int(n_ovf.p + 1, &m2)) != IIFE_OK) break;
      P->g_log_cap1 = log2_ceil(2 * (int64_t)m1 + 2);
      P->g_log_cap2 = log2_ceil(2 * (int64_t)m2 + 2);
      size_t per_team = ((size_t)1 << P->g_log_cap1) + ((size_t)1 << P->g_log_cap2);
      size_t budget = (size_t)4 << 30;
      int64_t teams = (int64_t)(budget / (per_team * 12));
      if (teams < 1) teams = 1;
      if (teams > n_last) teams = n_last;
      if (teams > c.sm_count * 2) teams = c.sm_count * 2;
      P->g_ctas = (int)teams;
      P->g_keys_n = per_team * (size_t)teams;
      P->g_vals_n = per_team * (size_t)teams;
      if ((rc = dev_alloc_t(&P->g_keys, P->g_keys_n)) != IIFE_OK) break;
      if ((rc = dev_alloc_t(&P->g_vals, P->g_vals_n)) != IIFE_OK) break;
    }
    cudaError_t e = cudaStreamSynchronize(c.stream);
    if (e != cudaSuccess) rc = set_err(IIFE_ERR_CUDA, "PtAP symbolic: %s", cudaGetErrorString(e));
    pt.lap("metadata + bins");
  } while (0);
  if (rc != IIFE_OK) {
    cudaStreamSynchronize(c.stream);
    plan_free(P);
    return rc;
  }
  *out = P;
  return IIFE_OK;
}

int gather_vals_launch(const double *val, const int *perm, double *out, int64_t nnz);  // mat.cu

// ------------------------------------------------------------------------------------------------
// template plan (ptap_tpl.cuh / ptap_tpl_host.h)
// ------------------------------------------------------------------------------------------------
static int env_int(const char *name, int dflt) {
  const char *e = getenv(name);
  return e ? atoi(e) : dflt;
}

static void tpl_parse_raw(const unsigned char *rec, tpl::Raw &r) {
  const int *hd = (const int *)(rec + TPLR_INTS);
  r.n0 = hd[0];
  r.n1 = hd[1];
  r.n2 = hd[2];
  const int T1 = hd[3], T2 = hd[4];
  r.len1.resize((size_t)r.n0);
  r.w.resize((size_t)r.n0);
  for (int q = 0; q < r.n0; ++q) {
    r.len1[(size_t)q] = rec[TPLR_LEN1 + q];
    r.w[(size_t)q] = ((const double *)(rec + TPLR_W))[q];
  }
  r.slot1.assign(rec + TPLR_SLOT1, rec + TPLR_SLOT1 + T1);
  r.len2.resize((size_t)r.n1);
  for (int q = 0; q < r.n1; ++q) r.len2[(size_t)q] = rec[TPLR_LEN2 + q];
  r.slot2.assign(rec + TPLR_SLOT2, rec + TPLR_SLOT2 + T2);
  r.mval.assign((const double *)(rec + TPLR_MVAL), (const double *)(rec + TPLR_MVAL) + T2);
}

// Groups the rows of the slot-plan bins by (structure, values of R and M), compiles a gather program per group
// on the host and splits the bins into templated rows and the rest.  `a` carries the operands and plan arrays
// with the CURRENT values of R (= M^T) and M.  *rebuilt = true when the row lists changed.
static int tpl_ensure(Plan *P, const Mat *Rm, const Mat *M, PtapArgs a, bool *rebuilt) {
  Ctx &c = ctx();
  *rebuilt = false;
  if (P->tp_state == 1 && P->tp_m_uid == M->uid && P->tp_m_version == M->val_version && P->tp_r_uid == Rm->uid &&
      P->tp_r_version == Rm->val_version)
    return IIFE_OK;
  tpl_free(P);
  *rebuilt = true;
  P->tp_m_uid = M->uid;
  P->tp_m_version = M->val_version;
  P->tp_r_uid = Rm->uid;
  P->tp_r_version = Rm->val_version;
  P->tp_state = 1;
  const int64_t n5 = P->bin_off[N_NUM_LEVELS + 1] - P->bin_off[N_NUM_LEVELS];
  const int64_t n = P->bin_off[N_NUM_LEVELS + 2] - P->bin_off[N_NUM_LEVELS];
  int min_rows = env_int("IIFE_TPL_MIN_ROWS", 32);
  // Small problems are launch-latency bound either way, and compiling programs costs ~0.1 ms of host time each (config 2,
  // 33 k rows of an unstructured mesh: 1 000 groups = 150 ms of cold time for nothing): no templates below
  // IIFE_TPL_MIN_PROBLEM rows, and at most one template per 4 096 rows (at least 32) above it.
  const int64_t min_problem = env_int("IIFE_TPL_MIN_PROBLEM", 32768);
  const int max_tpl = (int)std::max<int64_t>(1, std::min<int64_t>(std::min(env_int("IIFE_TPL_MAX", 1024), 4096),
                                                                std::max<int64_t>(32, n / 4096)));
  if (n < min_rows || n < min_problem || !P->packed_meta || n > 0x7fffffff) return IIFE_OK;
  const int *list = P->bin_rows + P->bin_off[N_NUM_LEVELS];
  a.rows = list;
  a.n_rows = n;
  Tmp<unsigned long long> h1, h2, h1s;
  Tmp<int> pos, spos, head, run_of, run_start, sel_d, n_sel_d;
  IIFE_TRY(h1.alloc((size_t)n));
  IIFE_TRY(h2.alloc((size_t)n));
  IIFE_TRY(h1s.alloc((size_t)n));
  IIFE_TRY(pos.alloc((size_t)n));
  IIFE_TRY(spos.alloc((size_t)n));
  IIFE_TRY(head.alloc((size_t)n + 1));
  IIFE_TRY(run_of.alloc((size_t)n + 1));
  PlanTimer pt;
  const int hgrid = (int)std::min<int64_t>((n + 7) / 8, (int64_t)c.sm_count * 8);
  IIFE_LAUNCH(k_tpl_hash, hgrid, 256, 0, a, h1.p, h2.p);
  IIFE_LAUNCH(k_tpl_iota, grid_for(n), 256, 0, pos.p, (long long)n);
  IIFE_CHECK_LAUNCH();
  {
    size_t tmp_bytes = 0;
    IIFE_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, h1.p, h1s.p, pos.p, spos.p, (int)n, 0, 64, c.stream));
    Tmp<unsigned char> tmp;
    IIFE_TRY(tmp.alloc(tmp_bytes));
    IIFE_CUDA(cub::DeviceRadixSort::SortPairs(tmp.p, tmp_bytes, h1.p, h1s.p, pos.p, spos.p, (int)n, 0, 64, c.stream));
    c.launches += 8;  // the passes of the library sort (plan build only, never in a numeric call)
  }
  pt.lap("tpl: hash + sort");
  IIFE_LAUNCH(k_tpl_heads, grid_for(n), 256, 0, h1s.p, (long long)n, head.p);
  int64_t n_runs = 0;
  IIFE_TRY(exclusive_scan_i32(head.p, run_of.p, n, &n_runs));
  if (n_runs == n) return IIFE_OK;  // nothing repeats
  IIFE_TRY(run_start.alloc((size_t)n_runs + 1));
  IIFE_LAUNCH(k_tpl_run_starts, grid_for(n), 256, 0, head.p, run_of.p, (long long)n, run_start.p);
  constexpr int SEL_CAP = 8192;
  IIFE_TRY(sel_d.alloc(2 * SEL_CAP));
  IIFE_TRY(n_sel_d.alloc(1));
  int n_sel = 0;
  for (;;) {  // more candidates than SEL_CAP: raise the bar (which ones were kept would not be deterministic)
    IIFE_CUDA(cudaMemsetAsync(n_sel_d.p, 0, sizeof(int), c.stream));
    IIFE_LAUNCH(k_tpl_select, grid_for(n_runs), 256, 0, run_start.p, (long long)n_runs, (long long)n, min_rows, SEL_CAP, sel_d.p, n_sel_d.p);
    IIFE_CHECK_LAUNCH();
    IIFE_TRY(read_int(n_sel_d.p, &n_sel));
    if (n_sel <= SEL_CAP) break;
    min_rows *= 4;
  }
  if (n_sel == 0) return IIFE_OK;
  std::vector<int> sel((size_t)2 * n_sel);
  IIFE_CUDA(cudaMemcpyAsync(sel.data(), sel_d.p, sel.size() * sizeof(int), cudaMemcpyDeviceToHost, c.stream));
  IIFE_CUDA(cudaStreamSynchronize(c.stream));
  {
    std::vector<std::pair<int, int>> runs((size_t)n_sel);
    for (int k = 0; k < n_sel; ++k) runs[(size_t)k] = {sel[(size_t)2 * k], sel[(size_t)2 * k + 1]};
    std::sort(runs.begin(), runs.end(), [](const std::pair<int, int> &x, const std::pair<int, int> &y) {
      return x.second != y.second ? x.second > y.second : x.first < y.first;
    });
    if (n_sel > max_tpl) n_sel = max_tpl;
    runs.resize((size_t)n_sel);
    std::sort(runs.begin(), runs.end());  // templated rows are laid out in sorted-hash order
    sel.resize((size_t)2 * n_sel);
    for (int k = 0; k < n_sel; ++k) {
      sel[(size_t)2 * k] = runs[(size_t)k].first;
      sel[(size_t)2 * k + 1] = runs[(size_t)k].second;
    }
  }
  const int n_tpl = n_sel;
  IIFE_CUDA(cudaMemcpyAsync(sel_d.p, sel.data(), sel.size() * sizeof(int), cudaMemcpyHostToDevice, c.stream));
  // ---- representative rows -> host -> programs
  std::vector<unsigned char> raw((size_t)n_tpl * TPLR_STRIDE);
  {
    Tmp<unsigned char> raw_d;
    IIFE_TRY(raw_d.alloc(raw.size()));
    IIFE_LAUNCH(k_tpl_extract, std::min(n_tpl, c.sm_count * 4), 256, 0, a, (const int *)sel_d.p, (const int *)spos.p, n_tpl, raw_d.p);
    IIFE_CHECK_LAUNCH();
    IIFE_CUDA(cudaMemcpyAsync(raw.data(), raw_d.p, raw.size(), cudaMemcpyDeviceToHost, c.stream));
    IIFE_CUDA(cudaStreamSynchronize(c.stream));
  }
  pt.lap("tpl: runs + extract");
  const size_t warp_budget = (size_t)env_int("IIFE_TPL_WARP_SMEM", 20 * 1024);
  std::vector<tpl::Program> progs((size_t)n_tpl);
  std::vector<int> valid((size_t)n_tpl, 0);
  std::vector<long long> blob_off((size_t)n_tpl, 0);
  size_t blob_total = 0;
  int s_cap = 2, o1_cap = 2, o2_cap = 2, n0_max = 1, n_valid = 0;
  for (int t = 0; t < n_tpl; ++t) {
    tpl::Raw r;
    tpl_parse_raw(raw.data() + (size_t)t * TPLR_STRIDE, r);
    tpl::Program &pr = progs[(size_t)t];
    if (!tpl::compile(r, pr)) continue;
    if (((size_t)pr.s_cap + pr.o1_cap + pr.o2_cap) * 8 > warp_budget) continue;  // per row; a warp holds two
    valid[(size_t)t] = 1;
    ++n_valid;
    blob_off[(size_t)t] = (long long)blob_total;
    blob_total += pr.blob.size();
    n0_max = std::max(n0_max, pr.n0);
    s_cap = std::max(s_cap, pr.s_cap);
    o1_cap = std::max(o1_cap, pr.o1_cap);
    o2_cap = std::max(o2_cap, pr.o2_cap);
  }
  pt.lap("tpl: host compile");
  if (pt.on) fprintf(stderr, "[plan] templates: %d candidates, %d compiled\n", n_tpl, n_valid);
  if (n_valid == 0) return IIFE_OK;
  // buffers start on 16-byte boundaries
  s_cap = (s_cap + 1) & ~1;
  o1_cap = (o1_cap + 1) & ~1;
  o2_cap = (o2_cap + 1) & ~1;
  std::vector<unsigned char> blobs(blob_total);
  for (int t = 0; t < n_tpl; ++t)
    if (valid[(size_t)t]) memcpy(blobs.data() + blob_off[(size_t)t], progs[(size_t)t].blob.data(), progs[(size_t)t].blob.size());
  // ---- membership and row lists
  Tmp<int> valid_d, tpl_of, f_sorted, f_rest, off_sorted, off_rest, ranges_d;
  IIFE_TRY(valid_d.alloc((size_t)n_tpl));
  IIFE_TRY(tpl_of.alloc((size_t)n));
  IIFE_TRY(f_sorted.alloc((size_t)n + 1));
  IIFE_TRY(f_rest.alloc((size_t)n + 1));
  IIFE_TRY(off_sorted.alloc((size_t)n + 1));
  IIFE_TRY(off_rest.alloc((size_t)n + 1));
  IIFE_TRY(ranges_d.alloc((size_t)2 * n_tpl));
  IIFE_CUDA(cudaMemcpyAsync(valid_d.p, valid.data(), (size_t)n_tpl * sizeof(int), cudaMemcpyHostToDevice, c.stream));
  IIFE_CUDA(cudaMemsetAsync(tpl_of.p, 0xFF, (size_t)n * sizeof(int), c.stream));
  {
    dim3 g(64, (unsigned)std::min(n_tpl, 65535));
    k_tpl_assign<<<g, 256, 0, c.stream>>>(sel_d.p, valid_d.p, n_tpl, spos.p, h2.p, tpl_of.p);
    c.launches++;
  }
  IIFE_LAUNCH(k_tpl_flags, grid_for(n), 256, 0, tpl_of.p, spos.p, (long long)n, f_sorted.p, f_rest.p);
  IIFE_CHECK_LAUNCH();
  int64_t n_t = 0, n_rest = 0;
  IIFE_TRY(exclusive_scan_i32(f_sorted.p, off_sorted.p, n, &n_t));
  IIFE_TRY(exclusive_scan_i32(f_rest.p, off_rest.p, n, &n_rest));
  if (n_t == 0) return IIFE_OK;
  int rest5 = 0;
  IIFE_TRY(read_int(off_rest.p + n5, &rest5));
  P->tp_n_rows = n_t;
  P->tp_list_n = n;
  P->tp_n_tpl = n_tpl;
  IIFE_TRY(dev_alloc_t(&P->tp_rows, (size_t)n_t));
  IIFE_TRY(dev_alloc_t(&P->tp_rest_rows, (size_t)n));
  IIFE_LAUNCH(k_tpl_scatter, grid_for(n), 256, 0, list, tpl_of.p, spos.p, off_sorted.p, off_rest.p, (long long)n, P->tp_rows, P->tp_rest_rows);
  IIFE_LAUNCH(k_tpl_ranges, (n_tpl + 255) / 256, 256, 0, sel_d.p, n_tpl, off_sorted.p, ranges_d.p);
  IIFE_CHECK_LAUNCH();
  std::vector<int> ranges((size_t)2 * n_tpl);
  IIFE_CUDA(cudaMemcpyAsync(ranges.data(), ranges_d.p, ranges.size() * sizeof(int), cudaMemcpyDeviceToHost, c.stream));
  IIFE_CUDA(cudaStreamSynchronize(c.stream));
  const int chunk_rows = std::max(1, std::min(env_int("IIFE_TPL_CHUNK", TPL_CHUNK), 1024));
  std::vector<int> chunks;
  double u1 = 0.0, u2 = 0.0;
  for (int t = 0; t < n_tpl; ++t) {
    if (!valid[(size_t)t]) continue;
    const int b = ranges[(size_t)2 * t], e = ranges[(size_t)2 * t + 1];
    u1 += progs[(size_t)t].use1 * (e - b);
    u2 += progs[(size_t)t].use2 * (e - b);
    for (int r0 = b; r0 < e; r0 += chunk_rows) {
      chunks.push_back(t);
      chunks.push_back(r0);
      chunks.push_back(std::min(chunk_rows, e - r0));
    }
  }
  P->tp_use1 = u1 / (double)n_t;
  P->tp_use2 = u2 / (double)n_t;
  P->tp_chunk_rows = chunk_rows;
  P->tp_n_chunks = (int)(chunks.size() / 3);
  P->tp_blob_bytes = blob_total;
  IIFE_TRY(dev_alloc_t(&P->tp_chunks, chunks.size()));
  IIFE_TRY(dev_alloc_t(&P->tp_blobs, blob_total));
  IIFE_TRY(dev_alloc_t(&P->tp_blob_off, (size_t)n_tpl));
  IIFE_CUDA(cudaMemcpyAsync(P->tp_chunks, chunks.data(), chunks.size() * sizeof(int), cudaMemcpyHostToDevice, c.stream));
  IIFE_CUDA(cudaMemcpyAsync(P->tp_blobs, blobs.data(), blob_total, cudaMemcpyHostToDevice, c.stream));
  IIFE_CUDA(cudaMemcpyAsync(P->tp_blob_off, blob_off.data(), (size_t)n_tpl * sizeof(long long), cudaMemcpyHostToDevice, c.stream));
  IIFE_CUDA(cudaStreamSynchronize(c.stream));  // the host vectors above go out of scope
  pt.lap("tpl: membership + lists");
  P->tp_rest5 = rest5;
  P->tp_rest6 = n_rest - rest5;
  P->tp_n0_cap = (n0_max + 31) & ~31;
  P->tp_s_cap = s_cap;
  P->tp_o1_cap = o1_cap;
  P->tp_o2_cap = o2_cap;
  return IIFE_OK;
}

static int tpl_launch(Plan *P, PtapArgs a) {
  Ctx &c = ctx();
  TplArgs t{};
  t.blobs = P->tp_blobs;
  t.blob_off = P->tp_blob_off;
  t.chunks = P->tp_chunks;
  t.n_chunks = P->tp_n_chunks;
  t.rows = P->tp_rows;
  t.s_cap = P->tp_s_cap;
  t.o1_cap = P->tp_o1_cap;
  t.o2_cap = P->tp_o2_cap;
  t.n0_cap = P->tp_n0_cap;
  // two rows per warp: sbegA/B + 2 x (S, O1, O2)
  const size_t row_bytes = ((size_t)t.s_cap + t.o1_cap + t.o2_cap) * 8;
  const size_t smem_max = (size_t)(c.max_smem_optin ? c.max_smem_optin : 227 * 1024);
  const size_t per_warp = (size_t)t.n0_cap * 8 + 2 * row_bytes;
  // warps per CTA: whatever keeps the most warps resident per SM (shared memory is the limit; the kernel's latency
  // tolerance is its warp count: 24 resident warps ran 8.2 ms where 16 ran 10.0 ms)
  if (per_warp > smem_max) return set_err(IIFE_ERR_STATE, "template kernel needs %zu B of shared memory per warp", per_warp);
  IIFE_CUDA(cudaFuncSetAttribute(k_ptap_numeric_tpl, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_max));
  int wpc = 1, per_sm = 1, best_warps = 0;
  const int forced = env_int("IIFE_TPL_WPC", 0);
  for (int cand = 8; cand >= 1; --cand) {
    if (forced > 0 && cand != std::min(forced, 8)) continue;
    if (per_warp * cand > smem_max) continue;
    int blocks = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks, k_ptap_numeric_tpl, cand * 32, per_warp * cand) != cudaSuccess || blocks < 1) {
      cudaGetLastError();
      blocks = 1;
    }
    if (blocks * cand > best_warps) {
      best_warps = blocks * cand;
      wpc = cand;
      per_sm = blocks;
    }
  }
  const size_t smem = per_warp * wpc;
  const int64_t ctas = std::min<int64_t>(((int64_t)t.n_chunks + wpc - 1) / wpc, (int64_t)c.sm_count * per_sm);
  k_ptap_numeric_tpl<<<(int)ctas, wpc * 32, smem, c.stream>>>(a, t);
  c.launches++;
  cudaError_t le = cudaGetLastError();
  if (le != cudaSuccess) return set_err(IIFE_ERR_CUDA, "template kernel launch: %s", cudaGetErrorString(le));
  return IIFE_OK;
}

// Builds (or refreshes, when the values of M changed) the stage-2 gather program of the small-row bin.
// `a` carries the operands, the plan arrays and the bin's row list.  pg_state stays -1 when it does not fit.
static int prog_ensure(Plan *P, const Mat *M, PtapArgs a) {
  Ctx &c = ctx();
  if (P->pg_state == -1) return IIFE_OK;
  if (P->pg_state == 1 && P->pg_m_uid == M->uid && P->pg_m_version == M->val_version) return IIFE_OK;
  const int64_t n = a.n_rows;
  const int grid = (int)std::min<int64_t>((n + 7) / 8, (int64_t)c.sm_count * 8);
  ProgArgs pg{};
  if (P->pg_state == 1) pg_free(P);  // values of M changed: rebuild from scratch
  struct Guard {  // an allocation failure below leaves no half-built program behind
    Plan *P;
    bool armed = true;
    ~Guard() {
      if (armed) {
        pg_free(P);
        P->pg_state = -1;
      }
    }
  } guard{P};
  if (P->pg_state == 0) {
    P->pg_rows = n;
    IIFE_TRY(dev_alloc_t(&P->pg_steps, (size_t)n));
    IIFE_TRY(dev_alloc_t(&P->pg_off, (size_t)n + 1));
    pg.steps = P->pg_steps;
    IIFE_LAUNCH(k_prog_build<false>, grid, 256, 0, a, pg);
    IIFE_CHECK_LAUNCH();
    int64_t total = 0;
    IIFE_TRY(exclusive_scan_i32_i64(P->pg_steps, P->pg_off, n, &total));
    size_t free_b = 0, total_b = 0;
    IIFE_CUDA(cudaMemGetInfo(&free_b, &total_b));
    if ((size_t)total * 32 * 9 + ((size_t)n * 33) > free_b / 2) return IIFE_OK;  // guard: freed, state -1 (slot kernel)
    P->pg_total_steps = total;
    IIFE_TRY(dev_alloc_t(&P->pg_slot, (size_t)total * 32));
    IIFE_TRY(dev_alloc_t(&P->pg_val, (size_t)total * 32));
    IIFE_TRY(dev_alloc_t(&P->pg_lane_out, (size_t)n * 32));
    IIFE_TRY(dev_alloc_t(&P->pg_maxg, (size_t)n));
  }
  pg.steps = P->pg_steps;
  pg.off = P->pg_off;
  pg.lane_out = P->pg_lane_out;
  pg.maxg = P->pg_maxg;
  pg.slot = P->pg_slot;
  pg.val = P->pg_val;
  if (P->pg_total_steps) IIFE_CUDA(cudaMemsetAsync(P->pg_slot, 0xFF, (size_t)P->pg_total_steps * 32, c.stream));
  IIFE_LAUNCH(k_prog_build<true>, grid, 256, 0, a, pg);
  IIFE_CHECK_LAUNCH();
  P->pg_m_uid = M->uid;
  P->pg_m_version = M->val_version;
  P->pg_state = 1;
  guard.armed = false;
  return IIFE_OK;
}

static int ptap_numeric_impl(Plan *P, Mat *R, Mat *M, Mat *A, Mat **C_io) {
  Ctx &c = ctx();
  if (M->n_rows != P->n_k || M->n_cols != P->n_ccols || M->nnz != P->nnzM || A->n_rows != P->n_f || A->nnz != P->nnzA)
    return set_err(IIFE_ERR_STATE, "PtAP numeric: operands do not match the symbolic plan (shape/nnz)");
  if (P->general_rap != (R != nullptr) || (R && (R->n_rows != P->n_b || R->nnz != P->nnzR)))
    return set_err(IIFE_ERR_STATE, "PtAP numeric: restriction operand does not match the symbolic plan");
  {  // the slot-plan / template kernels never look at column indices: a different pattern of the same size would give a
     // silently wrong A_b.  The fingerprints are cached per matrix, so this is three 64-bit compares per call.
    uint64_t fM = 0, fA = 0, fR = 0;
    IIFE_TRY(mat_fingerprint(M, &fM));
    IIFE_TRY(mat_fingerprint(A, &fA));
    if (R) IIFE_TRY(mat_fingerprint(R, &fR));
    if (fM != P->fpM || fA != P->fpA || (R && fR != P->fpR))
      return set_err(IIFE_ERR_STATE, "PtAP numeric: the sparsity pattern of an operand differs from the one the plan was built for");
  }
  Mat *C = *C_io;
  bool created = false;
  if (!C) {
    IIFE_TRY(mat_alloc(&C, P->n_b, P->n_ccols, P->nnz_c));
    created = true;
    cudaMemcpyAsync(C->rowptr, P->c_rowptr, ((size_t)P->n_b + 1) * sizeof(int), cudaMemcpyDeviceToDevice, c.stream);
    if (P->nnz_c) cudaMemcpyAsync(C->colind, P->c_colind, (size_t)P->nnz_c * sizeof(int), cudaMemcpyDeviceToDevice, c.stream);
  } else if (C->n_rows != P->n_b || C->nnz != P->nnz_c) {
    return set_err(IIFE_ERR_STATE, "PtAP numeric: result matrix does not match the plan");
  }
  int rc = IIFE_OK;
  do {
    // refresh the values of M^T unless they are already those of this M
    if (!R && (rc = gather_vals_launch(M->val, P->mt_perm, P->MT->val, P->nnzM)) != IIFE_OK) break;
    Mat *Rm = R ? R : P->MT;
    PtapArgs a{};
    a.mt_rowptr = Rm->rowptr;
    a.mt_col = Rm->colind;
    a.mt_val = Rm->val;
    a.a_rowptr = A->rowptr;
    a.a_col = A->colind;
    a.a_val = A->val;
    a.m_rowptr = M->rowptr;
    a.m_col = M->colind;
    a.m_val = M->val;
    a.c_rowptr = P->c_rowptr;
    a.c_col = P->c_colind;
    a.c_val = C->val;
    a.logG1 = P->logG1;
    a.logG2 = P->logG2;
    a.err_flag = P->err_flag;
    a.s1_off = P->s1_off;
    a.s2_off = P->s2_off;
    a.slot1 = P->slot1;
    a.slot2 = P->slot2;
    a.inter_rowptr = P->inter_rowptr;
    a.inter_col = P->inter_col;
    a.mt_abeg = P->packed_meta ? P->mt_abeg : nullptr;
    a.mt_alen = P->packed_meta ? P->mt_alen : nullptr;
    a.inter_mbeg = P->packed_meta ? P->inter_mbeg : nullptr;
    a.inter_mlen = P->packed_meta ? P->inter_mlen : nullptr;
    // template rows first (ptap_tpl.cuh): IIFE_PTAP_TPL=0 keeps every row on the per-row kernels
    bool use_tpl = false;
    {
      const char *et = getenv("IIFE_PTAP_TPL");
      if (!et || atoi(et) != 0) {
        bool rebuilt = false;
        if ((rc = tpl_ensure(P, Rm, M, a, &rebuilt)) != IIFE_OK) break;
        if (rebuilt) pg_free(P);  // the gather program of ptap_prog.cuh is laid out for the old "rest" list
        use_tpl = P->tp_n_rows > 0;
        if (use_tpl && (rc = tpl_launch(P, a)) != IIFE_OK) break;
      } else if (P->tp_state == 1) {
        tpl_free(P);
        pg_free(P);
      }
    }
    for (int sb = 0; sb < N_SLOT_BINS; ++sb) {  // slot-plan rows (ptap_slots.cuh)
      int l = N_NUM_LEVELS + sb;
      const bool wide = sb == 2;
      int64_t cnt = P->bin_off[l + 1] - P->bin_off[l];
      a.rows = P->bin_rows + P->bin_off[l];
      if (use_tpl && !wide) {  // what the templates did not take
        cnt = sb == 0 ? P->tp_rest5 : P->tp_rest6;
        a.rows = P->tp_rest_rows + (sb == 0 ? 0 : P->tp_rest5);
      }
      if (cnt == 0) continue;
      a.n_rows = cnt;
      int lg1 = a.logG1 < 3 ? 3 : (a.logG1 > 5 ? 5 : a.logG1);
      int lg2 = a.logG2 < 2 ? 2 : (a.logG2 > 5 ? 5 : a.logG2);
      if (const char *e1v = getenv("IIFE_PTAP_LG1")) lg1 = atoi(e1v);
      if (const char *e2v = getenv("IIFE_PTAP_LG2")) lg2 = atoi(e2v);
      if (wide) {
        lg1 = lg1 < 4 ? 4 : lg1;
        lg2 = lg2 < 3 ? 3 : lg2;
      }
      if (sb == 0) {  // stage 2 as a per-row gather program (ptap_prog.cuh): small-row bin only, while it fits in memory
        const char *e3 = getenv("IIFE_PTAP_PROG");
        if ((!e3 || atoi(e3) != 0) && SLOT_CAP2[0] <= 32 && SLOT_CAP1[0] < PROG_IDLE) {
          if ((rc = prog_ensure(P, M, a)) != IIFE_OK) break;
          prog_kernel_t pk = P->pg_state == 1 ? pick_prog_kernel(lg1) : nullptr;
          if (pk) {
            ProgArgs pg{};
            pg.steps = P->pg_steps;
            pg.off = P->pg_off;
            pg.lane_out = P->pg_lane_out;
            pg.maxg = P->pg_maxg;
            pg.slot = P->pg_slot;
            pg.val = P->pg_val;
            const int cap1p = SLOT_CAP1[0];
            int wpcp = 8;
            const size_t smemp = prog_per_warp_bytes(lg1, cap1p) * wpcp;
            if (smemp > 48 * 1024) {
              cudaError_t e = cudaFuncSetAttribute(pk, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smemp);
              if (e != cudaSuccess) { rc = set_err(IIFE_ERR_CUDA, "smem attribute: %s", cudaGetErrorString(e)); break; }
            }
            int per_smp = 0;
            if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_smp, pk, wpcp * 32, smemp) != cudaSuccess || per_smp < 1) {
              cudaGetLastError();
              per_smp = 1;
            }
            int64_t ctasp = std::min<int64_t>((cnt + wpcp - 1) / wpcp, (int64_t)c.sm_count * per_smp);
            pk<<<(int)ctasp, wpcp * 32, smemp, c.stream>>>(a, pg, cap1p);
            c.launches++;
            cudaError_t le = cudaGetLastError();
            if (le != cudaSuccess) { rc = set_err(IIFE_ERR_CUDA, "program kernel launch: %s", cudaGetErrorString(le)); break; }
            continue;
          }
        }
      }
      slot_kernel_t kern = pick_slot_kernel(lg1, lg2, wide);
      if (!kern) { rc = set_err(IIFE_ERR_ARG, "no slot kernel for group sizes 2^%d / 2^%d", lg1, lg2); break; }
      int cap1 = wide ? P->wide_cap1 : SLOT_CAP1[sb], cap2 = wide ? P->wide_cap2 : SLOT_CAP2[sb];
      size_t per_warp = slot_per_warp_bytes(lg1, lg2, cap1, cap2);
      size_t smem_max = (size_t)(c.max_smem_optin ? c.max_smem_optin : 227 * 1024);
      int wpc = 8;
      if (const char *ew = getenv("IIFE_PTAP_WPC")) wpc = atoi(ew);
      while (wpc > 1 && per_warp * wpc > smem_max / 2) wpc >>= 1;
      size_t smem = per_warp * wpc;
      if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) { rc = set_err(IIFE_ERR_CUDA, "smem attribute: %s", cudaGetErrorString(e)); break; }
      }
      int per_sm = 0;
      if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, wpc * 32, smem) != cudaSuccess || per_sm < 1) {
        cudaGetLastError();
        per_sm = 1;
      }
      int64_t ctas = (cnt + wpc - 1) / wpc;
      int64_t cap = (int64_t)c.sm_count * per_sm;
      if (ctas > cap) ctas = cap;
      kern<<<(int)ctas, wpc * 32, smem, c.stream>>>(a, cap1, cap2);
      c.launches++;
      cudaError_t le = cudaGetLastError();  // report here: a later occupancy query would clear it
      if (le != cudaSuccess) { rc = set_err(IIFE_ERR_CUDA, "slot-plan kernel launch: %s", cudaGetErrorString(le)); break; }
    }
    if (rc != IIFE_OK) break;
    for (int l = 0; l < N_NUM_LEVELS; ++l) {
      int64_t cnt = P->bin_off[l + 1] - P->bin_off[l];
      if (cnt == 0) continue;
      const Level &L = NUM_LEVELS[l];
      bool global_tables = (L.log_cap1 == 0);
      a.rows = P->bin_rows + P->bin_off[l];
      a.n_rows = cnt;
      a.log_cap1 = global_tables ? P->g_log_cap1 : L.log_cap1;
      a.log_cap2 = global_tables ? P->g_log_cap2 : L.log_cap2;
      a.g_keys = global_tables ? P->g_keys : nullptr;
      a.g_vals = global_tables ? P->g_vals : nullptr;
      if (L.warp_team) {
        // privatised-table warp kernel (ptap_warp.cuh)
        int lg1 = a.logG1 < 3 ? 3 : (a.logG1 > 5 ? 5 : a.logG1);
        int lg2 = a.logG2 < 2 ? 2 : (a.logG2 > 5 ? 5 : a.logG2);
        if (const char *e1 = getenv("IIFE_PTAP_LG1")) lg1 = atoi(e1);
        if (const char *e2 = getenv("IIFE_PTAP_LG2")) lg2 = atoi(e2);
        warp_kernel_t kern = pick_warp_kernel(lg1, lg2);
        if (!kern) { rc = set_err(IIFE_ERR_ARG, "no warp kernel for group sizes 2^%d / 2^%d", lg1, lg2); break; }
        size_t per_warp = warp_kernel_smem_per_warp(lg1, lg2, a.log_cap1, a.log_cap2);
        size_t smem_max = (size_t)(c.max_smem_optin ? c.max_smem_optin : 227 * 1024);
        int wpc = 4;
        if (const char *ew = getenv("IIFE_PTAP_WPC")) wpc = atoi(ew);
        while (wpc > 1 && per_warp * wpc > smem_max / 2) wpc >>= 1;
        size_t smem = per_warp * wpc;
        if (smem > smem_max) { rc = set_err(IIFE_ERR_UNSUPPORTED, "numeric level %d needs %zu B of shared memory", l, smem); break; }
        int ctas_sm = (int)((smem_max + 1024) / (smem + 1024));
        if (ctas_sm * wpc > 64) ctas_sm = 64 / wpc;
        if (ctas_sm > 32) ctas_sm = 32;
        int64_t ctas = (cnt + wpc - 1) / wpc;
        int64_t cap = (int64_t)c.sm_count * ctas_sm;
        if (ctas > cap) ctas = cap;
        if (smem > 48 * 1024) {
          cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
          if (e != cudaSuccess) { rc = set_err(IIFE_ERR_CUDA, "smem attribute: %s", cudaGetErrorString(e)); break; }
        }
        kern<<<(int)ctas, wpc * 32, smem, c.stream>>>(a);
        c.launches++;
        continue;
      }
      int teams_per_cta = L.warp_team ? L.threads / 32 : 1;
      int T = L.warp_team ? 32 : L.threads;
      size_t smem = team_smem_bytes(true, global_tables, 1 << a.log_cap1, 1 << a.log_cap2, T) * teams_per_cta;
      int max_ctas_sm = (int)((size_t)(c.max_smem_optin ? c.max_smem_optin : 227 * 1024) / (smem + 1024));
      if (max_ctas_sm < 1) { rc = set_err(IIFE_ERR_UNSUPPORTED, "numeric level %d needs %zu B of shared memory", l, smem); break; }
      int by_threads = 2048 / L.threads;
      if (max_ctas_sm > by_threads) max_ctas_sm = by_threads;
      int64_t ctas = (cnt + teams_per_cta - 1) / teams_per_cta;
      int64_t cap = (int64_t)c.sm_count * max_ctas_sm;
      if (ctas > cap) ctas = cap;
      if (global_tables && ctas > P->g_ctas) ctas = P->g_ctas;
      if ((rc = set_smem(k_ptap_numeric<false>, smem)) != IIFE_OK) break;
      IIFE_LAUNCH(k_ptap_numeric<false>, (int)ctas, L.threads, smem, a);
    }
    if (rc != IIFE_OK) break;
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { rc = set_err(IIFE_ERR_CUDA, "PtAP numeric launch: %s", cudaGetErrorString(e)); break; }
    C->T_vals_valid = false;
    C->dinv_valid = false;
    C->sell_vals_valid = false;
    C->val_version++;
  } while (0);
  if (rc != IIFE_OK) {
    if (created) mat_free(C);
    return rc;
  }
  *C_io = C;
  return IIFE_OK;
}

// LRU cache of plans for the handle-less AT_R_A entry point
struct CacheEntry {
  uint64_t fpM, fpA;
  Plan *plan;
};
static std::list<CacheEntry> g_cache;
constexpr size_t CACHE_MAX = 8;

// a plan of the 50 M-DOF case holds ~12 GB (transpose, patterns, slot plan): before building another one,
// drop least-recently-used plans while less than a quarter of the device memory is free
static void cache_make_room() {
  size_t free_b = 0, total_b = 0;
  while (!g_cache.empty() && cudaMemGetInfo(&free_b, &total_b) == cudaSuccess && free_b < total_b / 4) {
    cudaStreamSynchronize(ctx().stream);
    plan_free(g_cache.back().plan);
    g_cache.pop_back();
    dev_release_cached();
  }
}

}  // namespace iife

using namespace iife;

extern "C" {

int iife_ptap_symbolic(iife_mat M_, iife_mat A_, iife_plan *out) {
  IIFE_NEED_INIT();
  if (!M_ || !A_ || !out) return set_err(IIFE_ERR_ARG, "NULL argument");
  *out = nullptr;
  Plan *P = nullptr;
  IIFE_TRY(ptap_symbolic_impl(nullptr, (Mat *)M_, (Mat *)A_, &P));
  *out = (iife_plan)P;
  return IIFE_OK;
}

int iife_plan_matches(iife_plan P_, iife_mat M_, iife_mat A_, int *matches) {
  IIFE_NEED_INIT();
  Plan *P = (Plan *)P_;
  if (!P || !M_ || !A_ || !matches) return set_err(IIFE_ERR_ARG, "NULL argument");
  uint64_t fm = 0, fa = 0;
  IIFE_TRY(mat_fingerprint((Mat *)M_, &fm));
  IIFE_TRY(mat_fingerprint((Mat *)A_, &fa));
  *matches = (fm == P->fpM && fa == P->fpA) ? 1 : 0;
  return IIFE_OK;
}

int iife_plan_get_info(iife_plan P_, int64_t *n_b, int64_t *nnz_c, int64_t *nnz_inter) {
  Plan *P = (Plan *)P_;
  if (!P) return set_err(IIFE_ERR_ARG, "NULL plan");
  if (n_b) *n_b = P->n_b;
  if (nnz_c) *nnz_c = P->nnz_c;
  if (nnz_inter) *nnz_inter = P->nnz_inter;
  return IIFE_OK;
}

int iife_ptap_numeric(iife_plan P_, iife_mat M_, iife_mat A_, iife_mat *C) {
  IIFE_NEED_INIT();
  if (!P_ || !M_ || !A_ || !C) return set_err(IIFE_ERR_ARG, "NULL argument");
  Mat *Cm = (Mat *)*C;
  IIFE_TRY(ptap_numeric_impl((Plan *)P_, nullptr, (Mat *)M_, (Mat *)A_, &Cm));
  *C = (iife_mat)Cm;
  return IIFE_OK;
}

int iife_rap_symbolic(iife_mat R_, iife_mat A_, iife_mat P_, iife_plan *out) {
  IIFE_NEED_INIT();
  if (!R_ || !A_ || !P_ || !out) return set_err(IIFE_ERR_ARG, "NULL argument");
  *out = nullptr;
  Plan *P = nullptr;
  IIFE_TRY(ptap_symbolic_impl((Mat *)R_, (Mat *)P_, (Mat *)A_, &P));
  *out = (iife_plan)P;
  return IIFE_OK;
}

int iife_rap_numeric(iife_plan plan_, iife_mat R_, iife_mat A_, iife_mat P_, iife_mat *C) {
  IIFE_NEED_INIT();
  if (!plan_ || !R_ || !A_ || !P_ || !C) return set_err(IIFE_ERR_ARG, "NULL argument");
  Mat *Cm = (Mat *)*C;
  IIFE_TRY(ptap_numeric_impl((Plan *)plan_, (Mat *)R_, (Mat *)P_, (Mat *)A_, &Cm));
  *C = (iife_mat)Cm;
  return IIFE_OK;
}

int iife_plan_bin_counts(iife_plan P_, int64_t *counts7) { return iife_plan_bin_counts_n(P_, counts7, 7); }

int iife_plan_bin_counts_n(iife_plan P_, int64_t *counts, int n) {
  Plan *P = (Plan *)P_;
  if (!P || !counts || n < 0) return set_err(IIFE_ERR_ARG, "NULL argument");
  for (int l = 0; l < n; ++l) counts[l] = l < N_BINS ? P->bin_off[l + 1] - P->bin_off[l] : 0;
  return IIFE_OK;
}

int iife_plan_tpl_info(iife_plan P_, int64_t *n_templates, int64_t *n_rows, int64_t *n_chunks, double *lane_use2) {
  Plan *P = (Plan *)P_;
  if (!P) return set_err(IIFE_ERR_ARG, "NULL plan");
  int64_t nt = 0;
  if (P->tp_state == 1 && P->tp_n_rows > 0) nt = P->tp_n_tpl;
  if (n_templates) *n_templates = nt;
  if (n_rows) *n_rows = P->tp_state == 1 ? P->tp_n_rows : 0;
  if (n_chunks) *n_chunks = P->tp_state == 1 ? P->tp_n_chunks : 0;
  if (lane_use2) {
    lane_use2[0] = P->tp_use1;
    lane_use2[1] = P->tp_use2;
  }
  return IIFE_OK;
}

// Host-only: compiles the gather program of one row description and runs the CPU interpreter on it (the CUDA
// kernel executes the same program).  No device needed; used by the CPU tests of the program compiler.
int iife_tpl_emulate_row(int n0, const int *len1, const double *w, const unsigned char *slot1, int n1, const int *len2,
                         const double *mval, const unsigned char *slot2, int n2, const double *a_vals, double *c_out,
                         int *info10) {
  if (!len1 || !w || !slot1 || !len2 || !mval || !slot2 || !a_vals || !c_out) return set_err(IIFE_ERR_ARG, "NULL argument");
  tpl::Raw r;
  r.n0 = n0;
  r.n1 = n1;
  r.n2 = n2;
  if (n0 < 0 || n1 < 0 || n2 < 0) return set_err(IIFE_ERR_ARG, "negative size");
  r.len1.assign(len1, len1 + n0);
  r.w.assign(w, w + n0);
  int T1 = 0, T2 = 0;
  for (int q = 0; q < n0; ++q) T1 += len1[q];
  r.len2.assign(len2, len2 + n1);
  for (int q = 0; q < n1; ++q) T2 += len2[q];
  r.slot1.assign(slot1, slot1 + T1);
  r.slot2.assign(slot2, slot2 + T2);
  r.mval.assign(mval, mval + T2);
  tpl::Program pr;
  if (!tpl::compile(r, pr)) return set_err(IIFE_ERR_UNSUPPORTED, "row does not fit the template kernel's limits");
  std::vector<const double *> rows((size_t)n0);
  {
    const double *p = a_vals;
    for (int q = 0; q < n0; ++q) {
      rows[(size_t)q] = p;
      p += len1[q];
    }
  }
  tpl::interpret(pr.blob.data(), rows.data(), c_out);
  if (info10) {
    tpl::Header h;
    memcpy(&h, pr.blob.data(), sizeof(h));
    info10[0] = h.S1;
    info10[1] = h.S2;
    info10[2] = h.nx1;
    info10[3] = h.nx2;
    info10[4] = h.stg_steps;
    info10[5] = h.blob_bytes;
    info10[6] = (int)(pr.use1 * 1000.0);
    info10[7] = (int)(pr.use2 * 1000.0);
    info10[8] = (int)(pr.conf1 * 1000.0);
    info10[9] = (int)(pr.conf2 * 1000.0);
  }
  return IIFE_OK;
}

int iife_plan_check(iife_plan P_) {
  IIFE_NEED_INIT();
  Plan *P = (Plan *)P_;
  if (!P) return set_err(IIFE_ERR_ARG, "NULL plan");
  int h = 0;
  IIFE_CUDA(cudaMemcpyAsync(&h, P->err_flag, sizeof(int), cudaMemcpyDeviceToHost, ctx().stream));
  IIFE_CUDA(cudaStreamSynchronize(ctx().stream));
  if (h) {
    cudaMemsetAsync(P->err_flag, 0, sizeof(int), ctx().stream);
    return set_err(IIFE_ERR_STATE, "PtAP numeric: a product term had no slot in the plan's pattern (code %d): operands do not match the plan", h);
  }
  return IIFE_OK;
}

int iife_plan_destroy(iife_plan P_) {
  Plan *P = (Plan *)P_;
  if (!P) return IIFE_OK;
  if (ctx().init) cudaStreamSynchronize(ctx().stream);
  for (auto it = g_cache.begin(); it != g_cache.end(); ++it)
    if (it->plan == P) {
      g_cache.erase(it);
      break;
    }
  return plan_free(P);
}

int iife_ptap(iife_mat M_, iife_mat A_, iife_mat *C, int *plan_was_cached) {
  IIFE_NEED_INIT();
  if (!M_ || !A_ || !C) return set_err(IIFE_ERR_ARG, "NULL argument");
  *C = nullptr;
  Mat *M = (Mat *)M_, *A = (Mat *)A_;
  uint64_t fm = 0, fa = 0;
  IIFE_TRY(mat_fingerprint(M, &fm));
  IIFE_TRY(mat_fingerprint(A, &fa));
  Plan *P = nullptr;
  for (auto it = g_cache.begin(); it != g_cache.end(); ++it)
    if (it->fpM == fm && it->fpA == fa) {
      P = it->plan;
      g_cache.splice(g_cache.begin(), g_cache, it);
      break;
    }
  if (plan_was_cached) *plan_was_cached = P ? 1 : 0;
  if (!P) {
    cache_make_room();
    IIFE_TRY(ptap_symbolic_impl(nullptr, M, A, &P));
    g_cache.push_front({fm, fa, P});
    while (g_cache.size() > CACHE_MAX) {
      plan_free(g_cache.back().plan);
      g_cache.pop_back();
    }
  }
  Mat *Cm = nullptr;
  IIFE_TRY(ptap_numeric_impl(P, nullptr, M, A, &Cm));
  *C = (iife_mat)Cm;
  return IIFE_OK;
}

int iife_plan_cache_clear(void) {
  if (ctx().init) cudaStreamSynchronize(ctx().stream);
  for (auto &e : g_cache) plan_free(e.plan);
  g_cache.clear();
  return IIFE_OK;
}

}  // extern "C"
