// spmv.cu — CSR sparse matrix-vector products (replaces PETSc MatMult / MatMultTranspose /
// MatMultAdd reached from reference la_utils.py:141,162 and common.py:139,364, and the SpMV inside
// KSPSolve, common.py:636).
//
// Layout: plain CSR, int32 indices, fp64 values.  LPR lanes (a power-of-two slice of a warp) own one
// row: consecutive lanes read consecutive (colind, val) entries — coalesced 4- and 8-byte streams
// marked evict-first (each matrix byte is touched once per product) — gather x through the read-only
// path (x is the only operand with reuse; it stays L2-resident) and combine with xor-shuffles.
// A grid of SMs x resident CTAs walks the rows grid-stride, so partial dot products in the fused
// variant need a bounded scratch array and are reduced in a fixed order by the last CTA to finish
// (bit-reproducible for a given launch geometry).
#include "common.cuh"
#include "p2p_dev.cuh"
#include <stdlib.h>
#include <algorithm>

namespace iife {

constexpr int SPMV_THREADS = 256;

// Register budget of the SELL kernels: minBlocks = 4 gives them 64 registers.  ptxas' own choice (40) leaves the
// dot-fused variant with one load in flight where the plain one has U (SASS: 3.3 against 6.1 pending load registers
// on average; 12.5 for both with 64 registers).  Measured at N_b=184 (gpurun_out r2 A/B, profiles/r02_ab.md):
// plain SpMV 293 -> 280 us, CG iteration 421 -> 407 us, and 395 us with the dot fused (which was a loss at 40).
#ifndef IIFE_SELL_MINBLOCKS
#define IIFE_SELL_MINBLOCKS 4
#endif
#if IIFE_SELL_MINBLOCKS > 0
#define IIFE_SELL_BOUNDS __launch_bounds__(SPMV_THREADS, IIFE_SELL_MINBLOCKS)
#else
#define IIFE_SELL_BOUNDS __launch_bounds__(SPMV_THREADS)
#endif

__device__ __forceinline__ double ld_stream(const double *p) { return __ldcs(p); }
__device__ __forceinline__ int ld_stream(const int *p) { return __ldcs(p); }

template <int LPR>
__device__ __forceinline__ double row_dot(const int *__restrict__ colind, const double *__restrict__ val,
                                          const double *__restrict__ x, int b, int e, int lg) {
  double s = 0.0;
  int p = b + lg;
  // two entries per lane per trip for memory-level parallelism
  for (; p + LPR < e; p += 2 * LPR) {
    int c0 = ld_stream(colind + p), c1 = ld_stream(colind + p + LPR);
    double v0 = ld_stream(val + p), v1 = ld_stream(val + p + LPR);
    s = fma(v0, __ldg(x + c0), s);
    s = fma(v1, __ldg(x + c1), s);
  }
  if (p < e) s = fma(ld_stream(val + p), __ldg(x + ld_stream(colind + p)), s);
  return group_sum<LPR>(s);
}

template <int LPR, bool PLAIN>
__global__ void __launch_bounds__(SPMV_THREADS)
k_spmv(const int *__restrict__ rowptr, const int *__restrict__ colind, const double *__restrict__ val,
       int64_t n_rows, double alpha, const double *__restrict__ x, double beta, double *__restrict__ y) {
  constexpr int RPB = SPMV_THREADS / LPR;
  int lg = threadIdx.x % LPR;
  int64_t row0 = (int64_t)blockIdx.x * RPB + threadIdx.x / LPR;
  int64_t stride = (int64_t)gridDim.x * RPB;
  int64_t n_iter = (n_rows + stride - 1) / stride;  // uniform trip count: shuffles use the full mask
  for (int64_t it = 0; it < n_iter; ++it) {
    int64_t i = row0 + it * stride;
    int b = 0, e = 0;
    if (i < n_rows) {
      b = __ldg(rowptr + i);
      e = __ldg(rowptr + i + 1);
    }
    double s = row_dot<LPR>(colind, val, x, b, e, lg);
    if (i < n_rows && lg == 0) {
      if (PLAIN) y[i] = s;
      else y[i] = (beta == 0.0) ? alpha * s : fma(alpha, s, beta * y[i]);
    }
  }
}

// y = A x for operators with very short rows (the extraction operator M: 1-8 entries per row).  k_spmv keeps one
// row per lane group in flight, and a row is a chain of three dependent loads (row pointer -> column -> x):
// measured 0.45 of the copy peak on M.  Here a lane group walks TWO rows at a time with their loads interleaved,
// which doubles the bytes in flight per SM: 0.915 -> 0.771 ms on M at N_b=184.  IIFE_SPMV_ILP=0 disables it.
template <int LPR>
__global__ void __launch_bounds__(SPMV_THREADS)
k_spmv_ilp2(const int *__restrict__ rowptr, const int *__restrict__ colind, const double *__restrict__ val,
            int64_t n_rows, const double *__restrict__ x, double *__restrict__ y) {
  constexpr int RPB = SPMV_THREADS / LPR;
  const int lg = threadIdx.x % LPR;
  const int64_t row0 = (int64_t)blockIdx.x * RPB + threadIdx.x / LPR;
  const int64_t stride = (int64_t)gridDim.x * RPB;
  const int64_t n_iter = (n_rows + 2 * stride - 1) / (2 * stride);  // uniform trip count: shuffles use the full mask
  for (int64_t it = 0; it < n_iter; ++it) {
    const int64_t i0 = row0 + 2 * it * stride, i1 = i0 + stride;
    int b0 = 0, e0 = 0, b1 = 0, e1 = 0;
    if (i0 < n_rows) {
      b0 = __ldg(rowptr + i0);
      e0 = __ldg(rowptr + i0 + 1);
    }
    if (i1 < n_rows) {
      b1 = __ldg(rowptr + i1);
      e1 = __ldg(rowptr + i1 + 1);
    }
    double s0 = 0.0, s1 = 0.0;
    int p0 = b0 + lg, p1 = b1 + lg;
    while (__any_sync(0xffffffffu, p0 < e0 || p1 < e1)) {
      int c0 = 0, c1 = 0;
      double v0 = 0.0, v1 = 0.0;
      if (p0 < e0) {
        c0 = ld_stream(colind + p0);
        v0 = ld_stream(val + p0);
      }
      if (p1 < e1) {
        c1 = ld_stream(colind + p1);
        v1 = ld_stream(val + p1);
      }
      if (p0 < e0) s0 = fma(v0, __ldg(x + c0), s0);
      if (p1 < e1) s1 = fma(v1, __ldg(x + c1), s1);
      p0 += LPR;
      p1 += LPR;
    }
    s0 = group_sum<LPR>(s0);
    s1 = group_sum<LPR>(s1);
    if (lg == 0) {
      if (i0 < n_rows) y[i0] = s0;
      if (i1 < n_rows) y[i1] = s1;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// CSR-stream SpMV for operators with SHORT, ragged rows (the extraction operator M: 1-8 entries per row, rejected by
// SELL-32 for its padding): y = A x, transferToForeground (reference common.py:123-140, la_utils.py:129-141).
// The (colind, val) entries of a tile of consecutive rows are one contiguous range of the CSR arrays, so a tile is
// staged in shared memory with two bulk async copies (cp.async.bulk = the TMA engine, completion on an mbarrier):
// no register staging, no dependent pointer chase, STAGES tiles in flight per CTA, and the global reads are perfectly
// coalesced whatever the row lengths.  One thread then owns one row and walks its entries in shared memory; the only
// irregular access left is the gather of x (L1/L2 resident: consecutive foreground rows touch neighbouring
// background columns).  Persistent CTAs, tiles handed out round robin.
// ------------------------------------------------------------------------------------------------
constexpr int STREAM_ROWS = 128;  // rows per tile = threads per CTA
constexpr int STREAM_MAX_STAGES = 4;

__device__ __forceinline__ void mbar_init(unsigned bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(unsigned bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned phase) {
  unsigned done = 0;
  while (!done) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(done)
        : "r"(bar), "r"(phase)
        : "memory");
  }
}
__device__ __forceinline__ void bulk_g2s(unsigned dst, const void *src, unsigned bytes, unsigned bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src),
               "r"(bytes), "r"(bar)
               : "memory");
}

// Shared memory per stage: the tile's row pointers (STREAM_ROWS + 4 words), cap_entries + 4 column words and
// cap_entries + 2 values (alignment slack: bulk copies start on 16-byte boundaries of the global arrays).  cap_entries =
// the largest number of entries any tile of STREAM_ROWS consecutive rows holds (Mat::max_tile_entries), so the stages
// are sized by what the operator needs, not by rows x longest row: more CTAs per SM.  U = gathers of x in flight per
// thread and round.
template <int U, int R>
__global__ void __launch_bounds__(R)
k_spmv_stream(const int *__restrict__ rowptr, const int *__restrict__ colind, const double *__restrict__ val, int64_t n_rows,
              int64_t nnz, int cap_entries, int stages, const double *__restrict__ x, double *__restrict__ y) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ __align__(8) unsigned long long bars[STREAM_MAX_STAGES];
  __shared__ int s_base[STREAM_MAX_STAGES];  // first entry (rowptr[r0]) of the tile in each stage; < 0: tile read from global
  constexpr size_t STREAM_RP_BYTES = (R + 4) * 4;
  const int tid = threadIdx.x;
  const size_t val_bytes = ((size_t)cap_entries + 2) * 8, col_bytes = (((size_t)cap_entries + 4) * 4 + 15) & ~(size_t)15;
  const size_t stage_bytes = val_bytes + col_bytes + STREAM_RP_BYTES;
  const int64_t n_tiles = (n_rows + R - 1) / R;
  if (tid == 0) {
    for (int s = 0; s < stages; ++s) mbar_init((unsigned)__cvta_generic_to_shared(&bars[s]), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  // producer: thread 0 issues the copies of tile `t` into stage `s`
  auto issue = [&](int64_t t, int s) {
    const int64_t r0 = t * R;
    const int64_t r1 = r0 + R < n_rows ? r0 + R : n_rows;
    const int b = __ldg(rowptr + r0), e = __ldg(rowptr + r1);
    const int bc = b & ~3, bv = b & ~1;
    const unsigned cb = (unsigned)(((e - bc) * 4 + 15) & ~15), vb = (unsigned)(((e - bv) * 8 + 15) & ~15);
    const unsigned bar = (unsigned)__cvta_generic_to_shared(&bars[s]);
    // the rounded-up copies may not run past the arrays (last tile): such a tile is read with plain loads instead
    const bool fits = (int64_t)bc * 4 + cb <= nnz * 4 && (int64_t)bv * 8 + vb <= nnz * 8 && (e - b) <= cap_entries &&
                      r0 + R + 4 <= n_rows + 1;
    if (e == b || !fits) {
      s_base[s] = -1 - b;  // nothing staged (empty tile, or tail / oversized tile)
      mbar_expect_tx(bar, 0);
      return;
    }
    s_base[s] = b;
    unsigned char *st = smem + (size_t)s * stage_bytes;
    mbar_expect_tx(bar, cb + vb + (unsigned)STREAM_RP_BYTES);
    bulk_g2s((unsigned)__cvta_generic_to_shared(st), val + bv, vb, bar);
    bulk_g2s((unsigned)__cvta_generic_to_shared(st + val_bytes), colind + bc, cb, bar);
    bulk_g2s((unsigned)__cvta_generic_to_shared(st + val_bytes + col_bytes), rowptr + r0, (unsigned)STREAM_RP_BYTES, bar);
  };
  const int64_t first = blockIdx.x, step = gridDim.x;
  if (tid == 0)
    for (int s = 0; s < stages; ++s)
      if (first + (int64_t)s * step < n_tiles) issue(first + (int64_t)s * step, s);
  int s = 0;
  unsigned phase = 0;
  for (int64_t t = first; t < n_tiles; t += step) {
    mbar_wait((unsigned)__cvta_generic_to_shared(&bars[s]), phase);
    const int base = s_base[s];
    const int64_t i = t * R + tid;
    if (i < n_rows) {
      double acc = 0.0;
      if (base >= 0) {
        const unsigned char *st = smem + (size_t)s * stage_bytes;
        const double *sv = (const double *)st - (base & ~1);
        const int *sc = (const int *)(st + val_bytes) - (base & ~3);
        const int *srp = (const int *)(st + val_bytes + col_bytes);
        const int rb = srp[tid], re = srp[tid + 1];
        // U gathers of x in flight per thread (rows are short: a plain loop would leave one)
        for (int p = rb; p < re; p += U) {
          double xv[U], av[U];
#pragma unroll
          for (int u = 0; u < U; ++u) {
            const bool in = p + u < re;
            av[u] = in ? sv[p + u] : 0.0;
            xv[u] = in ? __ldg(x + sc[p + u]) : 0.0;
          }
#pragma unroll
          for (int u = 0; u < U; ++u) acc = fma(av[u], xv[u], acc);
        }
      } else {
        const int rb = __ldg(rowptr + i), re = __ldg(rowptr + i + 1);
        for (int p = rb; p < re; ++p) acc = fma(ld_stream(val + p), __ldg(x + ld_stream(colind + p)), acc);
      }
      y[i] = acc;
    }
    __syncthreads();  // every thread is done with stage s: refill it
    if (tid == 0 && t + (int64_t)stages * step < n_tiles) issue(t + (int64_t)stages * step, s);
    if (++s == stages) {
      s = 0;
      phase ^= 1u;
    }
  }
}

template <int U, int R>
static int spmv_stream_go(const Mat *A, int cap, int stages, const double *x, double *y) {
  Ctx &c = ctx();
  const size_t stage = ((size_t)cap + 2) * 8 + ((((size_t)cap + 4) * 4 + 15) & ~(size_t)15) + (R + 4) * 4;
  const size_t smem = stage * stages;
  auto kern = &k_spmv_stream<U, R>;
  if (smem > 48 * 1024) IIFE_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int per_sm = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, R, smem) != cudaSuccess || per_sm < 1) {
    cudaGetLastError();
    per_sm = 1;
  }
  const int64_t n_tiles = (A->n_rows + R - 1) / R;
  const int grid = (int)std::max<int64_t>(1, std::min<int64_t>(n_tiles, (int64_t)c.sm_count * per_sm));
  IIFE_LAUNCH(kern, grid, R, smem, A->rowptr, A->colind, A->val, A->n_rows, A->nnz, cap, stages, x, y);
  IIFE_CHECK_LAUNCH();
  return IIFE_OK;
}

// Measured on M of the N_b = 184 cube (50 M rows, 1-8 entries): what matters is the number of resident threads (every
// thread has one row's gathers in flight), not the depth of the copy pipeline: 256-row tiles x 3 stages sized by rows x
// longest row 0.655 ms; sized by the real tile maximum 0.612; 2 stages 0.462; 128-row tiles x 2 stages 0.433 ms
// (0.95 of the measured copy peak); 3 / 4 stages 0.55 / 0.73; 8 gathers per round 0.47.
int spmv_stream_launch(const Mat *A, int max_len, const double *x, double *y) {
  static const int stages_env = getenv("IIFE_STREAM_STAGES") ? atoi(getenv("IIFE_STREAM_STAGES")) : 2;
  const int stages = std::max(1, std::min(stages_env, STREAM_MAX_STAGES));
  int cap = STREAM_ROWS * max_len;
  if (A->max_tile_entries > 0 && A->max_tile_entries < cap) cap = (A->max_tile_entries + 3) & ~3;
  return spmv_stream_go<4, STREAM_ROWS>(A, cap, stages, x, y);
}

// w = A p, dot = (p, w).  Rows of A index p as well (A square on the local row block).
template <int LPR>
__global__ void __launch_bounds__(SPMV_THREADS)
k_spmv_dot(const int *__restrict__ rowptr, const int *__restrict__ colind, const double *__restrict__ val,
           int64_t n_rows, const double *__restrict__ p, double *__restrict__ w, double *__restrict__ dot_out,
           double *__restrict__ partials, unsigned int *__restrict__ counter, const int *__restrict__ flag, P2PRed pr) {
  if (flag && *flag != 0) return;
  constexpr int RPB = SPMV_THREADS / LPR;
  __shared__ double red[32];
  __shared__ bool is_last;
  int lg = threadIdx.x % LPR;
  int64_t row0 = (int64_t)blockIdx.x * RPB + threadIdx.x / LPR;
  int64_t stride = (int64_t)gridDim.x * RPB;
  int64_t n_iter = (n_rows + stride - 1) / stride;
  double acc = 0.0;
  for (int64_t it = 0; it < n_iter; ++it) {
    int64_t i = row0 + it * stride;
    int b = 0, e = 0;
    if (i < n_rows) {
      b = __ldg(rowptr + i);
      e = __ldg(rowptr + i + 1);
    }
    double s = row_dot<LPR>(colind, val, p, b, e, lg);
    if (i < n_rows && lg == 0) {
      w[i] = s;
      acc = fma(s, __ldg(p + i), acc);
    }
  }
  double bs = block_sum(acc, red);
  if (threadIdx.x == 0) {
    partials[blockIdx.x] = bs;
    __threadfence();
    unsigned int t = atomicAdd(counter, 1u);
    is_last = (t == gridDim.x - 1);
  }
  __syncthreads();
  if (is_last) {
    __threadfence();
    double s = 0.0;
    for (int k = threadIdx.x; k < (int)gridDim.x; k += blockDim.x) s += __ldcg(partials + k);
    s = block_sum(s, red);
    __shared__ double s_sum;
    if (threadIdx.x == 0) {
      *dot_out = s;
      *counter = 0u;
      s_sum = s;
    }
    if (pr.enabled) {  // row-partitioned solver: hand the partial to every rank (no wait here)
      __syncthreads();
      p2p_push(pr, 2ull * (*pr.iter + (unsigned long long)pr.k_off) + 1ull, &s_sum, 1, threadIdx.x);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// SELL-32 operator copy for the Krylov loop.
// The CSR kernels above give one row to LPR lanes: every row costs a chain of three dependent global
// loads (rowptr -> colind/val -> x) and a shuffle reduction, which the ncu profile showed to be
// latency bound (long-scoreboard stalls, 24 % DRAM utilisation at full occupancy).  A_b is solved with
// hundreds of times per extraction and has near-uniform rows ((2p+1)^d B-spline stencils), so it is
// re-laid out once per value update as sliced ELLPACK: slices of 32 consecutive rows, entries stored
// column-major inside a slice.  One THREAD owns one row: the warp's loads of (colind, val) are 128-
// and 256-byte coalesced streams, four of them are in flight per thread, there is no reduction, and
// the only irregular access is the gather of x (L2 resident).  Padding entries carry val = 0 and the
// row's own column.  Rejected (CSR stays in use) when padding exceeds 25 % of nnz.
// ------------------------------------------------------------------------------------------------
__global__ void k_sell_widths(const int *__restrict__ rowptr, int64_t n_rows, int64_t n_slices, int *__restrict__ entries) {
  int lane = threadIdx.x & 31;
  int64_t w = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  int64_t nw = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t s = w; s < n_slices; s += nw) {
    int64_t i = s * 32 + lane;
    int len = i < n_rows ? rowptr[i + 1] - rowptr[i] : 0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) len = max(len, __shfl_xor_sync(0xffffffffu, len, o));
    if (lane == 0) entries[s] = len * 32;
  }
}

// A slice is stored with SHARED column offsets (one word per slot instead of 32) when its rows can be ALIGNED to the
// offsets of its longest row: every entry of every row sits in the slot whose offset (column - row) it has, the slots a
// shorter row lacks are padded with zeros that read an owned entry of x.  Rows at the ends of a grid line then share the
// slice's 27 offsets with their interior neighbours instead of making the whole slice carry 32 x width column words.
// The walk below is the same in the decision (k_sell_align) and the fill (k_sell_fill): in slot order, a row takes its
// next entry if the offsets agree, a zero otherwise; the slice aligns if every row ends with all entries placed.
struct SellWalk {
  int ref_lane, ref_b, ref_i;
};
__device__ __forceinline__ SellWalk sell_walk_begin(int len, int width, int b, int64_t i) {
  SellWalk wk;
  const unsigned longest = __ballot_sync(0xffffffffu, len == width);
  wk.ref_lane = longest ? __ffs(longest) - 1 : 0;
  wk.ref_b = __shfl_sync(0xffffffffu, b, wk.ref_lane);
  wk.ref_i = __shfl_sync(0xffffffffu, (int)i, wk.ref_lane);
  return wk;
}

// per slice: number of column words of the compact layout (width if the slice aligns, 32 x width otherwise)
__global__ void k_sell_align(const int *__restrict__ rowptr, const int *__restrict__ colind, int64_t n_rows, int64_t n_pad_lim,
                             int64_t n_slices, const int *__restrict__ sell_ptr, int *__restrict__ words, int force_full) {
  int lane = threadIdx.x & 31;
  int64_t w = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  int64_t nw = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t s = w; s < n_slices; s += nw) {
    const int64_t i = s * 32 + lane;
    const bool live = i < n_rows;
    int b = 0, len = 0;
    if (live) {
      b = rowptr[i];
      len = rowptr[i + 1] - b;
    }
    const int width = (sell_ptr[s + 1] - sell_ptr[s]) >> 5;
    const SellWalk wk = sell_walk_begin(len, width, b, i);
    bool ok = true;
    int q = 0;
    for (int k = 0; k < width; ++k) {
      const int off = __ldg(colind + wk.ref_b + k) - wk.ref_i;  // same address in all lanes: one broadcast load
      if (live) {
        if (q < len && __ldg(colind + b + q) - (int)i == off) ++q;
        else ok = ok && (i + off >= 0) && (i + off < n_pad_lim);  // a padded slot reads x[i + off]: must be an owned entry
      }
    }
    ok = ok && (q == len);
    ok = __all_sync(0xffffffffu, ok) && !force_full && width > 0;
    if (lane == 0) words[s] = ok ? width : width * 32;
  }
}

// values (and, with sell_col != nullptr, the compact columns) of every slice
__global__ void k_sell_fill(const int *__restrict__ rowptr, const int *__restrict__ colind, const double *__restrict__ val,
                            int64_t n_rows, int64_t n_cols, int64_t n_slices, const int *__restrict__ sell_ptr,
                            const int *__restrict__ sell_cptr, int *__restrict__ sell_col, double *__restrict__ sell_val,
                            double *__restrict__ dinv_out) {
  int lane = threadIdx.x & 31;
  int64_t w = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  int64_t nw = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t s = w; s < n_slices; s += nw) {
    const int64_t i = s * 32 + lane;
    int b = 0, len = 0;
    if (i < n_rows) {
      b = rowptr[i];
      len = rowptr[i + 1] - b;
    }
    const int sb = sell_ptr[s], width = (sell_ptr[s + 1] - sb) >> 5;
    const int cb = sell_cptr[s];
    const bool shared = (sell_cptr[s + 1] - cb) == width;
    const bool all_full = __all_sync(0xffffffffu, len == width);
    double diag = 0.0;  // the Jacobi diagonal falls out of the same pass (dinv_out: PCJACOBI's inverse, zero -> 1)
    if (!shared || all_full) {  // entries in CSR order, short rows padded at the end with (own column, 0)
      const int pad_col = (i < n_cols) ? (int)i : 0;
      for (int k = 0; k < width; ++k) {
        const bool in = k < len;
        const int c = in ? colind[b + k] : pad_col;
        const double v = in ? val[b + k] : 0.0;
        if (sell_col) {
          if (!shared) sell_col[cb + k * 32 + lane] = c;
          else if (lane == 0) sell_col[cb + k] = c - (int)i;
        }
        sell_val[sb + k * 32 + lane] = v;
        if (in && c == (int)i) diag += v;
      }
    } else {  // aligned slice with short rows: the walk of k_sell_align
      const SellWalk wk = sell_walk_begin(len, width, b, i);
      int q = 0;
      for (int k = 0; k < width; ++k) {
        const int off = __ldg(colind + wk.ref_b + k) - wk.ref_i;
        const bool take = q < len && __ldg(colind + b + q) - (int)i == off;
        const double v = take ? val[b + q] : 0.0;
        sell_val[sb + k * 32 + lane] = v;
        if (take) {
          ++q;
          if (off == 0) diag += v;
        }
        if (sell_col && lane == 0) sell_col[cb + k] = off;
      }
    }
    if (dinv_out && i < n_rows) dinv_out[i] = (diag == 0.0) ? 1.0 : 1.0 / diag;
  }
}

// one SELL slice (32 rows, one per lane): returns this lane's row sum.  CG: gather x with ld.global.cg
// (L2 only) — used for rows whose ghost entries were written by a peer GPU during this kernel.
// Column indices are stored COMPACTLY: a slice whose 32 rows all have the same column OFFSETS
// (col = row + off_k: every interior slice of a stencil operator such as the B-spline A_b) keeps one
// offset per k instead of 32 columns, which removes a third of the bytes an SpMV has to stream
// (12 -> 8.1 B per entry).  A slice is "uniform" iff it stores exactly `width` column words.
template <int U, bool CG = false>
__device__ __forceinline__ double sell_slice(const int *__restrict__ sell_ptr, const int *__restrict__ sell_cptr,
                                             const int *__restrict__ sell_col, const double *__restrict__ sell_val,
                                             const double *__restrict__ x, int64_t s, int lane, int64_t n_rows) {
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
    const int64_t i = s * 32 + lane;
    const int row = (i < n_rows) ? (int)i : 0;
    const int live = (i < n_rows) ? 1 : 0;
    const int *op = sell_col + cb;
    for (; k + U <= width; k += U) {
      int c[U];
      double v[U];
#pragma unroll
      for (int u = 0; u < U; ++u) c[u] = row + live * __ldg(op + k + u);
#pragma unroll
      for (int u = 0; u < U; ++u) v[u] = ld_stream(vp + (k + u) * 32);
#pragma unroll
      for (int u = 0; u < U; ++u) a[u] = fma(v[u], CG ? __ldcg(x + c[u]) : __ldg(x + c[u]), a[u]);
    }
    for (; k < width; ++k) {
      const int c1 = row + live * __ldg(op + k);
      a[0] = fma(ld_stream(vp + k * 32), CG ? __ldcg(x + c1) : __ldg(x + c1), a[0]);
    }
  } else {
    const int *cp = sell_col + cb + lane;
    for (; k + U <= width; k += U) {
      int c[U];
      double v[U];
#pragma unroll
      for (int u = 0; u < U; ++u) c[u] = ld_stream(cp + (k + u) * 32);
#pragma unroll
      for (int u = 0; u < U; ++u) v[u] = ld_stream(vp + (k + u) * 32);
#pragma unroll
      for (int u = 0; u < U; ++u) a[u] = fma(v[u], CG ? __ldcg(x + c[u]) : __ldg(x + c[u]), a[u]);
    }
    for (; k < width; ++k) {
      const int c1 = ld_stream(cp + k * 32);
      a[0] = fma(ld_stream(vp + k * 32), CG ? __ldcg(x + c1) : __ldg(x + c1), a[0]);
    }
  }
  double acc = a[0];
#pragma unroll
  for (int u = 1; u < U; ++u) acc += a[u];
  return acc;
}

template <bool DOT, int U, bool IFIRST>
__global__ void IIFE_SELL_BOUNDS
k_spmv_sell(const int *__restrict__ sell_ptr, const int *__restrict__ sell_cptr, const int *__restrict__ sell_col,
            const double *__restrict__ sell_val, int64_t n_rows, int64_t n_slices, const double *__restrict__ x, double *__restrict__ y,
            double *__restrict__ dot_out, double *__restrict__ partials, unsigned int *__restrict__ counter,
            const int *__restrict__ flag, P2PRed pr, HaloWait hw) {
  if (flag && *flag != 0) return;  // converged: the rest of the enqueued chunk is a row of no-ops
  __shared__ double red[32];
  __shared__ bool is_last;
  if (hw.flags && blockIdx.x == 0 && threadIdx.x == 0) cg_trace(pr, TR_S_IN);
  if (!IFIRST && hw.flags) {  // three-kernel CG iteration: the neighbours' k_cg_p_push stores the ghost entries of x
    const int q = threadIdx.x;
    if (q < hw.nranks && ((hw.recv_mask >> q) & 1u)) spin_until(hw.flags + q, *hw.seq_base + (unsigned long long)hw.k_off + 1ull, hw.err);
    __syncthreads();
    if (blockIdx.x == 0 && threadIdx.x == 0) cg_trace(pr, TR_S_WAITED);
  }
  const int lane = threadIdx.x & 31;
  const int64_t w0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nw = ((int64_t)gridDim.x * blockDim.x) >> 5;
  double dsum = 0.0;
  if (!IFIRST) {
    for (int64_t s = w0; s < n_slices; s += nw) {
      const double acc = sell_slice<U>(sell_ptr, sell_cptr, sell_col, sell_val, x, s, lane, n_rows);
      const int64_t i = s * 32 + lane;
      if (i < n_rows) {
        y[i] = acc;
        if (DOT) dsum = fma(acc, __ldg(x + i), dsum);
      }
    }
  } else {
    // interior slices (no ghost column): the ghost entries may still be in flight
    for (int64_t v = w0; v < hw.n_int; v += nw) {
      const int64_t s = hw.int_lo + v;
      const double acc = sell_slice<U>(sell_ptr, sell_cptr, sell_col, sell_val, x, s, lane, n_rows);
      const int64_t i = s * 32 + lane;
      if (i < n_rows) {
        y[i] = acc;
        if (DOT) dsum = fma(acc, __ldg(x + i), dsum);
      }
    }
    // boundary slices [0, int_lo) and [int_lo + n_int, n_slices), handed out from the far end of the warp list (the
    // warps with one interior slice fewer); each warp waits for the flags itself; x is read past L1 (the ghost
    // entries landed during this kernel, and a line holding the last owned entries may sit in L1 with stale ghosts)
    const int64_t n_bnd = n_slices - hw.n_int;
    const int64_t b0 = nw - 1 - w0;
    if (b0 < n_bnd) {
      if (lane < hw.nranks && ((hw.recv_mask >> lane) & 1u))
        spin_until(hw.flags + lane, *hw.seq_base + (unsigned long long)hw.k_off + 1ull, hw.err);
      __syncwarp();
      if (b0 == 0 && lane == 0) cg_trace(pr, TR_S_WAITED);
      for (int64_t b = b0; b < n_bnd; b += nw) {
        const int64_t s = b < hw.int_lo ? b : b + hw.n_int;
        const double acc = sell_slice<U, true>(sell_ptr, sell_cptr, sell_col, sell_val, x, s, lane, n_rows);
        const int64_t i = s * 32 + lane;
        if (i < n_rows) {
          y[i] = acc;
          if (DOT) dsum = fma(acc, __ldcg(x + i), dsum);
        }
      }
    }
  }
  if (DOT) {
    double bs = block_sum(dsum, red);
    if (threadIdx.x == 0) {
      partials[blockIdx.x] = bs;
      __threadfence();
      unsigned int t = atomicAdd(counter, 1u);
      is_last = (t == gridDim.x - 1);
    }
    __syncthreads();
    if (is_last) {
      __threadfence();
      double sacc = 0.0;
      for (int kk = threadIdx.x; kk < (int)gridDim.x; kk += blockDim.x) sacc += __ldcg(partials + kk);
      sacc = block_sum(sacc, red);
      __shared__ double s_sum;
      if (threadIdx.x == 0) {
        *dot_out = sacc;
        *counter = 0u;
        s_sum = sacc;
      }
      if (pr.enabled) {  // row-partitioned solver: hand the partial to every rank (no wait here)
        __syncthreads();
        p2p_push(pr, 2ull * (*pr.iter + (unsigned long long)pr.k_off) + 1ull, &s_sum, 1, threadIdx.x);
        if (hw.flags && threadIdx.x == 0) cg_trace(pr, TR_S_OUT);
      }
    }
  }
}

// slices whose rows reference a ghost column (col >= n_owned)
__global__ void k_sell_ghost_flag(const int *__restrict__ sell_ptr, const int *__restrict__ sell_cptr,
                                  const int *__restrict__ sell_col, int64_t n_rows, int64_t n_slices, int n_owned,
                                  int *__restrict__ is_interior, int *__restrict__ is_boundary) {
  int lane = threadIdx.x & 31;
  int64_t w = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  int64_t nw = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t s = w; s < n_slices; s += nw) {
    int width = (sell_ptr[s + 1] - sell_ptr[s]) >> 5;
    int cb = sell_cptr[s], words = sell_cptr[s + 1] - cb;
    bool ghost = false;
    if (words == width) {
      int64_t i = s * 32 + lane;
      for (int k = 0; k < width; ++k) ghost |= (i < n_rows) && ((int)i + sell_col[cb + k] >= n_owned);
    } else {
      for (int p = cb + lane; p < cb + words; p += 32) ghost |= (sell_col[p] >= n_owned);
    }
    ghost = __any_sync(0xffffffffu, ghost);
    if (lane == 0) {
      is_interior[s] = ghost ? 0 : 1;
      is_boundary[s] = ghost ? 1 : 0;
    }
  }
}

__global__ void k_sell_order(const int *__restrict__ is_interior, const int *__restrict__ off_int,
                             const int *__restrict__ off_bnd, int64_t n_slices, int n_interior, int *__restrict__ order) {
  int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (; s < n_slices; s += stride) order[is_interior[s] ? off_int[s] : n_interior + off_bnd[s]] = (int)s;
}

// (x, y) with the usual last-CTA fixed-order reduction (and the push to the peers' mailboxes in the
// row-partitioned solver).  Used after the plain SELL SpMV: fusing the dot into the SpMV kernel made the
// compiler serialise the slice's loads (392 us fused against 294 us plain + ~20 us for this kernel).
__global__ void __launch_bounds__(SPMV_THREADS)
k_vec_dot(const double *__restrict__ x, const double *__restrict__ y, int64_t n, double *__restrict__ dot_out,
          double *__restrict__ partials, unsigned int *__restrict__ counter, const int *__restrict__ flag, P2PRed pr) {
  if (flag && *flag != 0) return;
  __shared__ double red[32];
  __shared__ bool is_last;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  double d0 = 0.0, d1 = 0.0, d2 = 0.0, d3 = 0.0;
  for (; i + 3 * stride < n; i += 4 * stride) {
    const double x0 = __ldg(x + i), x1 = __ldg(x + i + stride), x2 = __ldg(x + i + 2 * stride), x3 = __ldg(x + i + 3 * stride);
    const double y0 = __ldcs(y + i), y1 = __ldcs(y + i + stride), y2 = __ldcs(y + i + 2 * stride), y3 = __ldcs(y + i + 3 * stride);
    d0 = fma(x0, y0, d0);
    d1 = fma(x1, y1, d1);
    d2 = fma(x2, y2, d2);
    d3 = fma(x3, y3, d3);
  }
  for (; i < n; i += stride) d0 = fma(__ldg(x + i), y[i], d0);
  double bs = block_sum((d0 + d1) + (d2 + d3), red);
  if (threadIdx.x == 0) {
    partials[blockIdx.x] = bs;
    __threadfence();
    unsigned int t = atomicAdd(counter, 1u);
    is_last = (t == gridDim.x - 1);
  }
  __syncthreads();
  if (is_last) {
    __threadfence();
    double sacc = 0.0;
    for (int kk = threadIdx.x; kk < (int)gridDim.x; kk += blockDim.x) sacc += __ldcg(partials + kk);
    sacc = block_sum(sacc, red);
    __shared__ double s_sum;
    if (threadIdx.x == 0) {
      *dot_out = sacc;
      *counter = 0u;
      s_sum = sacc;
    }
    if (pr.enabled) {
      __syncthreads();
      p2p_push(pr, 2ull * (*pr.iter + (unsigned long long)pr.k_off) + 1ull, &s_sum, 1, threadIdx.x);
    }
  }
}

// grid = SMs x resident CTAs of THIS kernel (occupancy query), so the persistent loop has no tail wave
template <class K>
static int resident_grid(K kernel, int64_t need) {
  int per_sm = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, SPMV_THREADS, 0) != cudaSuccess || per_sm < 1) {
    cudaGetLastError();
    per_sm = 4;
  }
  int64_t cap = (int64_t)ctx().sm_count * per_sm;
  if (need > cap) {
    // even work per CTA: with a few items per warp (a rank's block of a partitioned operator) a grid of
    // exactly `cap` CTAs leaves some warps one item more than others — a 15-25 % tail
    int64_t per_cta = (need + cap - 1) / cap;
    need = (need + per_cta - 1) / per_cta;
  }
  if (need < 1) need = 1;
  return (int)need;
}

static int sell_unroll() {
  static int u = -1;
  if (u < 0) {
    const char *e = getenv("IIFE_SELL_UNROLL");
    u = e ? atoi(e) : 4;
    if (u != 2 && u != 4 && u != 8) u = 4;
  }
  return u;
}

int spmv_launch_signature() {
  const char *fd = getenv("IIFE_SELL_FUSED_DOT");
  return sell_unroll() | ((fd && atoi(fd) == 0) ? 16 : 0);
}

static int launch_sell(const Mat *A, bool dot, const double *x, double *y, double *dot_out, double *partials,
                       unsigned int *counter, const int *flag, const P2PRed *red_in = nullptr, const HaloWait *hw_in = nullptr) {
  P2PRed pr{};
  if (red_in) pr = *red_in;
  HaloWait hw{};
  if (hw_in) hw = *hw_in;
  int64_t need = (A->sell_slices + (SPMV_THREADS / 32) - 1) / (SPMV_THREADS / 32);
#define SELL_GO(D, UU, IF)                                                                                          \
  {                                                                                                                 \
    int g = resident_grid(k_spmv_sell<D, UU, IF>, need);                                                            \
    IIFE_LAUNCH((k_spmv_sell<D, UU, IF>), g, SPMV_THREADS, 0, A->sell_ptr, A->sell_cptr, A->sell_col, A->sell_val, A->n_rows, \
                A->sell_slices, x, y, dot_out, partials, counter, flag, pr, hw);                                    \
  }
  int u = sell_unroll();
  if (dot && hw.flags && hw.interior_first) {
    SELL_GO(true, 4, true)
  } else if (dot) {
    if (u == 2) SELL_GO(true, 2, false) else if (u == 8) SELL_GO(true, 8, false) else SELL_GO(true, 4, false)
  } else {
    if (u == 2) SELL_GO(false, 2, false) else if (u == 8) SELL_GO(false, 8, false) else SELL_GO(false, 4, false)
  }
#undef SELL_GO
  IIFE_CHECK_LAUNCH();
  return IIFE_OK;
}

static int sell_grid(int64_t n_slices);
static bool sell_ready(const Mat *A);

// Interior slices of a row-partitioned operator block (columns >= n_owned are ghosts): when they form ONE run
// [sell_int_lo, sell_int_lo + sell_n_interior) — row blocks of a banded operator: the boundary rows sit at the two ends —
// the three-kernel CG iteration multiplies them before it waits for the ghost entries (k_spmv_sell<.,.,true>).
// sell_int_lo = -1 otherwise (the wait then stays in the prologue).
int mat_ensure_sell_order(Mat *A, int64_t n_owned) {
  if (A->sell_state != 1) return IIFE_OK;
  if (A->sell_order_owned == n_owned) return IIFE_OK;
  const int64_t ns = A->sell_slices;
  Tmp<int> fi, fb, oi, ob, order;
  IIFE_TRY(fi.alloc((size_t)ns + 1));
  IIFE_TRY(fb.alloc((size_t)ns + 1));
  IIFE_TRY(oi.alloc((size_t)ns + 1));
  IIFE_TRY(ob.alloc((size_t)ns + 1));
  IIFE_TRY(order.alloc((size_t)ns));
  IIFE_LAUNCH(k_sell_ghost_flag, sell_grid(ns), SPMV_THREADS, 0, A->sell_ptr, A->sell_cptr, A->sell_col, A->n_rows, ns, (int)n_owned, fi.p, fb.p);
  IIFE_CHECK_LAUNCH();
  int64_t n_int = 0, n_bnd = 0;
  IIFE_TRY(exclusive_scan_i32(fi.p, oi.p, ns, &n_int));
  IIFE_TRY(exclusive_scan_i32(fb.p, ob.p, ns, &n_bnd));
  IIFE_LAUNCH(k_sell_order, sell_grid(ns), SPMV_THREADS, 0, fi.p, oi.p, ob.p, ns, (int)n_int, order.p);
  IIFE_CHECK_LAUNCH();
  int ends[2] = {0, 0};  // first and last interior slice (k_sell_order keeps the slices in order inside each class)
  if (n_int > 0) {
    IIFE_CUDA(cudaMemcpyAsync(&ends[0], order.p, sizeof(int), cudaMemcpyDeviceToHost, ctx().stream));
    IIFE_CUDA(cudaMemcpyAsync(&ends[1], order.p + (n_int - 1), sizeof(int), cudaMemcpyDeviceToHost, ctx().stream));
  }
  IIFE_CUDA(cudaStreamSynchronize(ctx().stream));
  A->sell_n_interior = n_int;
  A->sell_int_lo = (n_int > 0 && (int64_t)ends[1] - ends[0] + 1 == n_int) ? ends[0] : -1;
  A->sell_order_owned = n_owned;
  return IIFE_OK;
}

void mat_free_sell(Mat *A) {
  A->sell_order_owned = -1;
  if (A->sell_ptr) dev_free_t(A->sell_ptr, (size_t)A->sell_slices + 1);
  if (A->sell_col) dev_free_t(A->sell_col, (size_t)A->sell_cwords);
  if (A->sell_cptr) dev_free_t(A->sell_cptr, (size_t)A->sell_slices + 1);
  A->sell_cptr = nullptr;
  if (A->sell_val) dev_free_t(A->sell_val, (size_t)A->sell_padded);
  A->sell_ptr = A->sell_col = nullptr;
  A->sell_val = nullptr;
  A->sell_state = 0;
  A->sell_vals_valid = false;
}

static int sell_grid(int64_t n_slices) {
  int64_t need = (n_slices + (SPMV_THREADS / 32) - 1) / (SPMV_THREADS / 32);
  int64_t cap = (int64_t)ctx().sm_count * 8;
  if (need > cap) need = cap;
  if (need < 1) need = 1;
  return (int)need;
}

int mat_ensure_sell(Mat *A, bool want_dinv) {
  static const bool disabled = getenv("IIFE_NO_SELL") != nullptr;
  if (A->sell_state == -1 || disabled || A->n_rows == 0 || A->nnz == 0) {
    A->sell_state = -1;
    return IIFE_OK;
  }
  Ctx &c = ctx();
  // a fill pass that runs anyway also yields the inverse Jacobi diagonal (saves the separate CSR pass, 1.1 ms at N_b=184)
  double *dinv_out = nullptr;
  const bool will_fill = A->sell_state == 0 || !A->sell_vals_valid;
  if (want_dinv && will_fill && !A->dinv_valid) {
    if (!A->dinv) IIFE_TRY(dev_alloc_t(&A->dinv, (size_t)A->n_rows));
    dinv_out = A->dinv;
  }
  if (A->sell_state == 0) {
    int64_t n_slices = (A->n_rows + 31) / 32;
    Tmp<int> entries;
    IIFE_TRY(entries.alloc((size_t)n_slices + 1));
    IIFE_LAUNCH(k_sell_widths, sell_grid(n_slices), SPMV_THREADS, 0, A->rowptr, A->n_rows, n_slices, entries.p);
    IIFE_CHECK_LAUNCH();
    int *ptr = nullptr;
    IIFE_TRY(dev_alloc_t(&ptr, (size_t)n_slices + 1));
    int64_t padded = 0;
    int rc = exclusive_scan_i32(entries.p, ptr, n_slices, &padded);
    if (rc == IIFE_ERR_UNSUPPORTED || (rc == IIFE_OK && padded > A->nnz + A->nnz / 4 + 1024)) {
      dev_free_t(ptr, (size_t)n_slices + 1);
      A->sell_state = -1;  // too much padding (or > int32): stay on CSR
      return IIFE_OK;
    }
    if (rc != IIFE_OK) {
      dev_free_t(ptr, (size_t)n_slices + 1);
      return rc;
    }
    A->sell_ptr = ptr;
    A->sell_slices = n_slices;
    A->sell_padded = padded;
    rc = dev_alloc_t(&A->sell_val, (size_t)padded);
    if (rc != IIFE_OK) {
      mat_free_sell(A);
      return rc;
    }
    // which slices share their column offsets (k_sell_align), then values and compact columns in one pass
    {
      Tmp<int> words;
      int *cptr = nullptr;
      rc = words.alloc((size_t)n_slices + 1);
      if (rc == IIFE_OK) rc = dev_alloc_t(&cptr, (size_t)n_slices + 1);
      if (rc == IIFE_OK) {
        A->sell_cptr = cptr;
        static const bool no_compress = getenv("IIFE_SELL_NOCOMPRESS") != nullptr;
        IIFE_LAUNCH(k_sell_align, sell_grid(n_slices), SPMV_THREADS, 0, A->rowptr, A->colind, A->n_rows,
                    std::min<int64_t>(A->n_rows, A->n_cols), n_slices, A->sell_ptr, words.p, no_compress ? 1 : 0);
        int64_t cw = 0;
        rc = exclusive_scan_i32(words.p, cptr, n_slices, &cw);
        if (rc == IIFE_OK) {
          A->sell_cwords = cw;
          rc = dev_alloc_t(&A->sell_col, (size_t)cw);
        }
        if (rc == IIFE_OK) {
          IIFE_LAUNCH(k_sell_fill, sell_grid(n_slices), SPMV_THREADS, 0, A->rowptr, A->colind, A->val, A->n_rows, A->n_cols, n_slices,
                      A->sell_ptr, A->sell_cptr, A->sell_col, A->sell_val, dinv_out);
          cudaError_t e = cudaStreamSynchronize(c.stream);
          if (e != cudaSuccess) rc = set_err(IIFE_ERR_CUDA, "SELL build: %s", cudaGetErrorString(e));
        }
      }
      if (rc != IIFE_OK) {
        mat_free_sell(A);
        return rc;
      }
    }
    A->sell_state = 1;
    A->sell_vals_valid = true;  // k_sell_fill above wrote the values too
    if (dinv_out) A->dinv_valid = true;
  }
  if (!A->sell_vals_valid) {
    IIFE_LAUNCH(k_sell_fill, sell_grid(A->sell_slices), SPMV_THREADS, 0, A->rowptr, A->colind, A->val, A->n_rows, A->n_cols,
                A->sell_slices, A->sell_ptr, A->sell_cptr, (int *)nullptr, A->sell_val, dinv_out);
    IIFE_CHECK_LAUNCH();
    A->sell_vals_valid = true;
    if (dinv_out) A->dinv_valid = true;
  }
  (void)c;
  return IIFE_OK;
}

static bool sell_ready(const Mat *A) { return A->sell_state == 1 && A->sell_vals_valid; }
bool mat_sell_ready(const Mat *A) { return sell_ready(A); }

int spmv_pick_lpr(const Mat *A) {
  double mean = A->n_rows ? (double)A->nnz / (double)A->n_rows : 0.0;
  if (mean <= 2.5) return 2;
  if (mean <= 5.0) return 4;
  if (mean <= 10.0) return 8;
  if (mean <= 20.0) return 16;
  return 32;
}

static int spmv_grid(int64_t n_rows, int lpr) {
  int rpb = SPMV_THREADS / lpr;
  int64_t need = (n_rows + rpb - 1) / rpb;
  int64_t cap = (int64_t)ctx().sm_count * 8;  // 8 CTAs of 256 threads = 2048 resident threads per SM
  if (need > cap) need = cap;
  if (need < 1) need = 1;
  return (int)need;
}

int spmv_launch(const Mat *A, double alpha, const double *x, double beta, double *y) {
  if (A->n_rows == 0) return IIFE_OK;
  if (sell_ready(A) && alpha == 1.0 && beta == 0.0)
    return launch_sell(A, false, x, y, nullptr, nullptr, nullptr, nullptr);
  int lpr = spmv_pick_lpr(A);
  int g = spmv_grid(A->n_rows, lpr);
  bool plain = (alpha == 1.0 && beta == 0.0);
  if (plain && lpr <= 8 && A->max_row_len < 0) {  // first product with this operator: one reduction over the row lengths
    int ml = 0;
    IIFE_TRY(mat_max_row_len(const_cast<Mat *>(A), &ml));
  }
  if (plain && lpr <= 8 && A->max_row_len >= 0 && A->max_row_len <= 16 && A->nnz >= 1024) {
    // short ragged rows: tiles staged with bulk async copies (k_spmv_stream); IIFE_SPMV_STREAM=0 disables
    const char *st = getenv("IIFE_SPMV_STREAM");
    if (!st || atoi(st) != 0) return spmv_stream_launch(A, A->max_row_len < 1 ? 1 : A->max_row_len, x, y);
  }
  if (plain && lpr <= 8) {
    const char *ilp = getenv("IIFE_SPMV_ILP");  // two-rows-in-flight kernel for short rows (default on)
    if (!ilp || atoi(ilp) != 0) {
      int g2 = spmv_grid((A->n_rows + 1) / 2, lpr);
      if (lpr == 2) IIFE_LAUNCH((k_spmv_ilp2<2>), g2, SPMV_THREADS, 0, A->rowptr, A->colind, A->val, A->n_rows, x, y);
      else if (lpr == 4) IIFE_LAUNCH((k_spmv_ilp2<4>), g2, SPMV_THREADS, 0, A->rowptr, A->colind, A->val, A->n_rows, x, y);
      else IIFE_LAUNCH((k_spmv_ilp2<8>), g2, SPMV_THREADS, 0, A->rowptr, A->colind, A->val, A->n_rows, x, y);
      IIFE_CHECK_LAUNCH();
      return IIFE_OK;
    }
  }
#define SPMV_CASE(L)                                                                                              \
  case L:                                                                                                         \
    if (plain) IIFE_LAUNCH((k_spmv<L, true>), g, SPMV_THREADS, 0, A->rowptr, A->colind, A->val, A->n_rows, alpha, x, beta, y); \
    else IIFE_LAUNCH((k_spmv<L, false>), g, SPMV_THREADS, 0, A->rowptr, A->colind, A->val, A->n_rows, alpha, x, beta, y);      \
    break;
  switch (lpr) {
    SPMV_CASE(2)
    SPMV_CASE(4)
    SPMV_CASE(8)
    SPMV_CASE(16)
    default:
      SPMV_CASE(32)
  }
#undef SPMV_CASE
  IIFE_CHECK_LAUNCH();
  return IIFE_OK;
}

int spmv_dot_launch(const Mat *A, const double *p, double *w, double *dot_out, double *partials,
                    unsigned int *counter, const int *flag, const P2PRed *red, const HaloWait *hw) {
  if (hw && !sell_ready(A)) return set_err(IIFE_ERR_STATE, "the in-kernel halo wait needs the SELL operator copy");
  if (sell_ready(A)) {
    if (hw) return launch_sell(A, true, p, w, dot_out, partials, counter, flag, red, hw);
    // dot fused into the SpMV (default since the kernel has 64 registers: 407 -> 395 us per CG iteration);
    // IIFE_SELL_FUSED_DOT=0 selects the plain SpMV followed by the dot kernel
    static const bool fused = !(getenv("IIFE_SELL_FUSED_DOT") && atoi(getenv("IIFE_SELL_FUSED_DOT")) == 0);
    if (fused) return launch_sell(A, true, p, w, dot_out, partials, counter, flag, red);
    // plain SpMV followed by the dot kernel, both gated by the reason flag
    IIFE_TRY(launch_sell(A, false, p, w, nullptr, nullptr, nullptr, flag));
    P2PRed pr2{};
    if (red) pr2 = *red;
    int64_t need = (A->n_rows + SPMV_THREADS * 4 - 1) / (SPMV_THREADS * 4);
    int g = resident_grid(k_vec_dot, need);
    IIFE_LAUNCH(k_vec_dot, g, SPMV_THREADS, 0, p, (const double *)w, A->n_rows, dot_out, partials, counter, flag, pr2);
    IIFE_CHECK_LAUNCH();
    return IIFE_OK;
  }
  P2PRed pr{};
  if (red) pr = *red;
  int lpr = spmv_pick_lpr(A);
  int g = spmv_grid(A->n_rows, lpr);
#define SPMVD_CASE(L)                                                                                            \
  case L:                                                                                                        \
    IIFE_LAUNCH(k_spmv_dot<L>, g, SPMV_THREADS, 0, A->rowptr, A->colind, A->val, A->n_rows, p, w, dot_out, partials, counter, flag, pr); \
    break;
  switch (lpr) {
    SPMVD_CASE(2)
    SPMVD_CASE(4)
    SPMVD_CASE(8)
    SPMVD_CASE(16)
    default:
      SPMVD_CASE(32)
  }
#undef SPMVD_CASE
  IIFE_CHECK_LAUNCH();
  return IIFE_OK;
}

}  // namespace iife

using namespace iife;

extern "C" int iife_spmv(iife_mat A_, int trans, double alpha, const double *x, double beta, double *y, int mem) {
  IIFE_NEED_INIT();
  Mat *A = (Mat *)A_;
  if (!A || !x || !y) return set_err(IIFE_ERR_ARG, "NULL argument");
  if (trans) {
    IIFE_TRY(mat_ensure_transpose(A));
    A = A->T;
    // M^T has near-uniform rows (27 entries on the cube), so its products (AT_x, reference la_utils.py:143-163)
    // take the SELL-32 kernel on a SELL copy of the cached transpose: 1.05 -> 0.41 ms (0.37 -> 0.93 of the copy
    // peak) at N_b=184.  mat_ensure_sell keeps CSR when padding would exceed 25 %.  IIFE_SPMV_SELL_T=0 disables.
    const char *sell_t = getenv("IIFE_SPMV_SELL_T");  // read per call: scripts/compare_variants.py toggles it
    if ((!sell_t || atoi(sell_t) != 0) && alpha == 1.0 && beta == 0.0) IIFE_TRY(mat_ensure_sell(A));
  }
  if (mem == IIFE_MEM_DEVICE) return spmv_launch(A, alpha, x, beta, y);
  Tmp<double> dx, dy;
  IIFE_TRY(dx.alloc((size_t)A->n_cols));
  IIFE_TRY(dy.alloc((size_t)A->n_rows));
  cudaStream_t s = ctx().stream;
  IIFE_CUDA(cudaMemcpyAsync(dx.p, x, (size_t)A->n_cols * sizeof(double), cudaMemcpyHostToDevice, s));
  if (beta != 0.0) IIFE_CUDA(cudaMemcpyAsync(dy.p, y, (size_t)A->n_rows * sizeof(double), cudaMemcpyHostToDevice, s));
  IIFE_TRY(spmv_launch(A, alpha, dx.p, beta, dy.p));
  IIFE_CUDA(cudaMemcpyAsync(y, dy.p, (size_t)A->n_rows * sizeof(double), cudaMemcpyDeviceToHost, s));
  IIFE_CUDA(cudaStreamSynchronize(s));
  return IIFE_OK;
}
