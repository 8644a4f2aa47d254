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

namespace iife {

constexpr int SPMV_THREADS = 256;

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

// w = A p, dot = (p, w).  Rows of A index p as well (A square on the local row block).
template <int LPR>
__global__ void __launch_bounds__(SPMV_THREADS)
k_spmv_dot(const int *__restrict__ rowptr, const int *__restrict__ colind, const double *__restrict__ val,
           int64_t n_rows, const double *__restrict__ p, double *__restrict__ w, double *__restrict__ dot_out,
           double *__restrict__ partials, unsigned int *__restrict__ counter, const int *__restrict__ flag) {
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
    if (threadIdx.x == 0) {
      *dot_out = s;
      *counter = 0u;
    }
  }
}

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
  int lpr = spmv_pick_lpr(A);
  int g = spmv_grid(A->n_rows, lpr);
  bool plain = (alpha == 1.0 && beta == 0.0);
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
                    unsigned int *counter, const int *flag) {
  int lpr = spmv_pick_lpr(A);
  int g = spmv_grid(A->n_rows, lpr);
#define SPMVD_CASE(L)                                                                                            \
  case L:                                                                                                        \
    IIFE_LAUNCH(k_spmv_dot<L>, g, SPMV_THREADS, 0, A->rowptr, A->colind, A->val, A->n_rows, p, w, dot_out, partials, counter, flag); \
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
