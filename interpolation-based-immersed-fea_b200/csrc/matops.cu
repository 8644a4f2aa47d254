// matops.cu — row edits of the background operator A_b ("next" row N3 of SURVEY.md §8f): the matrix
// operations behind the reference's basis-function-removal helpers
//   trimNodes              reference common.py:262-332   (MatZeroRows, diagonal 1.0)
//   removeZeroDiagonal     reference common.py:236-251   (A += diag(vd): MatAXPY with a different pattern)
//   getIdentity            reference common.py:254-258
// Both produce a NEW device matrix; the Python mirror swaps the handle to give the in-place behaviour of
// petsc4py.  Semantics restated from PETSc (absent here):
//   MatZeroRows on AIJ without KEEP_NONZERO_PATTERN: a listed row keeps exactly one entry, (i, i) = diag, when
//   diag != 0 and i < n_cols, and no entry otherwise; other rows are untouched.
//   MatAXPY(A, 1, D, DIFFERENT_NONZERO_PATTERN) with D = MatDiagonalSet(vd, INSERT) on an empty matrix: pattern
//   = union(A, full diagonal), rows stay column-sorted, (i, i) = A_ii + vd_i.
#include "common.cuh"

namespace iife {

static int grid_rows(int64_t n, int threads = 256) {
  int64_t g = (n + threads - 1) / threads;
  int64_t cap = (int64_t)ctx().sm_count * 32;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return (int)g;
}

template <class I>
__global__ void k_mark_rows(const I *__restrict__ rows, int64_t n_listed, int64_t n_rows, unsigned char *__restrict__ flag,
                            int *__restrict__ bad) {
  int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (; t < n_listed; t += stride) {
    long long r = (long long)rows[t];
    if (r < 0 || r >= n_rows) atomicOr(bad, 1);
    else flag[r] = 1;  // duplicates are fine: same byte, same value
  }
}

__global__ void k_zero_rows_count(const int *__restrict__ rowptr, const unsigned char *__restrict__ flag, int64_t n_rows,
                                  int64_t n_cols, int keep_diag, int *__restrict__ new_len) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (; i < n_rows; i += stride)
    new_len[i] = flag[i] ? ((keep_diag && i < n_cols) ? 1 : 0) : (rowptr[i + 1] - rowptr[i]);
}

// warp per row: copy an untouched row, or write the single diagonal entry of a zeroed one
__global__ void __launch_bounds__(256)
k_zero_rows_fill(const int *__restrict__ rowptr, const int *__restrict__ colind, const double *__restrict__ val,
                 const unsigned char *__restrict__ flag, int64_t n_rows, double diag, const int *__restrict__ new_rowptr,
                 int *__restrict__ new_col, double *__restrict__ new_val) {
  const int lane = threadIdx.x & 31;
  int64_t w = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nw = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (; w < n_rows; w += nw) {
    const int ob = new_rowptr[w], on = new_rowptr[w + 1] - ob;
    if (flag[w]) {
      if (lane == 0 && on == 1) {
        new_col[ob] = (int)w;
        new_val[ob] = diag;
      }
    } else {
      const int b = rowptr[w];
      for (int e = lane; e < on; e += 32) {
        new_col[ob + e] = colind[b + e];
        new_val[ob + e] = val[b + e];
      }
    }
  }
}

// number of columns < i in the (sorted) row and whether i itself is stored
__device__ __forceinline__ void diag_position(const int *__restrict__ colind, int b, int n, int i, int *pos, bool *has) {
  int lo = 0, hi = n;
  while (lo < hi) {
    int mid = (lo + hi) >> 1;
    if (colind[b + mid] < i) lo = mid + 1;
    else hi = mid;
  }
  *pos = lo;
  *has = lo < n && colind[b + lo] == i;
}

__global__ void k_add_diag_count(const int *__restrict__ rowptr, const int *__restrict__ colind, int64_t n_rows,
                                 int64_t n_cols, int *__restrict__ new_len) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (; i < n_rows; i += stride) {
    const int b = rowptr[i], n = rowptr[i + 1] - b;
    int add = 0;
    if (i < n_cols) {
      int pos;
      bool has;
      diag_position(colind, b, n, (int)i, &pos, &has);
      add = has ? 0 : 1;
    }
    new_len[i] = n + add;
  }
}

__global__ void __launch_bounds__(256)
k_add_diag_fill(const int *__restrict__ rowptr, const int *__restrict__ colind, const double *__restrict__ val,
                const double *__restrict__ d, int64_t n_rows, int64_t n_cols, const int *__restrict__ new_rowptr,
                int *__restrict__ new_col, double *__restrict__ new_val) {
  const int lane = threadIdx.x & 31;
  int64_t w = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nw = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (; w < n_rows; w += nw) {
    const int b = rowptr[w], n = rowptr[w + 1] - b;
    const int ob = new_rowptr[w];
    if (w >= n_cols) {  // no diagonal position in this row
      for (int e = lane; e < n; e += 32) {
        new_col[ob + e] = colind[b + e];
        new_val[ob + e] = val[b + e];
      }
      continue;
    }
    int pos;
    bool has;
    diag_position(colind, b, n, (int)w, &pos, &has);  // every lane computes the same answer
    const int shift = has ? 0 : 1;
    for (int e = lane; e < n; e += 32) {
      const int o = e < pos ? e : e + shift;
      double v = val[b + e];
      if (has && e == pos) v += d[w];
      new_col[ob + o] = colind[b + e];
      new_val[ob + o] = v;
    }
    if (!has && lane == 0) {
      new_col[ob + pos] = (int)w;
      new_val[ob + pos] = 0.0 + d[w];
    }
  }
}

static int finish_rowptr(const int *new_len, int64_t n_rows, Mat **out, int64_t n_cols, Tmp<int> &new_rowptr) {
  IIFE_TRY(new_rowptr.alloc((size_t)n_rows + 1));
  int64_t total = 0;
  IIFE_TRY(exclusive_scan_i32(new_len, new_rowptr.p, n_rows, &total));
  if (total > 0x7fffffffLL) return set_err(IIFE_ERR_UNSUPPORTED, "result has more than 2^31-1 entries");
  IIFE_TRY(mat_alloc(out, n_rows, n_cols, total));
  IIFE_CUDA(cudaMemcpyAsync((*out)->rowptr, new_rowptr.p, ((size_t)n_rows + 1) * sizeof(int), cudaMemcpyDeviceToDevice,
                            ctx().stream));
  return IIFE_OK;
}

}  // namespace iife

using namespace iife;

extern "C" {

int iife_mat_zero_rows(iife_mat A_, const void *rows, int64_t n_listed, int idx_bytes, double diag, int mem,
                       iife_mat *out) {
  IIFE_NEED_INIT();
  Mat *A = (Mat *)A_;
  if (!A || !out || (n_listed > 0 && !rows)) return set_err(IIFE_ERR_ARG, "NULL argument");
  if (n_listed < 0) return set_err(IIFE_ERR_ARG, "negative row count");
  if (idx_bytes != 4 && idx_bytes != 8) return set_err(IIFE_ERR_ARG, "idx_bytes must be 4 or 8");
  cudaStream_t st = ctx().stream;
  const int64_t n = A->n_rows;
  Tmp<unsigned char> flag;
  Tmp<int> bad, new_len, new_rowptr;
  IIFE_TRY(flag.alloc((size_t)n));
  IIFE_TRY(bad.alloc(1));
  IIFE_TRY(new_len.alloc((size_t)n));
  IIFE_CUDA(cudaMemsetAsync(flag.p, 0, (size_t)(n ? n : 1), st));
  IIFE_CUDA(cudaMemsetAsync(bad.p, 0, sizeof(int), st));
  Tmp<unsigned char> stage;  // host row list staged on the device
  const void *rows_dev = rows;
  if (n_listed > 0 && mem == IIFE_MEM_HOST) {
    IIFE_TRY(stage.alloc((size_t)n_listed * idx_bytes));
    IIFE_CUDA(cudaMemcpyAsync(stage.p, rows, (size_t)n_listed * idx_bytes, cudaMemcpyHostToDevice, st));
    rows_dev = stage.p;
  }
  if (n_listed > 0) {
    if (idx_bytes == 4)
      IIFE_LAUNCH(k_mark_rows<int>, grid_rows(n_listed), 256, 0, (const int *)rows_dev, n_listed, n, flag.p, bad.p);
    else
      IIFE_LAUNCH(k_mark_rows<long long>, grid_rows(n_listed), 256, 0, (const long long *)rows_dev, n_listed, n, flag.p,
                  bad.p);
    IIFE_CHECK_LAUNCH();
    int h_bad = 0;
    IIFE_CUDA(cudaMemcpyAsync(&h_bad, bad.p, sizeof(int), cudaMemcpyDeviceToHost, st));
    IIFE_CUDA(cudaStreamSynchronize(st));
    if (h_bad) return set_err(IIFE_ERR_ARG, "row index out of range [0, %lld)", (long long)n);
  }
  const int keep_diag = diag != 0.0 ? 1 : 0;
  Mat *C = nullptr;
  if (n > 0) {
    IIFE_LAUNCH(k_zero_rows_count, grid_rows(n), 256, 0, A->rowptr, flag.p, n, A->n_cols, keep_diag, new_len.p);
    IIFE_CHECK_LAUNCH();
  }
  IIFE_TRY(finish_rowptr(new_len.p, n, &C, A->n_cols, new_rowptr));
  if (n > 0) {
    IIFE_LAUNCH(k_zero_rows_fill, grid_rows(n * 32), 256, 0, A->rowptr, A->colind, A->val, flag.p, n, diag, C->rowptr,
                C->colind, C->val);
    if (cudaGetLastError() != cudaSuccess) {
      mat_free(C);
      return set_err(IIFE_ERR_CUDA, "k_zero_rows_fill launch failed");
    }
  }
  IIFE_CUDA(cudaStreamSynchronize(st));  // temporaries are released on return
  *out = (iife_mat)C;
  return IIFE_OK;
}

int iife_mat_add_diagonal(iife_mat A_, const double *d, int mem, iife_mat *out) {
  IIFE_NEED_INIT();
  Mat *A = (Mat *)A_;
  if (!A || !out || (!d && A->n_rows > 0)) return set_err(IIFE_ERR_ARG, "NULL argument");
  cudaStream_t st = ctx().stream;
  const int64_t n = A->n_rows;
  Tmp<double> stage;
  const double *d_dev = d;
  if (n > 0 && mem == IIFE_MEM_HOST) {
    IIFE_TRY(stage.alloc((size_t)n));
    IIFE_CUDA(cudaMemcpyAsync(stage.p, d, (size_t)n * sizeof(double), cudaMemcpyHostToDevice, st));
    d_dev = stage.p;
  }
  Tmp<int> new_len, new_rowptr;
  IIFE_TRY(new_len.alloc((size_t)n));
  Mat *C = nullptr;
  if (n > 0) {
    IIFE_LAUNCH(k_add_diag_count, grid_rows(n), 256, 0, A->rowptr, A->colind, n, A->n_cols, new_len.p);
    IIFE_CHECK_LAUNCH();
  }
  IIFE_TRY(finish_rowptr(new_len.p, n, &C, A->n_cols, new_rowptr));
  if (n > 0) {
    IIFE_LAUNCH(k_add_diag_fill, grid_rows(n * 32), 256, 0, A->rowptr, A->colind, A->val, d_dev, n, A->n_cols, C->rowptr,
                C->colind, C->val);
    if (cudaGetLastError() != cudaSuccess) {
      mat_free(C);
      return set_err(IIFE_ERR_CUDA, "k_add_diag_fill launch failed");
    }
  }
  IIFE_CUDA(cudaStreamSynchronize(st));
  *out = (iife_mat)C;
  return IIFE_OK;
}

}  // extern "C"
