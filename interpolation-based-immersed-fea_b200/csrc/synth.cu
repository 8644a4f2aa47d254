// synth.cu — on-device generator of BASELINE config 5, the "S1 fitted cube" of SURVEY.md §8(d).
// Bench/test tooling, not part of the reference's path: it only produces the operands (A_f, M, b_f)
// the path consumes, directly in HBM, so the 50 M-DOF case does not have to be built on the host.
//
//   background  : N^3 cells, trilinear B-splines, n_b = (N+1)^3 nodes, id = bx + (N+1)(by + (N+1) bz)
//   foreground  : every background cell split 2x2x2, each sub-cube into 6 Kuhn tetrahedra,
//                 n_f = (2N+1)^3 vertices, id = x + nv (y + nv z)
//   A_f         : P1 stiffness + sigma * mass; vertex v couples to v+d for the 15 offsets d whose
//                 components all have the same sign (edges of the Kuhn triangulation)
//   M           : M[v, node] = trilinear hat of the node at v (1, 2, 4 or 8 entries of 1, 1/2, 1/4, 1/8)
//   b_f         : load vector of f = 1
// Values are sums of per-cell contributions read from a host-computed table (iife_b200/synthetic.py)
// in a fixed cell order, so the host generator reproduces them bit for bit.
#include "common.cuh"

namespace iife {

__device__ __forceinline__ bool kuhn_offset(int dx, int dy, int dz) {
  bool nonneg = dx >= 0 && dy >= 0 && dz >= 0;
  bool nonpos = dx <= 0 && dy <= 0 && dz <= 0;
  return nonneg || nonpos;
}

__global__ void k_synth_counts(int nv, int64_t row_begin, int64_t n_rows, int *__restrict__ lenA,
                               int *__restrict__ lenM) {
  int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (; t < n_rows; t += stride) {
    int64_t j = row_begin + t;
    int x = (int)(j % nv), y = (int)((j / nv) % nv), z = (int)(j / ((int64_t)nv * nv));
    int ca = 0;
    for (int dz = -1; dz <= 1; ++dz)
      for (int dy = -1; dy <= 1; ++dy)
        for (int dx = -1; dx <= 1; ++dx) {
          if (!kuhn_offset(dx, dy, dz)) continue;
          int xx = x + dx, yy = y + dy, zz = z + dz;
          if (xx < 0 || yy < 0 || zz < 0 || xx >= nv || yy >= nv || zz >= nv) continue;
          ++ca;
        }
    lenA[t] = ca;
    lenM[t] = ((x & 1) + 1) * ((y & 1) + 1) * ((z & 1) + 1);
  }
}

__global__ void k_synth_fill(int nv, int nb, int64_t row_begin, int64_t n_rows, const double *__restrict__ coef,
                             const double *__restrict__ load8, const int *__restrict__ a_rowptr,
                             int *__restrict__ a_col, double *__restrict__ a_val, const int *__restrict__ m_rowptr,
                             int *__restrict__ m_col, double *__restrict__ m_val, double *__restrict__ b_f) {
  __shared__ double s_coef[8 * 27];
  __shared__ double s_load[8];
  for (int k = threadIdx.x; k < 8 * 27; k += blockDim.x) s_coef[k] = coef[k];
  if (threadIdx.x < 8) s_load[threadIdx.x] = load8[threadIdx.x];
  __syncthreads();
  int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const int ncell = nv - 1;
  for (; t < n_rows; t += stride) {
    int64_t j = row_begin + t;
    int x = (int)(j % nv), y = (int)((j / nv) % nv), z = (int)(j / ((int64_t)nv * nv));
    // cells touching the vertex: lower corner = v - c, c in {0,1}^3
    bool cell_ok[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      int cx = x - (c & 1), cy = y - ((c >> 1) & 1), cz = z - ((c >> 2) & 1);
      cell_ok[c] = cx >= 0 && cy >= 0 && cz >= 0 && cx < ncell && cy < ncell && cz < ncell;
    }
    int p = a_rowptr[t];
    for (int dz = -1; dz <= 1; ++dz)
      for (int dy = -1; dy <= 1; ++dy)
        for (int dx = -1; dx <= 1; ++dx) {
          if (!kuhn_offset(dx, dy, dz)) continue;
          int xx = x + dx, yy = y + dy, zz = z + dz;
          if (xx < 0 || yy < 0 || zz < 0 || xx >= nv || yy >= nv || zz >= nv) continue;
          int d = (dx + 1) + 3 * (dy + 1) + 9 * (dz + 1);
          double v = 0.0;
#pragma unroll
          for (int c = 0; c < 8; ++c)
            if (cell_ok[c]) v += s_coef[c * 27 + d];  // table rows are zero where c+d leaves the cell
          a_col[p] = (int)(xx + (int64_t)nv * (yy + (int64_t)nv * zz));
          a_val[p] = v;
          ++p;
        }
    double bl = 0.0;
#pragma unroll
    for (int c = 0; c < 8; ++c)
      if (cell_ok[c]) bl += s_load[c];
    if (b_f) b_f[t] = bl;
    // M row: tensor product of the 1D hats
    int q = m_rowptr[t];
    int bx0 = x >> 1, by0 = y >> 1, bz0 = z >> 1;
    int nx = (x & 1) + 1, ny = (y & 1) + 1, nz = (z & 1) + 1;
    double wx = (x & 1) ? 0.5 : 1.0, wy = (y & 1) ? 0.5 : 1.0, wz = (z & 1) ? 0.5 : 1.0;
    for (int kz = 0; kz < nz; ++kz)
      for (int ky = 0; ky < ny; ++ky)
        for (int kx = 0; kx < nx; ++kx) {
          m_col[q] = (bx0 + kx) + nb * ((by0 + ky) + nb * (bz0 + kz));
          m_val[q] = wx * wy * wz;
          ++q;
        }
  }
}

}  // namespace iife

using namespace iife;

static int grid_for_rows(int64_t n) {
  int64_t g = (n + 255) / 256;
  int64_t cap = (int64_t)ctx().sm_count * 16;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return (int)g;
}

static int synth_check(int64_t n_bg_cells, int64_t row_begin, int64_t row_end, int64_t *nv_out) {
  if (n_bg_cells < 1 || n_bg_cells > 600) return set_err(IIFE_ERR_ARG, "n_bg_cells %lld out of range [1,600]", (long long)n_bg_cells);
  int64_t nv = 2 * n_bg_cells + 1;
  int64_t n_f = nv * nv * nv;
  if (n_f >= 0x7fffffffLL) return set_err(IIFE_ERR_UNSUPPORTED, "n_f %lld exceeds int32", (long long)n_f);
  if (row_begin < 0 || row_end < row_begin || row_end > n_f) return set_err(IIFE_ERR_ARG, "bad row range");
  *nv_out = nv;
  return IIFE_OK;
}

extern "C" int iife_synth_cube_counts(int64_t n_bg_cells, int64_t row_begin, int64_t row_end, int64_t *nnz_A,
                                      int64_t *nnz_M) {
  IIFE_NEED_INIT();
  int64_t nv = 0;
  IIFE_TRY(synth_check(n_bg_cells, row_begin, row_end, &nv));
  int64_t n = row_end - row_begin;
  Tmp<int> la, lm, oa, om;
  IIFE_TRY(la.alloc((size_t)n + 1));
  IIFE_TRY(lm.alloc((size_t)n + 1));
  IIFE_TRY(oa.alloc((size_t)n + 1));
  IIFE_TRY(om.alloc((size_t)n + 1));
  if (n) IIFE_LAUNCH(k_synth_counts, grid_for_rows(n), 256, 0, (int)nv, row_begin, n, la.p, lm.p);
  IIFE_CHECK_LAUNCH();
  int64_t ta = 0, tm = 0;
  // totals can exceed int32 only if the caller asks for too many rows at once: report, do not wrap
  int rc = exclusive_scan_i32(la.p, oa.p, n, &ta);
  if (rc != IIFE_OK && rc != IIFE_ERR_UNSUPPORTED) return rc;
  rc = exclusive_scan_i32(lm.p, om.p, n, &tm);
  if (rc != IIFE_OK && rc != IIFE_ERR_UNSUPPORTED) return rc;
  if (nnz_A) *nnz_A = ta;
  if (nnz_M) *nnz_M = tm;
  return IIFE_OK;
}

extern "C" int iife_synth_cube_build(int64_t n_bg_cells, int64_t row_begin, int64_t row_end, const double *coef_8x27,
                                     const double *load8, iife_mat *A_f, iife_mat *M, double *b_f_dev) {
  IIFE_NEED_INIT();
  if (!coef_8x27 || !load8 || !A_f || !M) return set_err(IIFE_ERR_ARG, "NULL argument");
  *A_f = nullptr;
  *M = nullptr;
  int64_t nv = 0;
  IIFE_TRY(synth_check(n_bg_cells, row_begin, row_end, &nv));
  int64_t nb = n_bg_cells + 1;
  int64_t n = row_end - row_begin;
  int64_t n_f = nv * nv * nv, n_b = nb * nb * nb;
  cudaStream_t s = ctx().stream;
  Tmp<int> la, lm;
  Tmp<double> dcoef, dload;
  IIFE_TRY(la.alloc((size_t)n + 1));
  IIFE_TRY(lm.alloc((size_t)n + 1));
  IIFE_TRY(dcoef.alloc(8 * 27));
  IIFE_TRY(dload.alloc(8));
  IIFE_CUDA(cudaMemcpyAsync(dcoef.p, coef_8x27, 8 * 27 * sizeof(double), cudaMemcpyHostToDevice, s));
  IIFE_CUDA(cudaMemcpyAsync(dload.p, load8, 8 * sizeof(double), cudaMemcpyHostToDevice, s));
  if (n) IIFE_LAUNCH(k_synth_counts, grid_for_rows(n), 256, 0, (int)nv, row_begin, n, la.p, lm.p);
  IIFE_CHECK_LAUNCH();
  // row pointers first (totals give the allocation sizes)
  Tmp<int> oa, om;
  IIFE_TRY(oa.alloc((size_t)n + 1));
  IIFE_TRY(om.alloc((size_t)n + 1));
  int64_t ta = 0, tm = 0;
  IIFE_TRY(exclusive_scan_i32(la.p, oa.p, n, &ta));
  IIFE_TRY(exclusive_scan_i32(lm.p, om.p, n, &tm));
  Mat *A = nullptr, *Mm = nullptr;
  IIFE_TRY(mat_alloc(&A, n, n_f, ta));
  int rc = mat_alloc(&Mm, n, n_b, tm);
  if (rc != IIFE_OK) {
    mat_free(A);
    return rc;
  }
  cudaMemcpyAsync(A->rowptr, oa.p, ((size_t)n + 1) * sizeof(int), cudaMemcpyDeviceToDevice, s);
  cudaMemcpyAsync(Mm->rowptr, om.p, ((size_t)n + 1) * sizeof(int), cudaMemcpyDeviceToDevice, s);
  if (n)
    IIFE_LAUNCH(k_synth_fill, grid_for_rows(n), 256, 0, (int)nv, (int)nb, row_begin, n, dcoef.p, dload.p, A->rowptr,
                A->colind, A->val, Mm->rowptr, Mm->colind, Mm->val, b_f_dev);
  cudaError_t e = cudaStreamSynchronize(s);
  if (e == cudaSuccess) e = cudaGetLastError();
  if (e != cudaSuccess) {
    mat_free(A);
    mat_free(Mm);
    return set_err(IIFE_ERR_CUDA, "synthetic cube: %s", cudaGetErrorString(e));
  }
  *A_f = (iife_mat)A;
  *M = (iife_mat)Mm;
  return IIFE_OK;
}
