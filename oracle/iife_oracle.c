/*
 * iife_oracle.c — CPU restatement of the extraction linear-algebra path of
 * jefromm/interpolation-based-immersed-fea.  TEST INFRASTRUCTURE ONLY: nothing under oracle/ is
 * imported, linked or executed by the product (libiife.so / iife_b200 / the la_utils mirror); it is
 * used by tests/, by __graft_entry__.smoke() as the checker, and by bench.py's cpu_baseline /
 * --impl reference legs.
 *
 * PARITY UNPINNED: the reference delegates all arithmetic of this path to PETSc through petsc4py
 * (version not pinned anywhere in the reference; FEniCS 2019.1-era, PETSc ~3.12-3.17), PETSc is not
 * vendored under /root/reference and is not installable here, and the reference ships no tests or
 * golden outputs for this path.  Every function below therefore restates the *published* algorithm of
 * the PETSc routine the reference calls, anchored on the reference call site it stands in for:
 *
 *   oracle_transpose          MatTranspose          la_utils.py:178,180  (A.transpose())
 *   oracle_spgemm_*           MatMatMult AIJ*AIJ    la_utils.py:179,181  (AT.matMult(R), ATR.matMult(ATT))
 *                             symbolic = structural boolean product, rows column-sorted, no numeric
 *                             dropping; numeric = row-wise Gustavson, inner index ascending
 *   oracle_spmv               MatMult               la_utils.py:141, common.py:139
 *   oracle_spmv_transpose     MatMultTranspose      la_utils.py:162      (AT_x)
 *   oracle_cg_jacobi          KSPCG + PCJACOBI      common.py:554-574, 628-636 (method='cg')
 *   oracle_fgmres_jacobi      KSPFGMRES + PCJACOBI  common.py:554-574, 628-636 (method='gmres' -> FGMRES, restart 300)
 *
 * Indices: int64 row pointers, int32 column indices; values fp64.  OpenMP is used over rows /
 * vector entries only (no change of per-row summation order), so results are independent of the
 * thread count except for the dot products, which use a fixed blocked reduction.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef int64_t i64;
typedef int32_t i32;

int oracle_max_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

void oracle_set_threads(int n) {
#ifdef _OPENMP
  if (n > 0) omp_set_num_threads(n);
#else
  (void)n;
#endif
}

/* ---- MatTranspose: stable counting sort, so every row of the result is column-sorted ---- */
void oracle_transpose(i64 n_rows, i64 n_cols, const i64 *rp, const i32 *ci, const double *v, i64 *t_rp, i32 *t_ci,
                      double *t_v) {
  i64 nnz = rp[n_rows];
  memset(t_rp, 0, (size_t)(n_cols + 1) * sizeof(i64));
  for (i64 p = 0; p < nnz; ++p) t_rp[ci[p] + 1]++;
  for (i64 c = 0; c < n_cols; ++c) t_rp[c + 1] += t_rp[c];
  i64 *cur = (i64 *)malloc((size_t)(n_cols + 1) * sizeof(i64));
  memcpy(cur, t_rp, (size_t)(n_cols + 1) * sizeof(i64));
  for (i64 i = 0; i < n_rows; ++i)
    for (i64 p = rp[i]; p < rp[i + 1]; ++p) {
      i64 q = cur[ci[p]]++;
      t_ci[q] = (i32)i;
      if (v) t_v[q] = v[p];
    }
  free(cur);
}

static int cmp_i32(const void *a, const void *b) {
  i32 x = *(const i32 *)a, y = *(const i32 *)b;
  return (x > y) - (x < y);
}

/* ---- MatMatMult symbolic: per-row number of structurally non-zero columns of X*Y ---- */
void oracle_spgemm_count(i64 n_rows, i64 n_cols_y, const i64 *xr, const i32 *xc, const i64 *yr, const i32 *yc,
                         i64 *cnt) {
#pragma omp parallel
  {
    i64 *mark = (i64 *)malloc((size_t)(n_cols_y > 0 ? n_cols_y : 1) * sizeof(i64));
    for (i64 c = 0; c < n_cols_y; ++c) mark[c] = -1;
#pragma omp for schedule(dynamic, 256)
    for (i64 i = 0; i < n_rows; ++i) {
      i64 n = 0;
      for (i64 p = xr[i]; p < xr[i + 1]; ++p) {
        i32 j = xc[p];
        for (i64 q = yr[j]; q < yr[j + 1]; ++q) {
          i32 k = yc[q];
          if (mark[k] != i) {
            mark[k] = i;
            ++n;
          }
        }
      }
      cnt[i] = n;
    }
    free(mark);
  }
}

/* ---- symbolic fill: sorted column lists (c_rp is the exclusive scan of cnt) ---- */
void oracle_spgemm_fill(i64 n_rows, i64 n_cols_y, const i64 *xr, const i32 *xc, const i64 *yr, const i32 *yc,
                        const i64 *c_rp, i32 *c_ci) {
#pragma omp parallel
  {
    i64 *mark = (i64 *)malloc((size_t)(n_cols_y > 0 ? n_cols_y : 1) * sizeof(i64));
    for (i64 c = 0; c < n_cols_y; ++c) mark[c] = -1;
#pragma omp for schedule(dynamic, 256)
    for (i64 i = 0; i < n_rows; ++i) {
      i64 o = c_rp[i];
      for (i64 p = xr[i]; p < xr[i + 1]; ++p) {
        i32 j = xc[p];
        for (i64 q = yr[j]; q < yr[j + 1]; ++q) {
          i32 k = yc[q];
          if (mark[k] != i) {
            mark[k] = i;
            c_ci[o++] = k;
          }
        }
      }
      qsort(c_ci + c_rp[i], (size_t)(o - c_rp[i]), sizeof(i32), cmp_i32);
    }
    free(mark);
  }
}

/* ---- MatMatMult numeric: row-wise Gustavson into the given (sorted) pattern.
 * C[i,:] = sum over p in X row i (ascending) of X[i,j_p] * Y[j_p,:]; a dense accumulator per thread
 * keeps the additions for one output entry in ascending-j order, as PETSc's sequential kernel does. */
void oracle_spgemm_numeric(i64 n_rows, i64 n_cols_y, const i64 *xr, const i32 *xc, const double *xv, const i64 *yr,
                           const i32 *yc, const double *yv, const i64 *c_rp, const i32 *c_ci, double *c_v) {
#pragma omp parallel
  {
    double *acc = (double *)calloc((size_t)(n_cols_y > 0 ? n_cols_y : 1), sizeof(double));
#pragma omp for schedule(dynamic, 256)
    for (i64 i = 0; i < n_rows; ++i) {
      for (i64 p = xr[i]; p < xr[i + 1]; ++p) {
        i32 j = xc[p];
        double a = xv[p];
        for (i64 q = yr[j]; q < yr[j + 1]; ++q) acc[yc[q]] += a * yv[q];
      }
      for (i64 o = c_rp[i]; o < c_rp[i + 1]; ++o) {
        c_v[o] = acc[c_ci[o]];
        acc[c_ci[o]] = 0.0;
      }
    }
    free(acc);
  }
}

/* ---- MatMult ---- */
void oracle_spmv(i64 n_rows, const i64 *rp, const i32 *ci, const double *v, const double *x, double *y) {
#pragma omp parallel for schedule(static)
  for (i64 i = 0; i < n_rows; ++i) {
    double s = 0.0;
    for (i64 p = rp[i]; p < rp[i + 1]; ++p) s += v[p] * x[ci[p]];
    y[i] = s;
  }
}

/* ---- MatMultTranspose: y = A^T x as PETSc's SeqAIJ kernel does it (scatter in row order) ---- */
void oracle_spmv_transpose(i64 n_rows, i64 n_cols, const i64 *rp, const i32 *ci, const double *v, const double *x,
                           double *y) {
  for (i64 c = 0; c < n_cols; ++c) y[c] = 0.0;
  for (i64 i = 0; i < n_rows; ++i) {
    double xi = x[i];
    for (i64 p = rp[i]; p < rp[i + 1]; ++p) y[ci[p]] += v[p] * xi;
  }
}

/* ---- MatGetDiagonal + PCJACOBI setup: missing diagonal reads 0; zero -> 1 ---- */
void oracle_jacobi_inverse(i64 n, const i64 *rp, const i32 *ci, const double *v, double *dinv) {
#pragma omp parallel for schedule(static)
  for (i64 i = 0; i < n; ++i) {
    double d = 0.0;
    for (i64 p = rp[i]; p < rp[i + 1]; ++p)
      if (ci[p] == i) d = v[p];
    dinv[i] = d == 0.0 ? 1.0 : 1.0 / d;
  }
}

static double dot(i64 n, const double *a, const double *b) {
  /* fixed blocked reduction: independent of the OpenMP thread count */
  const i64 B = 4096;
  i64 nb = (n + B - 1) / B;
  double total = 0.0;
  double *part = (double *)malloc((size_t)(nb > 0 ? nb : 1) * sizeof(double));
#pragma omp parallel for schedule(static)
  for (i64 k = 0; k < nb; ++k) {
    i64 lo = k * B, hi = lo + B < n ? lo + B : n;
    double s = 0.0;
    for (i64 i = lo; i < hi; ++i) s += a[i] * b[i];
    part[k] = s;
  }
  for (i64 k = 0; k < nb; ++k) total += part[k];
  free(part);
  return total;
}

/* KSPConvergedDefault: returns the reason (0 = keep iterating) */
static int converged_default(double rnorm, double ttol, double atol, double dtol, double rho0) {
  if (isnan(rnorm) || isinf(rnorm)) return -9;
  if (rnorm <= ttol) return rnorm < atol ? 3 : 2;
  if (rnorm >= dtol * rho0) return -4;
  return 0;
}

/* ---- KSPCG, left Jacobi, preconditioned norm.  dinv == NULL means PCNONE.
 * Returns the reason; *its_out iterations; hist[0..min(its,hist_len-1)] the preconditioned norms. */
int oracle_cg_jacobi(i64 n, const i64 *rp, const i32 *ci, const double *v, const double *dinv_in, const double *b,
                     double *x, double rtol, double atol, double dtol, i64 max_it, i64 *its_out, double *rnorm_out,
                     double *hist, i64 hist_len) {
  double *r = (double *)malloc((size_t)(n + 1) * sizeof(double));
  double *z = (double *)malloc((size_t)(n + 1) * sizeof(double));
  double *p = (double *)malloc((size_t)(n + 1) * sizeof(double));
  double *w = (double *)malloc((size_t)(n + 1) * sizeof(double));
  int reason = 0;
  i64 i = 0, its = 0;
  /* r = b - A x0 */
  oracle_spmv(n, rp, ci, v, x, w);
#pragma omp parallel for schedule(static)
  for (i64 k = 0; k < n; ++k) {
    r[k] = b[k] - w[k];
    z[k] = (dinv_in ? dinv_in[k] : 1.0) * r[k];
    w[k] = (dinv_in ? dinv_in[k] : 1.0) * b[k]; /* D^-1 b for the reference norm */
  }
  double dp = sqrt(dot(n, z, z));
  double rho0 = sqrt(dot(n, w, w));
  if (rho0 == 0.0) rho0 = dp; /* KSPConvergedDefault: zero rhs, nonzero guess -> initial residual norm */
  double ttol = fmax(rtol * rho0, atol);
  if (hist && hist_len > 0) hist[0] = dp;
  reason = converged_default(dp, ttol, atol, dtol, rho0);
  double beta = 0.0, betaold = 1.0;
  if (!reason && max_it <= 0) reason = -3;
  while (!reason) {
    its = i + 1;
    beta = dot(n, z, r);
    if (beta == 0.0) { reason = 3; break; }
    if (beta < 0.0) { reason = -8; break; }
    if (isnan(beta) || isinf(beta)) { reason = -9; break; }
    if (i == 0) {
      memcpy(p, z, (size_t)n * sizeof(double));
    } else {
      double bb = beta / betaold;
#pragma omp parallel for schedule(static)
      for (i64 k = 0; k < n; ++k) p[k] = z[k] + bb * p[k];
    }
    oracle_spmv(n, rp, ci, v, p, w);
    double dpi = dot(n, p, w);
    betaold = beta;
    if (!(dpi > 0.0)) { reason = isnan(dpi) ? -9 : -10; break; }
    double a = beta / dpi;
#pragma omp parallel for schedule(static)
    for (i64 k = 0; k < n; ++k) {
      x[k] += a * p[k];
      r[k] -= a * w[k];
      z[k] = (dinv_in ? dinv_in[k] : 1.0) * r[k];
    }
    dp = sqrt(dot(n, z, z));
    if (hist && i + 1 < hist_len) hist[i + 1] = dp;
    reason = converged_default(dp, ttol, atol, dtol, rho0);
    if (reason) break;
    ++i;
    if (i >= max_it) { reason = -3; its = i; break; }
  }
  if (its_out) *its_out = its;
  if (rnorm_out) *rnorm_out = dp;
  free(r); free(z); free(p); free(w);
  return reason;
}

/* ---- KSPFGMRES(m), right Jacobi, true residual norm, classical Gram-Schmidt ---- */
/* R_out (optional, m x m column-major, leading dimension m) receives the triangular factor of the LAST cycle's
 * Hessenberg matrix and *k_out its order: the rotations are orthogonal, so R has the singular values of the
 * (k+1) x k Hessenberg that KSPComputeExtremeSingularValues works on (reference common.py:483-507). */
static int fgmres_impl(i64 n, const i64 *rp, const i32 *ci, const double *v, const double *dinv_in, const double *b,
                       double *x, double rtol, double atol, double dtol, i64 max_it, int m, i64 *its_out,
                       double *rnorm_out, double *hist, i64 hist_len, double *R_out, i64 *k_out) {
  if (m < 1) m = 30;
  if (max_it > 0 && (i64)m > max_it) m = (int)max_it;
  double **V = (double **)calloc((size_t)m + 1, sizeof(double *));
  double **Z = (double **)calloc((size_t)m + 1, sizeof(double *));
  double *H = (double *)calloc((size_t)(m + 1) * (size_t)m, sizeof(double));
  double *cs = (double *)calloc((size_t)m + 1, sizeof(double));
  double *sn = (double *)calloc((size_t)m + 1, sizeof(double));
  double *rs = (double *)calloc((size_t)m + 2, sizeof(double));
  double *y = (double *)calloc((size_t)m + 1, sizeof(double));
  double *hcol = (double *)calloc((size_t)m + 2, sizeof(double));
  const int m1 = m + 1;
  int reason = 0;
  i64 its = 0;
  double rho0 = sqrt(dot(n, b, b));
  double ttol = fmax(rtol * rho0, atol);
  double res = 0.0;
  int first = 1;
  while (!reason) {
    if (!V[0]) V[0] = (double *)malloc((size_t)(n + 1) * sizeof(double));
    /* r = b - A x */
    oracle_spmv(n, rp, ci, v, x, V[0]);
#pragma omp parallel for schedule(static)
    for (i64 k = 0; k < n; ++k) V[0][k] = b[k] - V[0][k];
    res = sqrt(dot(n, V[0], V[0]));
    if (first && hist && hist_len > 0) hist[0] = res;
    if (first && rho0 == 0.0) { /* KSPConvergedDefault: zero rhs, nonzero guess -> initial residual norm */
      rho0 = res;
      ttol = fmax(rtol * rho0, atol);
    }
    first = 0;
    reason = converged_default(res, ttol, atol, dtol, rho0);
    if (!reason && its >= max_it) reason = -3;
    if (reason) break;
    rs[0] = res;
    double scale = res != 0.0 ? 1.0 / res : 0.0;
    int j = 0;
    int hapend = 0;
    while (!reason && j < m) {
      if (!V[j + 1]) V[j + 1] = (double *)malloc((size_t)(n + 1) * sizeof(double));
      if (!Z[j]) Z[j] = (double *)malloc((size_t)(n + 1) * sizeof(double));
      double *vj = V[j], *zj = Z[j], *wv = V[j + 1];
#pragma omp parallel for schedule(static)
      for (i64 k = 0; k < n; ++k) {
        vj[k] *= scale;
        zj[k] = (dinv_in ? dinv_in[k] : 1.0) * vj[k];
      }
      oracle_spmv(n, rp, ci, v, zj, wv);
      /* classical Gram-Schmidt: all coefficients from the same w, then one update */
      for (int k = 0; k <= j; ++k) hcol[k] = dot(n, wv, V[k]);
#pragma omp parallel for schedule(static)
      for (i64 t = 0; t < n; ++t) {
        double a = wv[t];
        for (int k = 0; k <= j; ++k) a -= hcol[k] * V[k][t];
        wv[t] = a;
      }
      double tt = sqrt(dot(n, wv, wv));
      double hapbnd = fabs(tt / rs[j]);
      if (hapbnd > 1e-30) hapbnd = 1e-30;
      if (tt > hapbnd) scale = 1.0 / tt; else { scale = 0.0; hapend = 1; }
      double *hh = H + (size_t)j * m1;
      for (int k = 0; k <= j; ++k) hh[k] = hcol[k];
      hh[j + 1] = tt;
      for (int k = 0; k < j; ++k) {
        double t1 = hh[k], t2 = hh[k + 1];
        hh[k] = cs[k] * t1 + sn[k] * t2;
        hh[k + 1] = -sn[k] * t1 + cs[k] * t2;
      }
      if (!hapend) {
        double t = sqrt(hh[j] * hh[j] + hh[j + 1] * hh[j + 1]);
        if (t == 0.0) { reason = -5; t = 1.0; }
        cs[j] = hh[j] / t;
        sn[j] = hh[j + 1] / t;
        rs[j + 1] = -sn[j] * rs[j];
        rs[j] = cs[j] * rs[j];
        hh[j] = cs[j] * hh[j] + sn[j] * hh[j + 1];
        res = fabs(rs[j + 1]);
      } else {
        res = 0.0;
        rs[j + 1] = 0.0;
      }
      ++j;
      ++its;
      if (hist && its < hist_len) hist[its] = res;
      if (!reason) reason = converged_default(res, ttol, atol, dtol, rho0);
      if (!reason) {
        if (hapend) reason = -5;
        else if (its >= max_it) reason = -3;
      }
    }
    if (R_out && j > 0) {
      for (int c = 0; c < j; ++c)
        for (int i = 0; i < j; ++i) R_out[(size_t)c * m + i] = (i <= c) ? H[(size_t)c * m1 + i] : 0.0;
      if (k_out) *k_out = j;
    }
    /* x += Z y with H(0:j,0:j) y = rs(0:j) */
    for (int i = j - 1; i >= 0; --i) {
      double s = 0.0;
      for (int c = i + 1; c < j; ++c) s += H[(size_t)c * m1 + i] * y[c];
      double d = H[(size_t)i * m1 + i];
      y[i] = d != 0.0 ? (rs[i] - s) / d : 0.0;
    }
#pragma omp parallel for schedule(static)
    for (i64 t = 0; t < n; ++t) {
      double a = x[t];
      for (int k = 0; k < j; ++k) a += y[k] * Z[k][t];
      x[t] = a;
    }
  }
  if (its_out) *its_out = its;
  if (rnorm_out) *rnorm_out = res;
  for (int k = 0; k <= m; ++k) { free(V[k]); free(Z[k]); }
  free(V); free(Z); free(H); free(cs); free(sn); free(rs); free(y); free(hcol);
  return reason;
}

int oracle_fgmres_jacobi(i64 n, const i64 *rp, const i32 *ci, const double *v, const double *dinv_in, const double *b,
                         double *x, double rtol, double atol, double dtol, i64 max_it, int m, i64 *its_out,
                         double *rnorm_out, double *hist, i64 hist_len) {
  return fgmres_impl(n, rp, ci, v, dinv_in, b, x, rtol, atol, dtol, max_it, m, its_out, rnorm_out, hist, hist_len, NULL,
                     NULL);
}

/* caller: m must already be clipped the way fgmres_impl clips it (m >= 1, m <= max_it when max_it > 0) */
int oracle_fgmres_hessenberg(i64 n, const i64 *rp, const i32 *ci, const double *v, const double *dinv_in,
                             const double *b, double *x, double rtol, double atol, double dtol, i64 max_it, int m,
                             i64 *its_out, double *rnorm_out, double *R_out, i64 *k_out) {
  if (k_out) *k_out = 0;
  return fgmres_impl(n, rp, ci, v, dinv_in, b, x, rtol, atol, dtol, max_it, m, its_out, rnorm_out, NULL, 0, R_out, k_out);
}

/* ---- synthetic S1 fitted cube (BASELINE config 5): the operands of bench.py's CPU arm ----
 * Same construction as the host generator of the product's bench tooling (iife_b200/synthetic.py:
 * cube_operators) and its device twin (csrc/synth.cu), restated here so that the CPU arm builds its operands
 * without importing the product: P1 stiffness + sigma*mass on the Kuhn triangulation of a (2N)^3 grid from the
 * 8x27 per-cell table `coef` (corner c, neighbour offset d), M = trilinear interpolation from the N^3
 * background grid, b_f = load of f = 1.  Pass 1 (a_rp/m_rp NULL-able value arrays): row lengths; pass 2: fill.
 * tests/test_oracle.py checks it bit for bit against the numpy generator. */
static const int S1_OFFS[15][3] = {
    /* (dx,dy,dz) in the order dz, dy, dx ascending, all components >= 0 or all <= 0 */
    {-1, -1, -1}, {0, -1, -1}, {-1, 0, -1}, {0, 0, -1}, {-1, -1, 0}, {0, -1, 0}, {-1, 0, 0}, {0, 0, 0},
    {1, 0, 0},    {0, 1, 0},   {1, 1, 0},   {0, 0, 1},  {1, 0, 1},   {0, 1, 1},  {1, 1, 1}};

void oracle_synth_cube_lengths(i64 n_bg_cells, i64 *a_len, i64 *m_len) {
  const i64 nv = 2 * n_bg_cells + 1, n_f = nv * nv * nv;
#pragma omp parallel for schedule(static)
  for (i64 j = 0; j < n_f; ++j) {
    const i64 x = j % nv, y = (j / nv) % nv, z = j / (nv * nv);
    int la = 0;
    for (int o = 0; o < 15; ++o) {
      const i64 xx = x + S1_OFFS[o][0], yy = y + S1_OFFS[o][1], zz = z + S1_OFFS[o][2];
      la += (xx >= 0 && yy >= 0 && zz >= 0 && xx < nv && yy < nv && zz < nv);
    }
    a_len[j] = la;
    m_len[j] = (i64)((x & 1) + 1) * ((y & 1) + 1) * ((z & 1) + 1);
  }
}

void oracle_synth_cube_fill(i64 n_bg_cells, const double *coef /* 8 x 27 */, const double *load8, const i64 *a_rp,
                            i32 *a_ci, double *a_v, const i64 *m_rp, i32 *m_ci, double *m_v, double *b_f) {
  const i64 nv = 2 * n_bg_cells + 1, nb = n_bg_cells + 1, n_f = nv * nv * nv, ncell = nv - 1;
#pragma omp parallel for schedule(static)
  for (i64 j = 0; j < n_f; ++j) {
    const i64 x = j % nv, y = (j / nv) % nv, z = j / (nv * nv);
    int cell_ok[8];
    double b = 0.0;
    for (int c = 0; c < 8; ++c) {
      const i64 cx = x - (c & 1), cy = y - ((c >> 1) & 1), cz = z - ((c >> 2) & 1);
      cell_ok[c] = (cx >= 0 && cy >= 0 && cz >= 0 && cx < ncell && cy < ncell && cz < ncell);
      if (cell_ok[c]) b += load8[c];
    }
    b_f[j] = b;
    i64 p = a_rp[j];
    for (int o = 0; o < 15; ++o) {
      const int dx = S1_OFFS[o][0], dy = S1_OFFS[o][1], dz = S1_OFFS[o][2];
      const i64 xx = x + dx, yy = y + dy, zz = z + dz;
      if (!(xx >= 0 && yy >= 0 && zz >= 0 && xx < nv && yy < nv && zz < nv)) continue;
      const int d = (dx + 1) + 3 * (dy + 1) + 9 * (dz + 1);
      double v = 0.0;
      for (int c = 0; c < 8; ++c)
        if (cell_ok[c]) v += coef[c * 27 + d];
      a_ci[p] = (i32)(xx + nv * (yy + nv * zz));
      a_v[p] = v;
      ++p;
    }
    const i64 bx0 = x >> 1, by0 = y >> 1, bz0 = z >> 1;
    const int nx = (int)(x & 1) + 1, ny = (int)(y & 1) + 1, nz = (int)(z & 1) + 1;
    const double w = ((x & 1) ? 0.5 : 1.0) * ((y & 1) ? 0.5 : 1.0) * ((z & 1) ? 0.5 : 1.0);
    i64 q = m_rp[j];
    for (int kz = 0; kz < nz; ++kz)
      for (int ky = 0; ky < ny; ++ky)
        for (int kx = 0; kx < nx; ++kx) {
          m_ci[q] = (i32)((bx0 + kx) + nb * ((by0 + ky) + nb * (bz0 + kz)));
          m_v[q] = w;
          ++q;
        }
  }
}

/* ---- KSPGCR + PCJACOBI (reference common.py:559-560: method='gcr'; PETSc's default restart of 30).  Restated from
 * the published algorithm (src/ksp/ksp/impls/gcr/gcr.c): per step  s = B r;  v = A s;  classical Gram-Schmidt of v
 * (and the same combination of s) against the stored v_i;  normalise;  x += (r, v) s;  r -= (r, v) v;  test the
 * unpreconditioned ||r|| with KSPConvergedDefault against ||b||.  The residual is recomputed from x at the start of
 * every cycle (and tested there).  dinv == NULL means PCNONE. */
int oracle_gcr_jacobi(i64 n, const i64 *rp, const i32 *ci, const double *v, const double *dinv_in, const double *b,
                      double *x, double rtol, double atol, double dtol, i64 max_it, int m, i64 *its_out,
                      double *rnorm_out, double *hist, i64 hist_len) {
  if (m < 1) m = 30;
  double **V = (double **)calloc((size_t)m, sizeof(double *));
  double **S = (double **)calloc((size_t)m, sizeof(double *));
  double *r = (double *)malloc((size_t)(n + 1) * sizeof(double));
  double *coef = (double *)calloc((size_t)m + 1, sizeof(double));
  int reason = 0, first = 1;
  i64 its = 0;
  double rho0 = sqrt(dot(n, b, b)), ttol = fmax(rtol * rho0, atol), res = 0.0;
  while (!reason) {
    oracle_spmv(n, rp, ci, v, x, r);
#pragma omp parallel for schedule(static)
    for (i64 k = 0; k < n; ++k) r[k] = b[k] - r[k];
    res = sqrt(dot(n, r, r));
    if (first) {
      if (hist && hist_len > 0) hist[0] = res;
      if (rho0 == 0.0) {
        rho0 = res;
        ttol = fmax(rtol * rho0, atol);
      }
      first = 0;
    }
    reason = converged_default(res, ttol, atol, dtol, rho0);
    if (!reason && its >= max_it) reason = -3;
    for (int k = 0; k < m && !reason; ++k) {
      if (!V[k]) V[k] = (double *)malloc((size_t)(n + 1) * sizeof(double));
      if (!S[k]) S[k] = (double *)malloc((size_t)(n + 1) * sizeof(double));
      double *vk = V[k], *sk = S[k];
#pragma omp parallel for schedule(static)
      for (i64 t = 0; t < n; ++t) sk[t] = (dinv_in ? dinv_in[t] : 1.0) * r[t];
      oracle_spmv(n, rp, ci, v, sk, vk);
      for (int i = 0; i < k; ++i) coef[i] = dot(n, vk, V[i]);
#pragma omp parallel for schedule(static)
      for (i64 t = 0; t < n; ++t) {
        double a = vk[t], c = sk[t];
        for (int i = 0; i < k; ++i) {
          a -= coef[i] * V[i][t];
          c -= coef[i] * S[i][t];
        }
        vk[t] = a;
        sk[t] = c;
      }
      double rv = dot(n, r, vk), nrm = sqrt(dot(n, vk, vk));
      double scale = 1.0 / nrm, tt = rv / nrm;
#pragma omp parallel for schedule(static)
      for (i64 t = 0; t < n; ++t) {
        vk[t] *= scale;
        sk[t] *= scale;
        x[t] += tt * sk[t];
        r[t] -= tt * vk[t];
      }
      res = sqrt(dot(n, r, r));
      ++its;
      if (hist && its < hist_len) hist[its] = res;
      reason = converged_default(res, ttol, atol, dtol, rho0);
      if (!reason && its >= max_it) reason = -3;
    }
  }
  for (int k = 0; k < m; ++k) {
    free(V[k]);
    free(S[k]);
  }
  free(V);
  free(S);
  free(r);
  free(coef);
  if (its_out) *its_out = its;
  if (rnorm_out) *rnorm_out = res;
  return reason;
}
