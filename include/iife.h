/*
 * iife.h — C ABI of libiife.so, the B200 (sm_100a) extraction linear-algebra library.
 *
 * It replaces, for ONE path, the PETSc routines that the reference
 * (jefromm/interpolation-based-immersed-fea) reaches through petsc4py:
 *
 *   reference call site (file:line, under /root/reference)          entry point(s) here
 *   ---------------------------------------------------------------------------------------------
 *   la_utils.py:178,180  Mat.transpose()            (MatTranspose)  iife_mat_transpose
 *   la_utils.py:179,181  Mat.matMult() x2           (MatMatMult)    iife_ptap_symbolic + iife_ptap_numeric
 *   la_utils.py:165-182  AT_R_A(M, A_f)                             iife_ptap (convenience, cached plan)
 *   la_utils.py:162      Mat.multTranspose()        (MatMultTranspose)  iife_spmv(trans=1)
 *   la_utils.py:141, common.py:139  Mat.mult()      (MatMult)       iife_spmv(trans=0)
 *   common.py:364        Mat.multAdd()              (MatMultAdd)    iife_spmv(alpha=1, beta=1 on a copy)
 *   common.py:554-574, 628-636  KSP create/setUp/solve (KSPCG, KSPFGMRES, PCJACOBI)  iife_ksp_solve
 *   common.py:222,305    Mat.getDiagonal()                          iife_mat_get_diagonal
 *   common.py:483-507    KSP GMRES + computeExtremeSingularValues (estimateConditionNumber)  iife_ksp_solve_hessenberg
 *   common.py:284,327    Mat.zeroRows()  (trimNodes)  (MatZeroRows)  iife_mat_zero_rows
 *   common.py:243-249    A0.setDiagonal(vd); A += A0  (removeZeroDiagonal, getIdentity)  iife_mat_add_diagonal
 *
 * Conventions
 *   - every function returns 0 on success, a positive IIFE_ERR_* code otherwise;
 *     iife_last_error() returns a thread-local, NUL-terminated description.
 *   - the CALLER owns every array passed in; the library never keeps or frees a
 *     caller pointer beyond the call.  The LIBRARY owns all device storage behind
 *     a handle until the matching *_destroy call.
 *   - `mem` says where caller arrays live: IIFE_MEM_HOST (pageable or pinned host
 *     memory; the call copies and returns when results are valid on the host) or
 *     IIFE_MEM_DEVICE (device pointers on the library's device; the call only
 *     enqueues work on the library stream, see iife_set_stream / iife_sync).
 *   - indices are 32- or 64-bit on the caller side (`idx_bytes` = 4 or 8, the two
 *     widths PetscInt can have); values are always fp64.  On the device the
 *     library stores int32 indices: a matrix with nnz >= 2^31-1 is refused with
 *     IIFE_ERR_UNSUPPORTED.
 *   - CSR inputs must have strictly ascending column indices inside each row
 *     (what PETSc AIJ guarantees); iife_mat_create_csr verifies it on the device.
 *   - one process drives one GPU.  Handles are not thread-safe.
 *   - there is no CPU fallback: without a usable CUDA device every compute entry
 *     point fails with IIFE_ERR_NO_DEVICE.
 */
#ifndef IIFE_H
#define IIFE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define IIFE_VERSION 100 /* 0.1.0 */

/* error codes */
#define IIFE_OK 0
#define IIFE_ERR_ARG 1         /* invalid argument / malformed CSR */
#define IIFE_ERR_CUDA 2        /* CUDA runtime error */
#define IIFE_ERR_NOMEM 3       /* device or host allocation failed */
#define IIFE_ERR_UNSUPPORTED 4 /* e.g. nnz does not fit int32 */
#define IIFE_ERR_NO_DEVICE 5   /* iife_init not called / no CUDA device */
#define IIFE_ERR_COMM 6        /* NCCL error */
#define IIFE_ERR_STATE 7       /* handle used in the wrong state (plan/pattern mismatch) */

#define IIFE_MEM_HOST 0
#define IIFE_MEM_DEVICE 1

/* KSP types / preconditioners (common.py:554-574: 'cg' -> KSPCG, 'gmres' -> KSPFGMRES; PC 'jacobi') */
#define IIFE_KSP_CG 0
#define IIFE_KSP_FGMRES 1
#define IIFE_KSP_GCR 2 /* common.py:559-560: 'gcr' -> KSPGCR; `restart` <= 0 selects PETSc's default of 30; single GPU */
#define IIFE_PC_NONE 0
#define IIFE_PC_JACOBI 1

/* converged reasons: numerically identical to PETSc's KSPConvergedReason */
#define IIFE_KSP_CONVERGED_RTOL 2
#define IIFE_KSP_CONVERGED_ATOL 3
#define IIFE_KSP_CONVERGED_ITS 4
#define IIFE_KSP_CONVERGED_ITERATING 0
#define IIFE_KSP_DIVERGED_ITS (-3)
#define IIFE_KSP_DIVERGED_DTOL (-4)
#define IIFE_KSP_DIVERGED_BREAKDOWN (-5)
#define IIFE_KSP_DIVERGED_INDEFINITE_PC (-8)
#define IIFE_KSP_DIVERGED_NANORINF (-9)
#define IIFE_KSP_DIVERGED_INDEFINITE_MAT (-10)

typedef struct iife_mat_s *iife_mat;   /* device-resident CSR (AIJ) matrix */
typedef struct iife_plan_s *iife_plan; /* symbolic PtAP plan (pattern of Mt, A_b, row bins) */
typedef struct iife_halo_s *iife_halo; /* ghost-entry exchange plan of a row-partitioned operator */

/* ---------------------------------------------------------------- context */
int iife_version(void);
const char *iife_last_error(void);
/* select the CUDA device of this process and create the library stream. */
int iife_init(int device);
int iife_finalize(void);
int iife_device_count(int *n);
/* run on a caller stream (a cudaStream_t, e.g. torch's current stream); NULL = library stream */
int iife_set_stream(void *cuda_stream);
int iife_sync(void);
/* bytes currently held by the library on the device */
int iife_device_bytes(int64_t *bytes);
/* number of kernels launched by the library since the last reset (bench.py "gpu_launches") */
int iife_launch_count(int64_t *n, int reset);

/* ---------------------------------------------------------------- matrices */
/* val may be NULL (pattern only; values zero until iife_mat_update_values). */
int iife_mat_create_csr(int64_t n_rows, int64_t n_cols, const void *rowptr, const void *colind,
                        const double *val, int idx_bytes, int mem, iife_mat *out);
/* flags: IIFE_CSR_UNSORTED_OK accepts rows whose columns are not ascending (the [owned | ghost]
 * renumbered operator block of the row-partitioned solver; SpMV and Jacobi do not need the order) */
#define IIFE_CSR_UNSORTED_OK 1
int iife_mat_create_csr_ex(int64_t n_rows, int64_t n_cols, const void *rowptr, const void *colind,
                           const double *val, int idx_bytes, int mem, int flags, iife_mat *out);
/* new values on the same pattern (the per-Newton-step path, common.py:432-435) */
int iife_mat_update_values(iife_mat A, const double *val, int mem);
int iife_mat_get_info(iife_mat A, int64_t *n_rows, int64_t *n_cols, int64_t *nnz);
/* copy out; any of rowptr / colind / val may be NULL */
int iife_mat_get_csr(iife_mat A, void *rowptr, void *colind, double *val, int idx_bytes, int mem);
/* raw device pointers (int32 rowptr[n_rows+1], int32 colind[nnz], double val[nnz]); valid until destroy.
 * A caller that WRITES values through `val` must call iife_mat_touch afterwards: plans and operator copies that
 * were precomputed from the old values (transpose, SELL copy, Jacobi diagonal, PtAP templates) key on it. */
int iife_mat_device_ptrs(iife_mat A, void **rowptr, void **colind, void **val);
int iife_mat_touch(iife_mat A);
/* 64-bit fingerprint of (shape, rowptr, colind): the key of the symbolic-plan cache */
int iife_mat_fingerprint(iife_mat A, uint64_t *fp);
/* explicit transpose as a new matrix, rows column-sorted (MatTranspose, la_utils.py:178) */
int iife_mat_transpose(iife_mat A, iife_mat *out);
/* diag[i] = A[i,i], 0 where the diagonal entry is not stored (MatGetDiagonal) */
int iife_mat_get_diagonal(iife_mat A, double *diag, int mem);
/* MatZeroRows as trimNodes uses it (common.py:284,327; no KEEP_NONZERO_PATTERN): *out = copy of A in which
 * every listed row (idx_bytes-wide 0-based indices, duplicates allowed) holds exactly one entry (i,i) = diag
 * when diag != 0 and i < n_cols, and nothing otherwise.  A is unchanged; the caller swaps handles. */
int iife_mat_zero_rows(iife_mat A, const void *rows, int64_t n_listed, int idx_bytes, double diag, int mem,
                       iife_mat *out);
/* *out = A + diag(d) with pattern union(A, full diagonal), rows column-sorted: the `A += A0` of
 * removeZeroDiagonal (common.py:243-249; A0 = MatDiagonalSet(vd) on an empty matrix). d has n_rows entries. */
int iife_mat_add_diagonal(iife_mat A, const double *d, int mem, iife_mat *out);
int iife_mat_destroy(iife_mat A);

/* ---------------------------------------------------------------- SpMV */
/* y = alpha * op(A) x + beta * y, op = A (trans=0) or A^T (trans=1, through a cached explicit
 * transpose: deterministic, no atomics).  beta == 0 ignores the incoming y (may be uninitialised). */
int iife_spmv(iife_mat A, int trans, double alpha, const double *x, double beta, double *y, int mem);

/* ---------------------------------------------------------------- PtAP: A_b = M^T A_f M */
/* symbolic phase: transpose of M, structural pattern of (M^T A) M — bit-exact to the boolean
 * product PETSc's MatMatMult/MatPtAP symbolic phases produce — row bins for the numeric phase. */
int iife_ptap_symbolic(iife_mat M, iife_mat A, iife_plan *out);
/* pattern-compatibility check of (M, A) against a plan (fingerprints) */
int iife_plan_matches(iife_plan P, iife_mat M, iife_mat A, int *matches);
int iife_plan_get_info(iife_plan P, int64_t *n_b, int64_t *nnz_c, int64_t *nnz_inter);
/* rows per numeric kernel: [0] warp hashing 256/64, [1] warp hashing 1K/256, [2] CTA hashing 2K/1K, [3] CTA hashing
 * 8K/4K, [4] CTA hashing with global-memory tables, [5] slot plan 128/32, [6] slot plan 256/256 (tests use it to
 * prove that every kernel of the ladder is exercised) */
int iife_plan_bin_counts(iife_plan P, int64_t *counts7);
/* the same for the first n bins; [7] = slot plan for wide rows (intermediate row <= 2040, output row <= 512 entries, two
 * bytes per product term: quadratic / 3-D unfitted backgrounds) */
int iife_plan_bin_counts_n(iife_plan P, int64_t *counts, int n);
/* template plan of the numeric phase (built on the first numeric call, rebuilt when the values of M change):
 * rows of the slot-plan bins that share structure and M values with at least IIFE_TPL_MIN_ROWS (32) others run one
 * precompiled gather program per group.  n_templates groups cover n_rows rows in n_chunks work items;
 * lane_use2[0..1] = mean fraction of busy lanes in the two gather stages.  All zero before the first numeric call. */
int iife_plan_tpl_info(iife_plan P, int64_t *n_templates, int64_t *n_rows, int64_t *n_chunks, double *lane_use2);
/* host-only self-check hook of the template compiler (no device needed): compiles the gather program of one output
 * row given as raw description — len1[n0] operand row lengths of A_f, w[n0] = R[i,:], slot1 = destination of every
 * stage-1 term, len2[n1] lengths of the M rows, mval / slot2 per stage-2 term — and interprets it on the CPU with
 * a_vals = the operand rows of A_f back to back; c_out[n2] receives the output row.  info10: S1, S2, extra slots of
 * both stages, staging steps, program bytes, lane use x1000 of both stages, shared-memory conflict degree x1000 of
 * the gather reads of both stages (1000 = conflict free). */
int iife_tpl_emulate_row(int n0, const int *len1, const double *w, const unsigned char *slot1, int n1, const int *len2,
                         const double *mval, const unsigned char *slot2, int n2, const double *a_vals, double *c_out,
                         int *info10);
/* numeric phase; *C == NULL creates the result matrix, otherwise refills its values (reuse) */
int iife_ptap_numeric(iife_plan P, iife_mat M, iife_mat A, iife_mat *C);
/* general triple product C = R A P with an explicit restriction R (n_out x nJ), A (nJ x nK), P (nK x n_cols):
 * what one rank of the row-partitioned PtAP computes after gathering its block of M^T, the A_f rows
 * that block touches and the M rows those touch (MPIAIJ MatMatMult fetches off-process rows the same way) */
int iife_rap_symbolic(iife_mat R, iife_mat A, iife_mat P, iife_plan *out);
int iife_rap_numeric(iife_plan plan, iife_mat R, iife_mat A, iife_mat P, iife_mat *C);
/* synchronises and reports a numeric-phase inconsistency (operands whose pattern differs from the
 * plan's: a product term found no slot) as IIFE_ERR_STATE */
int iife_plan_check(iife_plan P);
int iife_plan_destroy(iife_plan P);
/* convenience used by la_utils.AT_R_A: looks the plan up in an internal LRU cache keyed by the
 * two pattern fingerprints, builds it on a miss, runs numeric, returns a NEW matrix. */
int iife_ptap(iife_mat M, iife_mat A, iife_mat *C, int *plan_was_cached);
int iife_plan_cache_clear(void);

/* ---------------------------------------------------------------- KSP */
typedef struct iife_ksp_result {
  int64_t iterations;
  int32_t reason;     /* IIFE_KSP_* */
  int32_t _pad;
  double rnorm;       /* final residual norm in the KSP's norm (CG: preconditioned, FGMRES: true) */
  double rnorm0;      /* reference norm of the relative test (norm of the (preconditioned) rhs) */
} iife_ksp_result;

/* Solve A x = b.  x holds the initial guess on entry (nonzero_initial_guess=True, common.py:634)
 * and the solution on return.  hist (optional, host memory, hist_len entries) receives the residual
 * norm of iterations 0..min(its, hist_len-1).  Never fails on non-convergence (common.py:635):
 * the outcome is in res->reason.  `halo` is NULL for a single-GPU operator. */
int iife_ksp_solve(iife_mat A, int ksp_type, int pc_type, double rtol, double atol, double dtol,
                   int64_t max_it, int restart, const double *b, double *x, int mem,
                   iife_halo halo, iife_ksp_result *res, double *hist, int64_t hist_len);
/* FGMRES solve (single GPU) that also returns the k x k upper-triangular factor R (column-major, leading
 * dimension k, host memory, r_capacity doubles available) of the last cycle's Hessenberg matrix; *k_out = k.
 * The Givens rotations are orthogonal, so R has the singular values PETSc's
 * KSPComputeExtremeSingularValues reports for that cycle: estimateConditionNumber, common.py:483-507
 * (GMRES restart 1000, PC none).  pc_type NONE reproduces PETSc's Hessenberg; JACOBI is the right-preconditioned
 * operator A D^-1 (PETSc would use D^-1 A). */
int iife_ksp_solve_hessenberg(iife_mat A, int pc_type, double rtol, double atol, double dtol, int64_t max_it,
                              int restart, const double *b, double *x, int mem, iife_ksp_result *res, double *R,
                              int64_t r_capacity, int64_t *k_out);

/* ---------------------------------------------------------------- multi-GPU (one process per GPU) */
/* 128-byte NCCL unique id, created on rank 0 and broadcast by the host framework (torch.distributed) */
int iife_comm_unique_id(void *id128);
int iife_comm_init(int rank, int nranks, const void *id128);
int iife_comm_finalize(void);
int iife_comm_info(int *rank, int *nranks);
/* Halo plan of a row-partitioned square operator whose local column space is
 * [owned 0..n_owned) ++ [ghost n_owned..n_owned+n_ghost): for each peer, how many ghost entries
 * come from it (recv_counts, contiguous in ghost order, peers ascending) and which owned entries go
 * to it (send_idx, concatenated in peer order with send_counts).  Arrays are host memory. */
int iife_halo_create(int64_t n_owned, int64_t n_ghost, const int64_t *send_counts,
                     const int32_t *send_idx, const int64_t *recv_counts, iife_halo *out);
/* fill x[n_owned .. n_owned+n_ghost) from the peers' owned entries (device pointer) */
int iife_halo_exchange(iife_halo H, double *x_dev);
int iife_halo_destroy(iife_halo H);
/* NVLink peer-memory path of the row-partitioned solver (p2p.cu): export allocates the halo's peer-visible
 * vector + mailbox and returns two 64-byte cudaIpc handles; the host framework all-gathers the handles of
 * all ranks (rank order, 128 bytes each) and tells every rank where its send block starts inside each
 * peer's vector (dst_start[q] = n_owned_q + offset of this rank's block in q's ghost section).  After
 * attach, iife_ksp_solve_dist exchanges ghosts and reduces dot products with library kernels that store
 * into peer memory, without NCCL calls inside the iteration. */
int iife_halo_p2p_export(iife_halo H, void *handles128);
int iife_halo_p2p_attach(iife_halo H, const void *all_handles, const int64_t *dst_start);
int iife_halo_p2p_error(iife_halo H, int *err);
/* row-partitioned y_local = A_local * [x_owned; x_ghost] (halo exchange + SpMV), device pointers */
int iife_spmv_dist(iife_mat A_local, iife_halo H, double *x_dev, double *y_dev);
/* row-partitioned KSP: A_local is n_owned x (n_owned + n_ghost), b and x are device vectors of the
 * owned entries; halo exchange before every SpMV, NCCL allreduce for every reduction. */
int iife_ksp_solve_dist(iife_mat A_local, iife_halo H, int ksp_type, int pc_type, double rtol, double atol,
                        double dtol, int64_t max_it, int restart, const double *b_dev, double *x_dev,
                        iife_ksp_result *res, double *hist, int64_t hist_len);
/* sum-allreduce of n fp64 on the library communicator (device pointer, in place) */
int iife_allreduce_sum(double *buf_dev, int64_t n);
/* exchange of variable-size byte blocks between ranks (send/recv displacements in bytes, host
 * arrays of nranks+1 entries); buffers are device pointers.  Used for ghost rows of A_f / M. */
int iife_alltoallv_bytes(const void *send_dev, const int64_t *send_displs, void *recv_dev,
                         const int64_t *recv_displs);

/* ---------------------------------------------------------------- synthetic workload (bench / tests) */
/* BASELINE config 5, "S1 fitted cube" (SURVEY.md §8d): background N_b^3 trilinear B-spline grid,
 * foreground = 2x refinement split into Kuhn tetrahedra, A_f = K + sigma*Mass (P1), M = trilinear
 * interpolation, b_f = load of f=1.  Rows [row_begin,row_end) of the foreground operators are
 * generated directly on the device (global column ids).  coef is the 8x27 table of per-cell
 * contributions (cell corner c in {0,1}^3 relative to the vertex, neighbour offset d in {-1,0,1}^3),
 * load8[c] the per-cell load contribution; both are produced by iife_b200.synthetic on the host. */
int iife_synth_cube_counts(int64_t n_bg_cells, int64_t row_begin, int64_t row_end,
                           int64_t *nnz_A, int64_t *nnz_M);
int iife_synth_cube_build(int64_t n_bg_cells, int64_t row_begin, int64_t row_end,
                          const double *coef_8x27, const double *load8, iife_mat *A_f, iife_mat *M,
                          double *b_f_dev);

#ifdef __cplusplus
}
#endif
#endif /* IIFE_H */
