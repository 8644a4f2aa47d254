// common.cuh — context, error plumbing and small device helpers shared by all translation units
// of libiife.so.  Hand-written CUDA for sm_100a; no torch types, no CPU fallback.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>
#include <string>
#include <vector>
#include "../../include/iife.h"

namespace iife {

struct Ctx {
  bool init = false;
  int device = -1;
  int sm_count = 148;
  int max_smem_optin = 0;
  cudaStream_t stream = nullptr;      // stream all work is enqueued on
  cudaStream_t own_stream = nullptr;  // created by iife_init
  int64_t dev_bytes = 0;
  int64_t launches = 0;
  // NCCL (comm.cu)
  void *nccl_comm = nullptr;
  int rank = 0, nranks = 1;
};
Ctx &ctx();

int set_err(int code, const char *fmt, ...);
const char *get_err();

#define IIFE_CUDA(expr)                                                                         \
  do {                                                                                          \
    cudaError_t _e = (expr);                                                                    \
    if (_e != cudaSuccess)                                                                      \
      return ::iife::set_err(_e == cudaErrorMemoryAllocation ? IIFE_ERR_NOMEM : IIFE_ERR_CUDA, \
                             "%s:%d: %s: %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
  } while (0)

#define IIFE_TRY(expr)        \
  do {                        \
    int _rc = (expr);         \
    if (_rc != IIFE_OK) return _rc; \
  } while (0)

#define IIFE_NEED_INIT()                                                                   \
  do {                                                                                     \
    if (!::iife::ctx().init)                                                               \
      return ::iife::set_err(IIFE_ERR_NO_DEVICE, "iife_init() has not been called (no CUDA device bound)"); \
  } while (0)

// launch bookkeeping: every kernel launch of the library goes through this macro so that
// iife_launch_count() is an honest count.
#define IIFE_LAUNCH(kernel, grid, block, smem, ...)                          \
  do {                                                                       \
    kernel<<<(grid), (block), (smem), ::iife::ctx().stream>>>(__VA_ARGS__);  \
    ::iife::ctx().launches++;                                                \
  } while (0)

#define IIFE_CHECK_LAUNCH() IIFE_CUDA(cudaGetLastError())

// device allocation with accounting
int dev_alloc(void **p, size_t bytes);
int dev_free(void *p, size_t bytes);
void dev_release_cached();  // hand the allocator's free blocks back to the driver
template <class T>
inline int dev_alloc_t(T **p, size_t n) {
  return dev_alloc((void **)p, (n ? n : 1) * sizeof(T));
}
template <class T>
inline int dev_free_t(T *p, size_t n) {
  return dev_free((void *)p, (n ? n : 1) * sizeof(T));
}

// RAII temporary device buffer (freed on scope exit; stream-ordered with cudaFreeAsync)
template <class T>
struct Tmp {
  T *p = nullptr;
  size_t n = 0;
  int alloc(size_t count) {
    n = count;
    return dev_alloc_t(&p, n);
  }
  ~Tmp() {
    if (p) dev_free_t(p, n);
  }
  Tmp() = default;
  Tmp(const Tmp &) = delete;
  Tmp &operator=(const Tmp &) = delete;
};

struct Mat {
  int64_t n_rows = 0, n_cols = 0, nnz = 0;
  int *rowptr = nullptr;  // [n_rows+1]
  int *colind = nullptr;  // [nnz]
  double *val = nullptr;  // [nnz]
  // cached explicit transpose (built on first use) and the gather permutation of its values
  Mat *T = nullptr;
  int *T_perm = nullptr;
  bool T_vals_valid = false;
  // cached inverse Jacobi diagonal (zero -> 1), invalidated when values change
  double *dinv = nullptr;
  bool dinv_valid = false;
  uint64_t fp = 0;
  bool fp_valid = false;
  int max_row_len = -1;  // lazily computed (mat_max_row_len)
  int max_tile_entries = -1;  // with it: most entries in any 128 consecutive rows starting at a multiple of 128 (k_spmv_stream)
  // identity of the VALUES: uid is unique per matrix object, val_version counts iife_mat_update_values calls
  // (plans that precompute from an operand's values — ptap_prog.cuh — key on the pair)
  uint64_t uid = 0, val_version = 0;
  // SELL-32 copy used by the KSP operator (spmv.cu): slices of 32 rows, column-major inside a slice
  int sell_state = 0;  // 0 not tried, 1 built, -1 rejected (padding too large)
  int64_t sell_slices = 0, sell_padded = 0;
  int *sell_ptr = nullptr;  // [sell_slices+1] entry offsets
  int *sell_cptr = nullptr;  // [sell_slices+1] offsets into sell_col (compact column words)
  int64_t sell_cwords = 0;
  int *sell_col = nullptr;   // per slice: `width` shared offsets (uniform slice) or width*32 explicit columns
  double *sell_val = nullptr;
  bool sell_vals_valid = false;
  // row-partitioned solver: slice order "interior first" for the halo-fused SpMV
  int64_t sell_n_interior = 0, sell_int_lo = -1, sell_order_owned = -1;  // mat_ensure_sell_order
};

int mat_alloc(Mat **out, int64_t n_rows, int64_t n_cols, int64_t nnz);
int mat_free(Mat *A);
int mat_ensure_transpose(Mat *A);  // builds A->T (+values)
int mat_ensure_dinv(Mat *A);       // Jacobi inverse diagonal with zero -> 1
int mat_fingerprint(Mat *A, uint64_t *fp);
int mat_max_row_len(Mat *A, int *out);

// exclusive scan of int32 counts into int32 offsets (n+1 outputs: out[n] = total); total64 returned
// on the host (synchronises the stream).
int exclusive_scan_i32(const int *in, int *out, int64_t n, int64_t *total64);
int exclusive_scan_i32_i64(const int *in, long long *out, int64_t n, int64_t *total64);

// SpMV launchers (spmv.cu)
int spmv_launch(const Mat *A, double alpha, const double *x, double beta, double *y);
// w = A p and *dot_out = sum_i p_i w_i over the rows of A (partials + deterministic last-block
// reduce); if flag != nullptr and *flag != 0 the kernel is a no-op.
int spmv_dot_launch(const Mat *A, const double *p, double *w, double *dot_out, double *partials,
                    unsigned int *counter, const int *flag, const struct P2PRed *red = nullptr,
                    const struct HaloWait *hw = nullptr);
bool mat_sell_ready(const Mat *A);
int spmv_pick_lpr(const Mat *A);
// SELL-32 operator copy: builds it on first use (returns IIFE_OK with A->sell_state == -1 if the
// padding would exceed 1.25x nnz, in which case callers stay on CSR) and refreshes values.
int mat_ensure_sell(Mat *A, bool want_dinv = false);  // want_dinv: the fill pass also writes A->dinv (PCJACOBI)
int mat_ensure_sell_order(Mat *A, int64_t n_owned);
int spmv_launch_signature();       // everything env-selected that shapes the SpMV launches (key of cached graphs)
void ksp_release_cached_graphs();  // ksp.cu: captured CG chunks kept between solves
void mat_free_sell(Mat *A);

// ghost-entry exchange plan of a row-partitioned operator (comm.cu)
struct Halo {
  int64_t n_owned = 0, n_ghost = 0;
  int nranks = 1;
  std::vector<int64_t> send_counts, recv_counts, send_off, recv_off;
  int *send_idx = nullptr;     // device [total_send]
  double *send_buf = nullptr;  // device [total_send]
  int64_t total_send = 0;
  // ---- peer-memory path (p2p.cu): ghost entries are stored directly into the neighbours' vectors
  // and the dot products are reduced through per-rank mailboxes, all over NVLink peer mappings
  bool p2p = false;
  int me = 0;
  double *xbuf = nullptr;            // [n_owned + n_ghost], IPC-exported: the multiplied vector of the solver
  struct Mailbox *mbox = nullptr;    // IPC-exported
  double *peer_xbuf[16] = {nullptr};
  struct Mailbox *peer_mbox[16] = {nullptr};
  long long dst_start[16] = {0};     // where my block starts inside peer q's xbuf
  unsigned char *send_peer = nullptr;  // device [total_send]: destination rank of every send entry
  int *send_off_dev = nullptr;         // device [nranks+1]
  unsigned long long *dev_seq = nullptr;  // device [3]: halo / allreduce sequence numbers, CG iteration counter
  unsigned int *p2p_counter = nullptr;    // device [1]
  int *p2p_err = nullptr;                 // device [1]
  unsigned int send_mask = 0, recv_mask = 0;
  // push plan by OWNED ROW (three-kernel CG iteration, ksp.cu): the thread that updates p[i] also stores it into the
  // neighbours' ghost slots.  brow: sorted owned rows that are sent anywhere; entries bptr[k]..bptr[k+1] of
  // (bpeer, bdst) say where; bmask: per 32 owned rows the pair (bits of the rows in brow, number of brow entries before)
  int n_brow = 0;
  int *brow = nullptr, *bptr = nullptr;
  unsigned char *bpeer = nullptr;
  long long *bdst = nullptr;
  unsigned int *bmask = nullptr;
};

constexpr int P2P_MAX_RANKS = 16;
struct Mailbox {
  double ar_vals[2][P2P_MAX_RANKS][4];
  unsigned long long ar_flag[2][P2P_MAX_RANKS];
  unsigned long long halo_flag[P2P_MAX_RANKS];
  // reductions fused into the CG kernels (two per iteration, slot = sequence parity)
  double it_vals[2][P2P_MAX_RANKS][4];
  unsigned long long it_flag[2][P2P_MAX_RANKS];
  // the same reductions in "low latency" form: every double travels as two 8-byte words (half of the value | low 32
  // bits of the sequence number), each store atomic and self-validating: no fence and no separate flag
  unsigned long long it_ll[2][P2P_MAX_RANKS][8];
};
// enqueue on the library stream: push ghost entries of H->xbuf to the peers and wait for mine
int p2p_halo_exchange(Halo *H, const int *reason_flag);
// in-place sum over ranks of vals[0..n) (n <= 4) through the mailboxes; mode 0: plain, 1: then the CG
// update scalar step (ksp.cu supplies the kernel through p2p_set_cg_scalars)
int p2p_allreduce(Halo *H, double *vals, int n, const int *reason_flag);
int halo_exchange(Halo *H, double *x_dev);     // fills x[n_owned .. n_owned+n_ghost)
int allreduce_sum(double *buf_dev, int64_t n);  // in place, on the library stream

static inline int div_up(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }

// ---------------------------------------------------------------- device helpers
#ifdef __CUDACC__
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
  return v;
}
// sum over the lanes of a power-of-two sub-group of the warp (xor butterfly: all lanes get it)
template <int G>
__device__ __forceinline__ double group_sum(double v) {
#pragma unroll
  for (int o = G / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
// block-wide sum (blockDim.x multiple of 32, <= 1024); result valid in thread 0
__device__ __forceinline__ double block_sum(double v, double *smem32) {
  v = warp_sum(v);
  int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  __syncthreads();
  if (lane == 0) smem32[w] = v;
  __syncthreads();
  if (w == 0) {
    int nw = (blockDim.x + 31) >> 5;
    v = lane < nw ? smem32[lane] : 0.0;
    v = warp_sum(v);
  }
  return v;
}
#endif

}  // namespace iife
