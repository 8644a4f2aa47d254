// ksp.cu — device-resident Krylov solvers: Jacobi-preconditioned CG and FGMRES(m).
//
// Replaces the Krylov branch of the reference's solveKSP (common.py:554-574 + 628-636), i.e. PETSc's
// KSPCG / KSPFGMRES + PCJACOBI as configured there, with PETSc's defaults restated in SURVEY.md A.5-A.8:
//   CG      left preconditioning, preconditioned residual norm, test against max(rtol*||D^-1 b||, atol)
//   FGMRES  right (flexible) preconditioning, true residual norm, classical Gram-Schmidt, no refinement,
//           Givens-updated residual estimate, restart m, test against max(rtol*||b||, atol)
//   Jacobi  z = r / diag(A), zero diagonal entries replaced by 1
//
// All scalars of the iteration (alpha, beta, norms, the Hessenberg matrix, Givens rotations, the
// iteration counter and the converged reason) live in device memory and are produced by the *last
// CTA to finish* of the kernel that reduces them, in a fixed order (bit-reproducible).  The host only
// enqueues kernels — as a CUDA graph of a chunk of iterations for CG — and polls the reason flag once
// per chunk; after convergence the remaining kernels of the chunk are no-ops.  There is no host
// round trip per iteration.
#include "common.cuh"
#include "p2p_dev.cuh"
#include <math.h>
#include <algorithm>
#include <stdlib.h>
#include <chrono>

namespace iife {

// scalar slots
enum {
  S_BETA = 0, S_BETA_OLD, S_DELTA, S_DP, S_TTOL, S_RHO0, S_RTOL, S_ATOL, S_DTOL, S_SCALE, S_RES, S_TT,
  S_RAW = 12,  // 3 raw (rank-local) sums awaiting the allreduce in the row-partitioned solver
  S_BETA2 = 16,  // three-kernel CG iteration: (z_k, r_k) of iteration k in slot k & 1
  S_COUNT = 20
};
// flag slots
enum { F_REASON = 0, F_ITS, F_LOC_IT, F_MAXIT, F_HAPEND, F_COUNT = 8 };

constexpr int VEC_THREADS = 256;
#ifndef IIFE_VEC_ILP
#define IIFE_VEC_ILP 4
#endif
constexpr int VEC_ILP = IIFE_VEC_ILP;  // elements per thread and round in the CG vector kernels
constexpr int MAX_PARTIALS = 2048;

struct KspWork {
  double *sc = nullptr;         // [S_COUNT]
  int *fl = nullptr;            // [F_COUNT]
  double *partials = nullptr;   // [4 * MAX_PARTIALS] (+ multi-dot partials allocated separately)
  unsigned int *counters = nullptr;  // [4]
  double *hist = nullptr;       // device residual history [hist_len]
  int64_t hist_len = 0;
};

__device__ __forceinline__ void log_hist(double *hist, long long hist_len, int its, double v) {
  if (hist && its < hist_len) hist[its] = v;
}

// KSPConvergedDefault evaluated by one thread
__device__ __forceinline__ void converged_default(double *sc, int *fl, int its, double rnorm) {
  if (fl[F_REASON] != 0) return;
  if (isnan(rnorm) || isinf(rnorm)) {
    fl[F_REASON] = IIFE_KSP_DIVERGED_NANORINF;
  } else if (rnorm <= sc[S_TTOL]) {
    fl[F_REASON] = (rnorm < sc[S_ATOL]) ? IIFE_KSP_CONVERGED_ATOL : IIFE_KSP_CONVERGED_RTOL;
  } else if (rnorm >= sc[S_DTOL] * sc[S_RHO0]) {
    fl[F_REASON] = IIFE_KSP_DIVERGED_DTOL;
  }
}

// ---- generic "last CTA reduces" epilogue: returns true in ALL threads of the last block, with the
// NR reduced sums available in out[] (shared).  partials layout: [r * MAX_PARTIALS + block].
template <int NR>
__device__ __forceinline__ bool grid_reduce(double (&acc)[NR], double *partials, unsigned int *counter, double *out_sh,
                                            double *red_sh, bool *last_sh) {
#pragma unroll
  for (int r = 0; r < NR; ++r) {
    double b = block_sum(acc[r], red_sh);
    if (threadIdx.x == 0) partials[r * MAX_PARTIALS + blockIdx.x] = b;
  }
  if (threadIdx.x == 0) {
    __threadfence();
    unsigned int t = atomicAdd(counter, 1u);
    *last_sh = (t == gridDim.x - 1);
  }
  __syncthreads();
  if (!*last_sh) return false;
  __threadfence();
#pragma unroll
  for (int r = 0; r < NR; ++r) {
    double s = 0.0;
    for (int k = threadIdx.x; k < (int)gridDim.x; k += blockDim.x) s += __ldcg(partials + r * MAX_PARTIALS + k);
    s = block_sum(s, red_sh);
    if (threadIdx.x == 0) out_sh[r] = s;
  }
  if (threadIdx.x == 0) *counter = 0u;
  __syncthreads();
  return true;
}

// ------------------------------------------------------------------------------------------------
// CG kernels
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void cg_init_scalars(double *sc, int *fl, double zz, double zr, double zbzb, double *hist,
                                                long long hist_len) {
  double dp = sqrt(zz), beta = zr, rho0 = sqrt(zbzb);
  if (rho0 == 0.0) rho0 = dp;  // KSPConvergedDefault: zero right-hand side with a nonzero guess -> initial residual norm
  sc[S_DP] = dp;
  sc[S_BETA] = beta;
  sc[S_BETA_OLD] = beta;
  sc[S_RHO0] = rho0;
  sc[S_TTOL] = fmax(sc[S_RTOL] * rho0, sc[S_ATOL]);
  fl[F_ITS] = 0;
  log_hist(hist, hist_len, 0, dp);
  converged_default(sc, fl, 0, dp);
  if (fl[F_REASON] == 0) {
    // checks PETSc makes at the top of iteration i (KSPSolve_CG): its is already i+1 there
    if (beta == 0.0) { fl[F_REASON] = IIFE_KSP_CONVERGED_ATOL; fl[F_ITS] = 1; }
    else if (beta < 0.0) { fl[F_REASON] = IIFE_KSP_DIVERGED_INDEFINITE_PC; fl[F_ITS] = 1; }
    else if (isnan(beta) || isinf(beta)) { fl[F_REASON] = IIFE_KSP_DIVERGED_NANORINF; fl[F_ITS] = 1; }
    else if (fl[F_MAXIT] <= 0) fl[F_REASON] = IIFE_KSP_DIVERGED_ITS;
  }
}

__device__ __forceinline__ void cg_update_scalars(double *sc, int *fl, double zr, double zz, double *hist,
                                                  long long hist_len) {
  double beta = zr, dp = sqrt(zz);
  int its = fl[F_ITS] + 1;
  sc[S_BETA_OLD] = sc[S_BETA];
  sc[S_BETA] = beta;
  sc[S_DP] = dp;
  fl[F_ITS] = its;
  log_hist(hist, hist_len, its, dp);
  converged_default(sc, fl, its, dp);
  if (fl[F_REASON] == 0) {
    if (its >= fl[F_MAXIT]) fl[F_REASON] = IIFE_KSP_DIVERGED_ITS;
    else if (beta == 0.0) { fl[F_REASON] = IIFE_KSP_CONVERGED_ATOL; fl[F_ITS] = its + 1; }
    else if (beta < 0.0) { fl[F_REASON] = IIFE_KSP_DIVERGED_INDEFINITE_PC; fl[F_ITS] = its + 1; }
    else if (isnan(beta) || isinf(beta)) { fl[F_REASON] = IIFE_KSP_DIVERGED_NANORINF; fl[F_ITS] = its + 1; }
  }
}

// after r = b - A x0:  dp = ||D^-1 r||, beta = (D^-1 r, r), rho0 = ||D^-1 b||, test(0)
template <bool DIST>
__global__ void __launch_bounds__(VEC_THREADS)
k_cg_init(const double *__restrict__ r, const double *__restrict__ b, const double *__restrict__ dinv, int64_t n,
          double *sc, int *fl, double *partials, unsigned int *counter, double *hist, long long hist_len) {
  __shared__ double red[32];
  __shared__ double out[3];
  __shared__ bool last;
  double acc[3] = {0.0, 0.0, 0.0};
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    double d = dinv ? dinv[i] : 1.0;
    double ri = r[i], z = d * ri, zb = d * b[i];
    acc[0] = fma(z, z, acc[0]);
    acc[1] = fma(z, ri, acc[1]);
    acc[2] = fma(zb, zb, acc[2]);
  }
  if (grid_reduce<3>(acc, partials, counter, out, red, &last) && threadIdx.x == 0) {
    if (DIST) {
      sc[S_RAW + 0] = out[0];
      sc[S_RAW + 1] = out[1];
      sc[S_RAW + 2] = out[2];
    } else {
      cg_init_scalars(sc, fl, out[0], out[1], out[2], hist, hist_len);
    }
  }
}

__global__ void k_cg_init_scalars(double *sc, int *fl, double *hist, long long hist_len) {
  cg_init_scalars(sc, fl, sc[S_RAW + 0], sc[S_RAW + 1], sc[S_RAW + 2], hist, hist_len);
}

// p = z (first iteration) or p = z + (beta/beta_old) p, with z = D^-1 r recomputed (never stored)
__global__ void __launch_bounds__(VEC_THREADS)
k_cg_p(const double *__restrict__ r, const double *__restrict__ dinv, double *__restrict__ p, int64_t n,
       const double *__restrict__ sc, const int *__restrict__ fl) {
  if (fl[F_REASON] != 0) return;
  bool first = (fl[F_ITS] == 0);
  double bb = first ? 0.0 : sc[S_BETA] / sc[S_BETA_OLD];
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  // VEC_ILP elements per thread and round, all loads issued before the first use (a thread owns ~10 elements: one
  // element per round leaves one batch of loads in flight and the kernel on the latency, not the bandwidth)
  for (int64_t i0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i0 < n; i0 += VEC_ILP * stride) {
    double rv[VEC_ILP], dv[VEC_ILP], pv[VEC_ILP];
#pragma unroll
    for (int u = 0; u < VEC_ILP; ++u) {
      const int64_t i = i0 + u * stride;
      const bool in = i < n;
      rv[u] = in ? r[i] : 0.0;
      dv[u] = (in && dinv) ? dinv[i] : 1.0;
      pv[u] = (in && !first) ? p[i] : 0.0;
    }
#pragma unroll
    for (int u = 0; u < VEC_ILP; ++u) {
      const int64_t i = i0 + u * stride;
      const double z = dv[u] * rv[u];
      if (i < n) p[i] = first ? z : fma(bb, pv[u], z);
    }
  }
}

// alpha = beta/delta; x += alpha p; r -= alpha w; z = D^-1 r; (z,r), (z,z) -> beta, dp, test(i+1)
// MODE 0: single GPU (scalar step by the last CTA).  MODE 1: row-partitioned with NCCL (raw sums out).
// MODE 2: row-partitioned over peer memory: delta is summed from the ranks' partials waiting in the
// mailbox (prologue), the two new partial sums are pushed to every rank by the last CTA (epilogue).
// MODE 3: as 2 inside the three-kernel iteration (k_cg_p_push does the scalar step): beta of this iteration is in
// the slot S_BETA2 + (iteration & 1), `it_rel` = iteration index since the start of the solve.
template <int MODE>
__global__ void __launch_bounds__(VEC_THREADS)
k_cg_update(double *__restrict__ x, double *__restrict__ r, const double *__restrict__ p,
            const double *__restrict__ w, const double *__restrict__ dinv, int64_t n, double *sc, int *fl,
            double *partials, unsigned int *counter, double *hist, long long hist_len, P2PRed pr) {
  if (fl[F_REASON] != 0) return;
  __shared__ double red[32];
  __shared__ double out[2];
  __shared__ bool last;
  double delta;
  if (MODE == 2 || MODE == 3) {
    __shared__ double s_delta;
    if (MODE == 3 && blockIdx.x == 0 && threadIdx.x == 0) cg_trace(pr, TR_U_IN);
    if (threadIdx.x < 32) {
      double d;
      p2p_wait_sum(pr, 2ull * (*pr.iter + (unsigned long long)pr.k_off) + 1ull, &d, 1);
      if (threadIdx.x == 0) {
        s_delta = d;
        if (blockIdx.x == 0) sc[S_DELTA] = d;
        if (MODE == 3 && blockIdx.x == 0) cg_trace(pr, TR_U_WAITED);
      }
    }
    __syncthreads();
    delta = s_delta;
  } else {
    delta = sc[S_DELTA];
  }
  if (!(delta > 0.0)) {
    // (p, A p) <= 0 or NaN: DIVERGED_INDEFINITE_MAT, no update (PETSc: its = i+1 at that point)
    if (blockIdx.x == 0 && threadIdx.x == 0) {
      // every block takes this branch from the same delta; only one writes
      fl[F_ITS] = fl[F_ITS] + 1;
      __threadfence();
      fl[F_REASON] = isnan(delta) ? IIFE_KSP_DIVERGED_NANORINF : IIFE_KSP_DIVERGED_INDEFINITE_MAT;
    }
    return;
  }
  double beta_now = sc[S_BETA];
  if (MODE == 3) {
    const unsigned long long it_rel = *pr.iter + (unsigned long long)pr.k_off - pr.iter[1];
    beta_now = sc[S_BETA2 + (int)(it_rel & 1ull)];
  }
  double alpha = beta_now / delta;
  double acc[2] = {0.0, 0.0};
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i0 < n; i0 += VEC_ILP * stride) {
    double pv[VEC_ILP], wv[VEC_ILP], xv[VEC_ILP], rv[VEC_ILP], dv[VEC_ILP];
#pragma unroll
    for (int u = 0; u < VEC_ILP; ++u) {  // all loads of the round first (see k_cg_p)
      const int64_t i = i0 + u * stride;
      const bool in = i < n;
      pv[u] = in ? p[i] : 0.0;
      wv[u] = in ? __ldcs(w + i) : 0.0;
      xv[u] = in ? x[i] : 0.0;
      rv[u] = in ? r[i] : 0.0;
      dv[u] = (in && dinv) ? dinv[i] : 1.0;
    }
#pragma unroll
    for (int u = 0; u < VEC_ILP; ++u) {  // same order of the sums as one element per round
      const int64_t i = i0 + u * stride;
      if (i < n) {
        x[i] = fma(alpha, pv[u], xv[u]);
        const double ri = fma(-alpha, wv[u], rv[u]);
        r[i] = ri;
        const double z = dv[u] * ri;
        acc[0] = fma(z, ri, acc[0]);
        acc[1] = fma(z, z, acc[1]);
      }
    }
  }
  if (grid_reduce<2>(acc, partials, counter, out, red, &last)) {
    if (MODE == 2 || MODE == 3) {
      p2p_push(pr, 2ull * (*pr.iter + (unsigned long long)pr.k_off) + 2ull, out, 2, threadIdx.x);
      if (MODE == 3 && threadIdx.x == 0) cg_trace(pr, TR_U_OUT);
    } else if (threadIdx.x == 0) {
      if (MODE == 1) {
        sc[S_RAW + 0] = out[0];
        sc[S_RAW + 1] = out[1];
      } else {
        cg_update_scalars(sc, fl, out[0], out[1], hist, hist_len);
      }
    }
  }
}

// peer-memory path: wait for every rank's (z.r, z.z) partials, scalar step, advance the iteration counter
__global__ void k_cg_scalars_p2p(double *sc, int *fl, double *hist, long long hist_len, P2PRed pr,
                                 unsigned long long *iter) {
  if (fl[F_REASON] != 0) return;
  if (!(sc[S_DELTA] > 0.0)) return;  // reason was set by k_cg_update
  double v[2];
  p2p_wait_sum(pr, 2ull * (*iter) + 2ull, v, 2);  // one warp
  if (threadIdx.x != 0) return;
  cg_update_scalars(sc, fl, v[0], v[1], hist, hist_len);
  *iter = *iter + 1ull;
}

// ------------------------------------------------------------------------------------------------
// Three-kernel CG iteration of the row-partitioned solver over peer memory (IIFE_CG_FUSED3, default on):
//   k_cg_p_push      scalar step of the PREVIOUS iteration (every CTA sums the ranks' (z.r, z.z) partials from the
//                    mailbox in rank order and takes the same decision; CTA 0 records it), p = z + (beta/beta_old) p,
//                    and the thread that updates a boundary row stores it straight into the neighbours' ghost slots;
//                    the last CTA raises the halo flags
//   k_spmv_sell      waits for the neighbours' flags in its prologue (HaloWait), w = A p, partial (p, w) pushed
//   k_cg_update<3>   waits for delta, x / r update, partial (z.r, z.z) pushed
// instead of five (p update, halo kernel, SpMV, update, scalar kernel).  All sequence numbers are the device
// counters of the halo (dev_seq[0] exchanges, dev_seq[2] iterations, dev_seq[3] = iteration counter at the start
// of the solve) plus the iteration's index k inside the captured chunk; k_cg_chunk_end moves the counters once per chunk,
// so no kernel reads a counter that another CTA of the same kernel writes.
// ------------------------------------------------------------------------------------------------
// scalar step of the three-kernel iteration, by warp 0 of EVERY CTA (all reach the same verdict; CTA 0 records it):
// KSPSolve_CG after the update of iteration it_rel-1: beta_old <- beta, beta = (z, r), dp = ||z||, its, test.
// git / it_rel: global / solve-relative index of the iteration that is about to start.  Results in shared memory.
__device__ __forceinline__ void cg3_scalar_step(const P2PRed &pr, unsigned long long git, unsigned long long it_rel, double *sc, int *fl,
                                                double *hist, long long hist_len, double *s_bb, int *s_stop) {
  if (threadIdx.x < 32) {
    double v[2];
    p2p_wait_sum(pr, 2ull * (git - 1ull) + 2ull, v, 2);
    if (threadIdx.x == 0) {
      const double beta_old = sc[S_BETA2 + (int)((it_rel - 1ull) & 1ull)];
      const double beta = v[0], dp = sqrt(v[1]);
      const int its = (int)it_rel;
      int reason = 0, its_out = its;
      if (isnan(dp) || isinf(dp)) reason = IIFE_KSP_DIVERGED_NANORINF;
      else if (dp <= sc[S_TTOL]) reason = (dp < sc[S_ATOL]) ? IIFE_KSP_CONVERGED_ATOL : IIFE_KSP_CONVERGED_RTOL;
      else if (dp >= sc[S_DTOL] * sc[S_RHO0]) reason = IIFE_KSP_DIVERGED_DTOL;
      else if (its >= fl[F_MAXIT]) reason = IIFE_KSP_DIVERGED_ITS;
      else if (beta == 0.0) { reason = IIFE_KSP_CONVERGED_ATOL; its_out = its + 1; }
      else if (beta < 0.0) { reason = IIFE_KSP_DIVERGED_INDEFINITE_PC; its_out = its + 1; }
      else if (isnan(beta) || isinf(beta)) { reason = IIFE_KSP_DIVERGED_NANORINF; its_out = its + 1; }
      *s_bb = beta / beta_old;
      *s_stop = reason;
      if (blockIdx.x == 0) {
        sc[S_BETA2 + (int)(it_rel & 1ull)] = beta;
        sc[S_BETA_OLD] = beta_old;
        sc[S_BETA] = beta;
        sc[S_DP] = dp;
        log_hist(hist, hist_len, its, dp);
        fl[F_ITS] = its_out;
        if (reason) {
          __threadfence();
          fl[F_REASON] = reason;
        }
      }
    }
  }
}

// p = z + bb p (p = z when `first`), z = D^-1 r recomputed; the thread that updates a boundary row stores it into the
// neighbours' ghost slots; the last CTA to finish raises the halo flags with sequence number halo_seq.
__device__ __forceinline__ void cg3_p_and_push(const double *r, const double *__restrict__ dinv, double *p, int64_t n, double bb,
                                               bool first, const RowPush &rp, unsigned long long halo_seq, bool *last) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  bool stored = false;
  for (int64_t i0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i0 < n; i0 += VEC_ILP * stride) {
    double rv[VEC_ILP], dv[VEC_ILP], pv[VEC_ILP];
    uint2 mv[VEC_ILP];  // (boundary-row bits of the 32-row word, boundary rows before it)
#pragma unroll
    for (int u = 0; u < VEC_ILP; ++u) {  // all loads of the round first (see k_cg_p)
      const int64_t i = i0 + u * stride;
      const bool in = i < n;
      rv[u] = in ? r[i] : 0.0;
      dv[u] = (in && dinv) ? dinv[i] : 1.0;
      pv[u] = (in && !first) ? p[i] : 0.0;
      mv[u] = in ? __ldg((const uint2 *)rp.bmask + (i >> 5)) : make_uint2(0u, 0u);
    }
#pragma unroll
    for (int u = 0; u < VEC_ILP; ++u) {
      const int64_t i = i0 + u * stride;
      if (i >= n) continue;
      const double z = dv[u] * rv[u];
      const double pn = first ? z : fma(bb, pv[u], z);
      p[i] = pn;
      const unsigned bit = 1u << (i & 31);
      if (mv[u].x & bit) {  // a boundary row: its neighbours' ghost copies
        const int k = (int)mv[u].y + __popc(mv[u].x & (bit - 1u));
        for (int e = __ldg(rp.bptr + k); e < __ldg(rp.bptr + k + 1); ++e) rp.pt.xbuf[rp.bpeer[e]][rp.bdst[e]] = pn;
        stored = true;
      }
    }
  }
  if (stored) __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned int t = atomicAdd(rp.counter, 1u);
    *last = (t == gridDim.x - 1);
  }
  __syncthreads();
  if (!*last) return;
  __threadfence_system();
  const int q = threadIdx.x;
  if (q < rp.nranks && ((rp.send_mask >> q) & 1u)) st_flag(&rp.pt.mbox[q]->halo_flag[rp.me], halo_seq);
  if (threadIdx.x == 0) *rp.counter = 0u;
}

__global__ void __launch_bounds__(VEC_THREADS)
k_cg_p_push(const double *__restrict__ r, const double *__restrict__ dinv, double *__restrict__ p, int64_t n, double *sc,
            int *fl, double *hist, long long hist_len, P2PRed pr, RowPush rp) {
  if (fl[F_REASON] != 0) return;  // set by an earlier kernel (CTA 0 of THIS kernel can only reach the same verdict)
  __shared__ double s_bb;
  __shared__ int s_stop;
  __shared__ bool last;
  const unsigned long long git = *pr.iter + (unsigned long long)pr.k_off;
  const unsigned long long it_rel = git - pr.iter[1];
  const bool first = (it_rel == 0ull);
  if (blockIdx.x == 0 && threadIdx.x == 0) cg_trace(pr, TR_P_IN);
  if (!first) {
    cg3_scalar_step(pr, git, it_rel, sc, fl, hist, hist_len, &s_bb, &s_stop);
    if (blockIdx.x == 0 && threadIdx.x == 0) cg_trace(pr, TR_P_WAITED);
    __syncthreads();
    if (s_stop) return;
  }
  cg3_p_and_push(r, dinv, p, n, first ? 0.0 : s_bb, first, rp, *rp.seq_base + (unsigned long long)pr.k_off + 1ull, &last);
  if (last && threadIdx.x == 0) cg_trace(pr, TR_P_OUT);
}

// Two-kernel iteration (IIFE_CG_MERGED, default on): k_cg_update<3> of iteration k and k_cg_p_push of iteration k+1 in
// ONE kernel.  After its part of the x / r update every CTA waits for all ranks' (z.r, z.z) — the last CTA of every
// rank pushes them — takes the scalar step and updates its part of p.  Every CTA spins inside the kernel, so the grid
// must be co-resident (sized by the occupancy query in cg_solve); the spins are bounded like all the others.
__global__ void __launch_bounds__(VEC_THREADS)
k_cg_update_p(double *x, double *r, double *p, const double *__restrict__ w, const double *__restrict__ dinv, int64_t n,
              double *sc, int *fl, double *partials, unsigned int *counter, double *hist, long long hist_len, P2PRed pr,
              RowPush rp) {
  if (fl[F_REASON] != 0) return;
  __shared__ double red[32];
  __shared__ double out[2];
  __shared__ bool last;
  __shared__ double s_delta, s_bb;
  __shared__ int s_stop;
  const unsigned long long git = *pr.iter + (unsigned long long)pr.k_off;
  const unsigned long long it_rel = git - pr.iter[1];
  if (blockIdx.x == 0 && threadIdx.x == 0) cg_trace(pr, TR_U_IN);
  if (threadIdx.x < 32) {
    double d;
    p2p_wait_sum(pr, 2ull * git + 1ull, &d, 1);
    if (threadIdx.x == 0) {
      s_delta = d;
      if (blockIdx.x == 0) {
        sc[S_DELTA] = d;
        cg_trace(pr, TR_U_WAITED);
      }
    }
  }
  __syncthreads();
  const double delta = s_delta;
  if (!(delta > 0.0)) {  // (p, A p) <= 0 or NaN: DIVERGED_INDEFINITE_MAT, no update (PETSc: its = i+1 at that point)
    if (blockIdx.x == 0 && threadIdx.x == 0) {
      fl[F_ITS] = fl[F_ITS] + 1;
      __threadfence();
      fl[F_REASON] = isnan(delta) ? IIFE_KSP_DIVERGED_NANORINF : IIFE_KSP_DIVERGED_INDEFINITE_MAT;
    }
    return;
  }
  const double alpha = sc[S_BETA2 + (int)(it_rel & 1ull)] / delta;
  double acc[2] = {0.0, 0.0};
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i0 < n; i0 += VEC_ILP * stride) {
    double pv[VEC_ILP], wv[VEC_ILP], xv[VEC_ILP], rv[VEC_ILP], dv[VEC_ILP];
#pragma unroll
    for (int u = 0; u < VEC_ILP; ++u) {
      const int64_t i = i0 + u * stride;
      const bool in = i < n;
      pv[u] = in ? p[i] : 0.0;
      wv[u] = in ? __ldcs(w + i) : 0.0;
      xv[u] = in ? x[i] : 0.0;
      rv[u] = in ? r[i] : 0.0;
      dv[u] = (in && dinv) ? dinv[i] : 1.0;
    }
#pragma unroll
    for (int u = 0; u < VEC_ILP; ++u) {  // same order of the sums as k_cg_update
      const int64_t i = i0 + u * stride;
      if (i < n) {
        x[i] = fma(alpha, pv[u], xv[u]);
        const double ri = fma(-alpha, wv[u], rv[u]);
        r[i] = ri;
        const double z = dv[u] * ri;
        acc[0] = fma(z, ri, acc[0]);
        acc[1] = fma(z, z, acc[1]);
      }
    }
  }
  if (grid_reduce<2>(acc, partials, counter, out, red, &last)) {
    p2p_push(pr, 2ull * git + 2ull, out, 2, threadIdx.x);
    if (threadIdx.x == 0) cg_trace(pr, TR_U_OUT);
  }
  // ---- iteration git + 1: scalar step, p update, ghost push (the thread that wrote r[i] above reads it again here)
  P2PRed prn = pr;
  prn.k_off = pr.k_off + 1;
  if (blockIdx.x == 0 && threadIdx.x == 0) cg_trace(prn, TR_P_IN);
  cg3_scalar_step(pr, git + 1ull, it_rel + 1ull, sc, fl, hist, hist_len, &s_bb, &s_stop);
  if (blockIdx.x == 0 && threadIdx.x == 0) cg_trace(prn, TR_P_WAITED);
  __syncthreads();
  if (s_stop) return;
  cg3_p_and_push(r, dinv, p, n, s_bb, false, rp, *rp.seq_base + (unsigned long long)pr.k_off + 2ull, &last);
  if (last && threadIdx.x == 0) cg_trace(prn, TR_P_OUT);
}

// end of a captured chunk of `chunk` iterations: the counters the kernels above offset with k
__global__ void k_cg_chunk_end(unsigned long long *dev_seq, int chunk) {
  dev_seq[0] += (unsigned long long)chunk;
  dev_seq[2] += (unsigned long long)chunk;
}

// start of a solve on the three-kernel path: fresh iteration sequence (see k_bump_seq), its start, beta_0
__global__ void k_cg_fused3_begin(unsigned long long *dev_seq, double *sc) {
  dev_seq[3] = dev_seq[2];
  sc[S_BETA2] = sc[S_BETA];
}

// A solve that stopped on (p, A p) <= 0 has exchanged reduction 2*iter+1 without finishing iteration `iter`: the
// next solve on the same halo must not meet those flags again, so every solve starts on a fresh sequence
// number (all ranks run the same launches, so they stay in step).
__global__ void k_bump_seq(unsigned long long *iter) { *iter = *iter + 1ull; }

// row-partitioned solver: scalar step after the allreduce of the two raw sums
__global__ void k_cg_update_scalars(double *sc, int *fl, double *hist, long long hist_len) {
  if (fl[F_REASON] != 0) return;
  if (!(sc[S_DELTA] > 0.0)) return;  // reason was set by k_cg_update
  cg_update_scalars(sc, fl, sc[S_RAW + 0], sc[S_RAW + 1], hist, hist_len);
}

}  // namespace iife
namespace iife {

// ------------------------------------------------------------------------------------------------
// FGMRES kernels.  Device-side small arrays: H (m+1) x m column-major, cs[m], sn[m], rs[m+1], y[m].
// ------------------------------------------------------------------------------------------------
struct GmresSmall {
  double *H, *cs, *sn, *rs, *y;
  int m;  // restart
};

__device__ __forceinline__ void gm_cycle_scalars(double *sc, int *fl, GmresSmall gs, double vv, double bb2, int first_cycle,
                                                 double *hist, long long hist_len) {
  double res = sqrt(vv);
  if (first_cycle) {
    double rho0 = sqrt(bb2);
    if (rho0 == 0.0) rho0 = res;  // KSPConvergedDefault: zero right-hand side with a nonzero guess
    sc[S_RHO0] = rho0;
    sc[S_TTOL] = fmax(sc[S_RTOL] * rho0, sc[S_ATOL]);
    fl[F_ITS] = 0;
    log_hist(hist, hist_len, 0, res);
  }
  sc[S_RES] = res;
  gs.rs[0] = res;
  sc[S_SCALE] = res != 0.0 ? 1.0 / res : 0.0;
  fl[F_LOC_IT] = 0;
  fl[F_HAPEND] = 0;
  converged_default(sc, fl, fl[F_ITS], res);
  if (fl[F_REASON] == 0 && fl[F_ITS] >= fl[F_MAXIT]) fl[F_REASON] = IIFE_KSP_DIVERGED_ITS;
}

// ||v||^2 of the fresh residual in V0 -> rs[0], scale, test at cycle start (true residual)
template <bool DIST>
__global__ void __launch_bounds__(VEC_THREADS)
k_gm_cycle_start(const double *__restrict__ v0, const double *__restrict__ b, int64_t n, double *sc, int *fl,
                 GmresSmall gs, double *partials, unsigned int *counter, double *hist, long long hist_len,
                 int first_cycle) {
  if (fl[F_REASON] != 0) return;
  __shared__ double red[32];
  __shared__ double out[2];
  __shared__ bool last;
  double acc[2] = {0.0, 0.0};
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    double v = v0[i];
    acc[0] = fma(v, v, acc[0]);
    if (first_cycle) {
      double bi = b[i];
      acc[1] = fma(bi, bi, acc[1]);
    }
  }
  if (grid_reduce<2>(acc, partials, counter, out, red, &last) && threadIdx.x == 0) {
    if (DIST) {
      sc[S_RAW + 0] = out[0];
      sc[S_RAW + 1] = out[1];
    } else {
      gm_cycle_scalars(sc, fl, gs, out[0], out[1], first_cycle, hist, hist_len);
    }
  }
}

__global__ void k_gm_cycle_scalars(double *sc, int *fl, GmresSmall gs, int first_cycle, double *hist, long long hist_len) {
  if (fl[F_REASON] != 0) return;
  gm_cycle_scalars(sc, fl, gs, sc[S_RAW + 0], sc[S_RAW + 1], first_cycle, hist, hist_len);
}

// v_j *= scale (normalise in place), z_j = D^-1 v_j
__global__ void __launch_bounds__(VEC_THREADS)
k_gm_scale_pc(double *__restrict__ v, double *__restrict__ z, const double *__restrict__ dinv, int64_t n,
              const double *__restrict__ sc, const int *__restrict__ fl, int j) {
  if (fl[F_REASON] != 0 || fl[F_LOC_IT] != j) return;
  double s = sc[S_SCALE];
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    double vi = v[i] * s;
    v[i] = vi;
    z[i] = (dinv ? dinv[i] : 1.0) * vi;
  }
}

// classical Gram-Schmidt, part 1: h_k = (w, v_k), k = 0..j, into H[:, j].  KT vectors per sweep.
constexpr int KT = 8;
__global__ void __launch_bounds__(VEC_THREADS)
k_gm_dots(const double *__restrict__ w, double *const *__restrict__ V, int64_t n, int j, GmresSmall gs,
          const int *__restrict__ fl, double *mpartials, unsigned int *counter, int gate) {
  if (fl[F_REASON] != 0 || fl[F_LOC_IT] != gate) return;
  __shared__ double red[32];
  __shared__ bool last;
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  int64_t i0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  for (int k0 = 0; k0 <= j; k0 += KT) {
    double acc[KT];
    const double *vp[KT];
#pragma unroll
    for (int k = 0; k < KT; ++k) {
      acc[k] = 0.0;
      vp[k] = V[min(k0 + k, j)];
    }
    for (int64_t i = i0; i < n; i += stride) {
      double wi = w[i];
#pragma unroll
      for (int k = 0; k < KT; ++k) acc[k] = fma(wi, vp[k][i], acc[k]);
    }
#pragma unroll
    for (int k = 0; k < KT; ++k) {
      double b = block_sum(acc[k], red);
      if (threadIdx.x == 0 && k0 + k <= j) mpartials[(size_t)(k0 + k) * gridDim.x + blockIdx.x] = b;
    }
  }
  if (threadIdx.x == 0) {
    __threadfence();
    unsigned int t = atomicAdd(counter, 1u);
    last = (t == gridDim.x - 1);
  }
  __syncthreads();
  if (!last) return;
  __threadfence();
  // one warp per coefficient, fixed order
  int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  for (int k = warp; k <= j; k += nw) {
    double s = 0.0;
    for (int bb = lane; bb < (int)gridDim.x; bb += 32) s += __ldcg(mpartials + (size_t)k * gridDim.x + bb);
    s = warp_sum(s);
    if (lane == 0) gs.H[(size_t)j * (gs.m + 1) + k] = s;
  }
  if (threadIdx.x == 0) *counter = 0u;
}

// Hessenberg / Givens update, residual estimate and convergence test of one FGMRES step (one thread)
__device__ __forceinline__ void gm_update_scalars(GmresSmall gs, double *sc, int *fl, int j, double norm2, double *hist,
                                                  long long hist_len) {
    const int m1 = gs.m + 1;
    double *hh = gs.H + (size_t)j * m1;
    double tt = sqrt(norm2);
    // happy breakdown test (fgmres.c): hapbnd = min(|tt / rs[j]|, haptol = 1e-30)
    double hapbnd = fabs(tt / gs.rs[j]);
    if (hapbnd > 1e-30) hapbnd = 1e-30;
    bool hapend = false;
    if (tt > hapbnd) {
      sc[S_SCALE] = 1.0 / tt;
    } else {
      sc[S_SCALE] = 0.0;
      hapend = true;
    }
    hh[j + 1] = tt;
    // apply the previous rotations to the new column
    for (int k = 0; k < j; ++k) {
      double t1 = hh[k], t2 = hh[k + 1];
      hh[k] = gs.cs[k] * t1 + gs.sn[k] * t2;
      hh[k + 1] = -gs.sn[k] * t1 + gs.cs[k] * t2;
    }
    // new rotation annihilating hh[j+1]
    double res;
    if (!hapend) {
      double t = sqrt(hh[j] * hh[j] + hh[j + 1] * hh[j + 1]);
      if (t == 0.0) {
        fl[F_REASON] = IIFE_KSP_DIVERGED_BREAKDOWN;
        t = 1.0;
      }
      gs.cs[j] = hh[j] / t;
      gs.sn[j] = hh[j + 1] / t;
      gs.rs[j + 1] = -gs.sn[j] * gs.rs[j];
      gs.rs[j] = gs.cs[j] * gs.rs[j];
      hh[j] = gs.cs[j] * hh[j] + gs.sn[j] * hh[j + 1];
      res = fabs(gs.rs[j + 1]);
    } else {
      // happy breakdown: the Krylov space is invariant, the exact solution is in it
      res = 0.0;
      gs.rs[j + 1] = 0.0;
    }
    int its = fl[F_ITS] + 1;
    fl[F_ITS] = its;
    fl[F_LOC_IT] = j + 1;
    sc[S_RES] = res;
    log_hist(hist, hist_len, its, res);
    converged_default(sc, fl, its, res);
    if (fl[F_REASON] == 0) {
      if (hapend) fl[F_REASON] = IIFE_KSP_DIVERGED_BREAKDOWN;
      else if (its >= fl[F_MAXIT]) fl[F_REASON] = IIFE_KSP_DIVERGED_ITS;
    }
}

// classical Gram-Schmidt, part 2: w -= sum_k h_k v_k; ||w||^2; then (last CTA, one thread) the
// Hessenberg/Givens update, residual estimate and convergence test of PETSc's KSPFGMRESCycle.
constexpr int ROWS_PT = 2;
template <bool DIST>
__global__ void __launch_bounds__(VEC_THREADS)
k_gm_update(double *__restrict__ w, double *const *__restrict__ V, int64_t n, int j, GmresSmall gs, double *sc,
            int *fl, double *partials, unsigned int *counter, double *hist, long long hist_len) {
  if (fl[F_REASON] != 0 || fl[F_LOC_IT] != j) return;
  extern __shared__ double hsh[];  // j+1 coefficients
  __shared__ double red[32];
  __shared__ double out[1];
  __shared__ bool last;
  const double *Hj = gs.H + (size_t)j * (gs.m + 1);
  for (int k = threadIdx.x; k <= j; k += blockDim.x) hsh[k] = Hj[k];
  __syncthreads();
  double acc[1] = {0.0};
  int64_t stride = (int64_t)gridDim.x * blockDim.x * ROWS_PT;
  for (int64_t base = ((int64_t)blockIdx.x * blockDim.x) * ROWS_PT + threadIdx.x; base < n; base += stride) {
    double a[ROWS_PT];
    int64_t idx[ROWS_PT];
#pragma unroll
    for (int r = 0; r < ROWS_PT; ++r) {
      idx[r] = base + (int64_t)r * blockDim.x;
      a[r] = idx[r] < n ? w[idx[r]] : 0.0;
    }
    for (int k = 0; k <= j; ++k) {
      const double *vk = V[k];
      double h = hsh[k];
#pragma unroll
      for (int r = 0; r < ROWS_PT; ++r)
        if (idx[r] < n) a[r] = fma(-h, vk[idx[r]], a[r]);
    }
#pragma unroll
    for (int r = 0; r < ROWS_PT; ++r)
      if (idx[r] < n) {
        w[idx[r]] = a[r];
        acc[0] = fma(a[r], a[r], acc[0]);
      }
  }
  if (grid_reduce<1>(acc, partials, counter, out, red, &last) && threadIdx.x == 0) {
    if (DIST) sc[S_RAW + 0] = out[0];
    else gm_update_scalars(gs, sc, fl, j, out[0], hist, hist_len);
  }
}

__global__ void k_gm_update_scalars(GmresSmall gs, double *sc, int *fl, int j, double *hist, long long hist_len) {
  if (fl[F_REASON] != 0 || fl[F_LOC_IT] != j) return;
  gm_update_scalars(gs, sc, fl, j, sc[S_RAW + 0], hist, hist_len);
}

// back substitution H(0:k,0:k) y = rs(0:k), k = loc_it, one warp
__global__ void k_gm_solve_y(GmresSmall gs, const int *__restrict__ fl) {
  int k = fl[F_LOC_IT];
  if (k <= 0) return;
  int lane = threadIdx.x;
  const int m1 = gs.m + 1;
  for (int i = k - 1; i >= 0; --i) {
    double s = 0.0;
    for (int c = i + 1 + lane; c < k; c += 32) s += gs.H[(size_t)c * m1 + i] * gs.y[c];
    s = warp_sum(s);
    s = __shfl_sync(0xffffffffu, s, 0);
    if (lane == 0) {
      double d = gs.H[(size_t)i * m1 + i];
      gs.y[i] = d != 0.0 ? (gs.rs[i] - s) / d : 0.0;
    }
    __syncwarp();
  }
}

// x += sum_{k < loc_it} y_k z_k ; the last CTA resets loc_it
__global__ void __launch_bounds__(VEC_THREADS)
k_gm_build_x(double *__restrict__ x, double *const *__restrict__ Z, int64_t n, GmresSmall gs, int *fl,
             unsigned int *counter) {
  int kk = fl[F_LOC_IT];
  if (kk <= 0) return;
  extern __shared__ double ysh[];
  __shared__ bool last;
  for (int k = threadIdx.x; k < kk; k += blockDim.x) ysh[k] = gs.y[k];
  __syncthreads();
  int64_t stride = (int64_t)gridDim.x * blockDim.x * ROWS_PT;
  for (int64_t base = ((int64_t)blockIdx.x * blockDim.x) * ROWS_PT + threadIdx.x; base < n; base += stride) {
    double a[ROWS_PT];
    int64_t idx[ROWS_PT];
#pragma unroll
    for (int r = 0; r < ROWS_PT; ++r) {
      idx[r] = base + (int64_t)r * blockDim.x;
      a[r] = idx[r] < n ? x[idx[r]] : 0.0;
    }
    for (int k = 0; k < kk; ++k) {
      const double *zk = Z[k];
      double y = ysh[k];
#pragma unroll
      for (int r = 0; r < ROWS_PT; ++r)
        if (idx[r] < n) a[r] = fma(y, zk[idx[r]], a[r]);
    }
#pragma unroll
    for (int r = 0; r < ROWS_PT; ++r)
      if (idx[r] < n) x[idx[r]] = a[r];
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    unsigned int t = atomicAdd(counter, 1u);
    last = (t == gridDim.x - 1);
    if (last) {
      fl[F_LOC_IT] = 0;
      *counter = 0u;
    }
  }
}

// r = b  (copy; the SpMV r -= A x follows), gated by the reason flag
__global__ void __launch_bounds__(VEC_THREADS)
k_copy_gated(const double *__restrict__ src, double *__restrict__ dst, int64_t n, const int *__restrict__ fl) {
  if (fl && fl[F_REASON] != 0) return;
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) dst[i] = src[i];
}

// y = y - A x gated: implemented with the plain SpMV when the flag is clear (checked on the device
// inside a tiny wrapper kernel would cost a launch; instead FGMRES restarts are rare and the SpMV
// result is harmless after convergence because V0 is not read any more).

static int vec_grid(int64_t n) {
  int64_t g = (n + VEC_THREADS - 1) / VEC_THREADS;
  int64_t cap = (int64_t)ctx().sm_count * 8;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  if (g > MAX_PARTIALS) g = MAX_PARTIALS;
  return (int)g;
}

struct HostFlags {
  int fl[F_COUNT];
};

static int poll_flags(const KspWork &w, HostFlags *pinned) {
  IIFE_CUDA(cudaMemcpyAsync(pinned->fl, w.fl, sizeof(int) * F_COUNT, cudaMemcpyDeviceToHost, ctx().stream));
  IIFE_CUDA(cudaStreamSynchronize(ctx().stream));
  return IIFE_OK;
}

static double now_ms() {
  return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
}
static bool ksp_debug() {
  static int v = -1;
  if (v < 0) v = getenv("IIFE_KSP_DEBUG") ? 1 : 0;
  return v == 1;
}
#define KSP_DBG(tag)                                                          \
  do {                                                                        \
    if (ksp_debug()) {                                                        \
      double _t = now_ms();                                                   \
      fprintf(stderr, "[ksp] %-22s +%.3f ms\n", tag, _t - dbg_t0);            \
      dbg_t0 = _t;                                                            \
    }                                                                         \
  } while (0)

static int env_int(const char *name, int dflt) {
  const char *s = getenv(name);
  return s ? atoi(s) : dflt;
}

// ------------------------------------------------------------------------------------------------
// CG driver (device pointers)
// ------------------------------------------------------------------------------------------------
// Work arrays of a solve come from an arena that hands the k-th request of a solve the block the k-th request of the
// previous solve got (same size): the addresses in the kernels' arguments stay the same from solve to solve, which is
// what lets the captured CG chunk below be reused.  The blocks (a few vectors of the last system size) stay with the
// library until iife_finalize.
struct KspArena {
  std::vector<std::pair<void *, size_t>> blocks;
  size_t cursor = 0;
};
static KspArena g_arena;
static int arena_alloc(void **p, size_t bytes) {
  if (bytes < 8) bytes = 8;
  KspArena &a = g_arena;
  if (a.cursor < a.blocks.size() && a.blocks[a.cursor].second == bytes) {
    *p = a.blocks[a.cursor++].first;
    return IIFE_OK;
  }
  for (size_t k = a.cursor; k < a.blocks.size(); ++k) dev_free(a.blocks[k].first, a.blocks[k].second);
  a.blocks.resize(a.cursor);
  IIFE_TRY(dev_alloc(p, bytes));
  a.blocks.emplace_back(*p, bytes);
  a.cursor++;
  return IIFE_OK;
}
static void arena_release() {
  for (auto &b : g_arena.blocks) dev_free(b.first, b.second);
  g_arena.blocks.clear();
  g_arena.cursor = 0;
}
template <class T>
struct Ws {  // like Tmp<T>, owned by the arena
  T *p = nullptr;
  int alloc(size_t count) { return arena_alloc((void **)&p, (count ? count : 1) * sizeof(T)); }
};

// The captured chunk of CG iterations is kept between solves: a graph depends on nothing but its kernels' arguments
// and launch shapes, so it is reused while every value that enters them is unchanged (the caching allocator hands
// the work vectors of the next solve the same blocks; a new A_b of the same plan gets the SELL arrays of the old
// one).  Any difference in the key -> capture and instantiate again (0.2-0.4 ms).  IIFE_KSP_GRAPH_CACHE=0 disables.
struct CgGraphKey {
  const void *ptr[20];
  long long n, slices, hist_len;
  int ints[12];
  P2PRed pr;
  RowPush rp;
  HaloWait hw;
};
struct CgGraphEntry {
  bool valid = false;
  CgGraphKey key;
  cudaGraph_t graph = nullptr;
  cudaGraphExec_t exec = nullptr;
  int64_t launches = 0;
};
// a few entries: consecutive steps alternate between two sets of addresses (the previous A_b is still alive while the
// next one is built)
constexpr int CG_GRAPH_CACHE = 4;
static CgGraphEntry g_cg_graph[CG_GRAPH_CACHE];
static int g_cg_graph_next = 0;

static void cg_graph_entry_free(CgGraphEntry &e) {
  if (e.exec) cudaGraphExecDestroy(e.exec);
  if (e.graph) cudaGraphDestroy(e.graph);
  e.exec = nullptr;
  e.graph = nullptr;
  e.valid = false;
}

void ksp_release_cached_graphs() {
  for (auto &e : g_cg_graph) cg_graph_entry_free(e);
  arena_release();
}

static int cg_solve(Mat *A, Halo *H, const double *dinv, const double *b, double *x, int64_t max_it, KspWork &w,
                    HostFlags *hf) {
  Ctx &c = ctx();
  const int64_t n = A->n_rows;
  const bool dist = (H != nullptr) && c.nranks > 1;
  const int64_t n_ext = H ? H->n_owned + H->n_ghost : n;  // owned + ghost entries of a multiplied vector
  double dbg_t0 = now_ms();
  // peer-memory path (p2p.cu): p lives in the halo's IPC-exported vector so that neighbours can store
  // their boundary entries straight into its ghost section; exchanges are library kernels, not NCCL
  const bool p2p = dist && H->p2p;
  // timing experiments only (results are wrong with these set): drop the halo / the reductions
  static const bool dbg_nohalo = getenv("IIFE_DBG_NOHALO") != nullptr, dbg_nored = getenv("IIFE_DBG_NORED") != nullptr;
  struct PBuf { double *p = nullptr; } p;
  Ws<double> r, p_own, wv;
  Ws<unsigned long long> trace;
  IIFE_TRY(r.alloc((size_t)n));
  if (p2p) p.p = H->xbuf;
  else {
    IIFE_TRY(p_own.alloc((size_t)n_ext));
    p.p = p_own.p;
  }
  IIFE_TRY(wv.alloc((size_t)n));
  const int g = vec_grid(n);
  auto xchg = [&](const int *flag) -> int { return p2p ? p2p_halo_exchange(H, flag) : halo_exchange(H, p.p); };
  auto ar = [&](double *v, int cnt, const int *flag) -> int { return p2p ? p2p_allreduce(H, v, cnt, flag) : allreduce_sum(v, cnt); };
  P2PRed pr{};
  if (p2p) {
    pr.enabled = 1;
    pr.me = H->me;
    pr.nranks = H->nranks;
    pr.mbox = H->mbox;
    for (int q = 0; q < P2P_MAX_RANKS; ++q) pr.peer[q] = H->peer_mbox[q];
    pr.iter = H->dev_seq + 2;
    pr.err = H->p2p_err;
    pr.ll = env_int("IIFE_P2P_LL", 1) != 0 ? 1 : 0;  // packed 8-byte words instead of values + fence + flag
    if (env_int("IIFE_CG_TRACE", 0) != 0) {
      IIFE_TRY(trace.alloc((size_t)CG_TRACE_ITERS * CG_TRACE_SLOTS));
      IIFE_CUDA(cudaMemsetAsync(trace.p, 0, sizeof(unsigned long long) * CG_TRACE_ITERS * CG_TRACE_SLOTS, c.stream));
      pr.trace = trace.p;
    }
    IIFE_LAUNCH(k_bump_seq, 1, 1, 0, H->dev_seq + 2);
  }
  // r = b - A x0   (row-partitioned: x0 is staged in p to receive its ghost entries)
  IIFE_LAUNCH(k_copy_gated, g, VEC_THREADS, 0, b, r.p, n, (const int *)nullptr);
  if (H) {
    IIFE_LAUNCH(k_copy_gated, g, VEC_THREADS, 0, (const double *)x, p.p, n, (const int *)nullptr);
    if (dist) IIFE_TRY(xchg(nullptr));
    IIFE_TRY(spmv_launch(A, -1.0, p.p, 1.0, r.p));
  } else {
    IIFE_TRY(spmv_launch(A, -1.0, x, 1.0, r.p));
  }
  if (dist) {
    IIFE_LAUNCH(k_cg_init<true>, g, VEC_THREADS, 0, r.p, b, dinv, n, w.sc, w.fl, w.partials, w.counters, w.hist,
                (long long)w.hist_len);
    IIFE_TRY(ar(w.sc + S_RAW, 3, nullptr));
    IIFE_LAUNCH(k_cg_init_scalars, 1, 1, 0, w.sc, w.fl, w.hist, (long long)w.hist_len);
  } else {
    IIFE_LAUNCH(k_cg_init<false>, g, VEC_THREADS, 0, r.p, b, dinv, n, w.sc, w.fl, w.partials, w.counters, w.hist,
                (long long)w.hist_len);
  }
  IIFE_CHECK_LAUNCH();
  KSP_DBG("alloc+init enqueue");
  // The flag readback that tells whether iteration 0 already converged waits for everything enqueued so far (SELL fill,
  // initial residual).  The host-side preparation of the iteration (halo arguments, graph capture + instantiation:
  // 0.3-0.5 ms) does not depend on it, so it is done first, while the device is busy; the poll follows it.
  const bool poll_early = getenv("IIFE_KSP_TIMELINE") != nullptr;  // the timeline aid runs iterations before the capture
  if (poll_early) {
    IIFE_TRY(poll_flags(w, hf));
    KSP_DBG("init poll");
    if (hf->fl[F_REASON] != 0) return IIFE_OK;
  }

  int chunk = env_int("IIFE_KSP_CHUNK", 32);
  if (chunk < 1) chunk = 1;
  // three-kernel iteration of the peer-memory path (k_cg_p_push / SpMV with halo wait / k_cg_update<3>)
  const bool fused3 = p2p && mat_sell_ready(A) && H->bmask && env_int("IIFE_CG_FUSED3", 1) != 0 && !dbg_nohalo &&
                      !dbg_nored;
  RowPush rpush{};
  HaloWait hwait{};
  if (fused3) {
    IIFE_LAUNCH(k_cg_fused3_begin, 1, 1, 0, H->dev_seq, w.sc);
    rpush.n_brow = H->n_brow;
    rpush.brow = H->brow;
    rpush.bptr = H->bptr;
    rpush.bpeer = H->bpeer;
    rpush.bdst = H->bdst;
    rpush.bmask = H->bmask;
    for (int q = 0; q < P2P_MAX_RANKS; ++q) {
      rpush.pt.xbuf[q] = H->peer_xbuf[q];
      rpush.pt.mbox[q] = H->peer_mbox[q];
      rpush.pt.dst_start[q] = H->dst_start[q];
    }
    rpush.me = H->me;
    rpush.nranks = H->nranks;
    rpush.send_mask = H->send_mask;
    rpush.seq_base = H->dev_seq;
    rpush.counter = H->p2p_counter;
    hwait.flags = H->mbox->halo_flag;
    hwait.seq_base = H->dev_seq;
    hwait.nranks = H->nranks;
    hwait.recv_mask = H->recv_mask;
    hwait.err = H->p2p_err;
    // interior slices before the ghost wait when they form one run (IIFE_CG_INTERIOR_FIRST=0: wait in the prologue)
    if (env_int("IIFE_CG_INTERIOR_FIRST", 1) != 0) {
      IIFE_TRY(mat_ensure_sell_order(A, H->n_owned));
      if (A->sell_int_lo >= 0) {
        hwait.interior_first = 1;
        hwait.int_lo = A->sell_int_lo;
        hwait.n_int = A->sell_n_interior;
      }
    }
  }
  // NCCL calls inside the loop: keep to plain stream launches (no graph capture) in that case
  const bool use_graph = env_int("IIFE_KSP_GRAPH", 1) != 0 && (!dist || p2p);
  // two-kernel iteration: update of iteration k and p update of iteration k + 1 in one co-resident kernel
  bool merged = fused3 && env_int("IIFE_CG_MERGED", 1) != 0;
  int g_merged = g;
  if (merged) {
    int per_sm = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_cg_update_p, VEC_THREADS, 0) != cudaSuccess || per_sm < 1) {
      cudaGetLastError();
      merged = false;
    } else {
      g_merged = std::min(g, per_sm * c.sm_count);
    }
  }
  auto enqueue_iteration = [&](int k) -> int {
    if (fused3 && merged) {
      P2PRed prk = pr;
      prk.k_off = k;
      HaloWait hwk = hwait;
      hwk.k_off = k;
      IIFE_TRY(spmv_dot_launch(A, p.p, wv.p, w.sc + S_DELTA, w.partials + 2 * MAX_PARTIALS, w.counters + 1, w.fl, &prk, &hwk));
      IIFE_LAUNCH(k_cg_update_p, g_merged, VEC_THREADS, 0, x, r.p, p.p, (const double *)wv.p, dinv, n, w.sc, w.fl, w.partials, w.counters,
                  w.hist, (long long)w.hist_len, prk, rpush);
      return IIFE_OK;
    }
    if (fused3) {
      P2PRed prk = pr;
      prk.k_off = k;
      HaloWait hwk = hwait;
      hwk.k_off = k;
      IIFE_LAUNCH(k_cg_p_push, g, VEC_THREADS, 0, (const double *)r.p, dinv, p.p, n, w.sc, w.fl, w.hist, (long long)w.hist_len, prk, rpush);
      IIFE_TRY(spmv_dot_launch(A, p.p, wv.p, w.sc + S_DELTA, w.partials + 2 * MAX_PARTIALS, w.counters + 1, w.fl, &prk, &hwk));
      IIFE_LAUNCH(k_cg_update<3>, g, VEC_THREADS, 0, x, r.p, (const double *)p.p, (const double *)wv.p, dinv, n, w.sc, w.fl,
                      w.partials, w.counters, w.hist, (long long)w.hist_len, prk);
      return IIFE_OK;
    }
    IIFE_LAUNCH(k_cg_p, g, VEC_THREADS, 0, (const double *)r.p, dinv, p.p, n, (const double *)w.sc, (const int *)w.fl);
    if (dist && !dbg_nohalo) IIFE_TRY(xchg(w.fl));
    IIFE_TRY(spmv_dot_launch(A, p.p, wv.p, w.sc + S_DELTA, w.partials + 2 * MAX_PARTIALS, w.counters + 1, w.fl,
                             (p2p && !dbg_nored) ? &pr : nullptr));
    if (dbg_nored) {
      IIFE_LAUNCH(k_cg_update<0>, g, VEC_THREADS, 0, x, r.p, p.p, wv.p, dinv, n, w.sc, w.fl, w.partials, w.counters,
                  w.hist, (long long)w.hist_len, pr);
    } else if (p2p) {
      // reductions ride inside the compute kernels: partials pushed by the producer's last CTA, summed
      // in rank order by the consumer
      IIFE_LAUNCH(k_cg_update<2>, g, VEC_THREADS, 0, x, r.p, p.p, wv.p, dinv, n, w.sc, w.fl, w.partials, w.counters,
                  w.hist, (long long)w.hist_len, pr);
      IIFE_LAUNCH(k_cg_scalars_p2p, 1, 32, 0, w.sc, w.fl, w.hist, (long long)w.hist_len, pr, H->dev_seq + 2);
    } else if (dist) {
      IIFE_TRY(ar(w.sc + S_DELTA, 1, w.fl));
      IIFE_LAUNCH(k_cg_update<1>, g, VEC_THREADS, 0, x, r.p, p.p, wv.p, dinv, n, w.sc, w.fl, w.partials, w.counters,
                  w.hist, (long long)w.hist_len, pr);
      IIFE_TRY(ar(w.sc + S_RAW, 2, w.fl));
      IIFE_LAUNCH(k_cg_update_scalars, 1, 1, 0, w.sc, w.fl, w.hist, (long long)w.hist_len);
    } else {
      IIFE_LAUNCH(k_cg_update<0>, g, VEC_THREADS, 0, x, r.p, (const double *)p.p, (const double *)wv.p, dinv, n, w.sc, w.fl,
                      w.partials, w.counters, w.hist, (long long)w.hist_len, pr);
    }
    return IIFE_OK;
  };
  if (getenv("IIFE_KSP_TIMELINE") && !dist) {
    // development aid: device timeline of the first iterations (events between the kernels)
    const int NIT = 12;
    const int tmask = atoi(getenv("IIFE_KSP_TIMELINE"));  // bit 0: p update, bit 1: SpMV+dot, bit 2: x/r update
    cudaEvent_t ev[NIT * 3 + 1];
    for (auto &e : ev) cudaEventCreate(&e);
    cudaEventRecord(ev[0], c.stream);
    for (int k = 0; k < NIT; ++k) {
      if (tmask & 1) IIFE_LAUNCH(k_cg_p, g, VEC_THREADS, 0, r.p, dinv, p.p, n, w.sc, w.fl);
      cudaEventRecord(ev[3 * k + 1], c.stream);
      if (tmask & 2) IIFE_TRY(spmv_dot_launch(A, p.p, wv.p, w.sc + S_DELTA, w.partials + 2 * MAX_PARTIALS, w.counters + 1, w.fl, nullptr));
      if (tmask & 8) IIFE_TRY(spmv_launch(A, 1.0, p.p, 0.0, wv.p));
      cudaEventRecord(ev[3 * k + 2], c.stream);
      if (tmask & 4) IIFE_LAUNCH(k_cg_update<0>, g, VEC_THREADS, 0, x, r.p, p.p, wv.p, dinv, n, w.sc, w.fl, w.partials, w.counters,
                  w.hist, (long long)w.hist_len, pr);
      cudaEventRecord(ev[3 * k + 3], c.stream);
    }
    cudaStreamSynchronize(c.stream);
    for (int k = NIT - 4; k < NIT; ++k) {
      float a = 0, b = 0, d = 0;
      cudaEventElapsedTime(&a, ev[3 * k], ev[3 * k + 1]);
      cudaEventElapsedTime(&b, ev[3 * k + 1], ev[3 * k + 2]);
      cudaEventElapsedTime(&d, ev[3 * k + 2], ev[3 * k + 3]);
      fprintf(stderr, "[timeline] it %d: k_cg_p %.1f us, spmv_dot %.1f us, k_cg_update %.1f us\n", k, a * 1e3, b * 1e3, d * 1e3);
    }
    for (auto &e : ev) cudaEventDestroy(e);
  }
  cudaGraph_t graph = nullptr;
  cudaGraphExec_t exec = nullptr;
  int64_t launches_per_chunk = 0;
  bool exec_cached = false, graph_hit = false;
  if (use_graph) {
    CgGraphKey key;
    memset(&key, 0, sizeof key);
    const void *ptrs[20] = {A->rowptr, A->colind, A->val, A->sell_ptr, A->sell_cptr, A->sell_col, A->sell_val, dinv, b, x,
                            r.p, p.p, wv.p, w.sc, w.fl, w.partials, w.counters, w.hist, H ? (const void *)H->dev_seq : nullptr, nullptr};
    for (int k = 0; k < 20; ++k) key.ptr[k] = ptrs[k];
    key.n = n;
    key.slices = A->sell_slices;
    key.hist_len = (long long)w.hist_len;
    const int ints[12] = {chunk, g, g_merged, fused3 ? 1 : 0, merged ? 1 : 0, dist ? 1 : 0, p2p ? 1 : 0, A->sell_state,
                          mat_sell_ready(A) ? 1 : 0, spmv_pick_lpr(A), (dbg_nohalo ? 1 : 0) | (dbg_nored ? 2 : 0), spmv_launch_signature()};
    for (int k = 0; k < 12; ++k) key.ints[k] = ints[k];
    memcpy(&key.pr, &pr, sizeof pr);
    memcpy(&key.rp, &rpush, sizeof rpush);
    memcpy(&key.hw, &hwait, sizeof hwait);
    const bool cache_on = env_int("IIFE_KSP_GRAPH_CACHE", 1) != 0;
    const CgGraphEntry *hit = nullptr;
    if (cache_on)
      for (const auto &e : g_cg_graph)
        if (e.valid && memcmp(&e.key, &key, sizeof key) == 0) hit = &e;
    if (!hit && ksp_debug()) {  // which part of the key moved since the newest entry
      const CgGraphEntry &le = g_cg_graph[(g_cg_graph_next + CG_GRAPH_CACHE - 1) % CG_GRAPH_CACHE];
      if (le.valid) {
        char line[256];
        int len = snprintf(line, sizeof line, "[ksp] graph key differs from the newest entry in:");
        for (int k = 0; k < 20; ++k)
          if (le.key.ptr[k] != key.ptr[k]) len += snprintf(line + len, sizeof line - len, " ptr%d", k);
        for (int k = 0; k < 12; ++k)
          if (le.key.ints[k] != key.ints[k]) len += snprintf(line + len, sizeof line - len, " int%d", k);
        if (memcmp(&le.key.pr, &key.pr, sizeof key.pr)) len += snprintf(line + len, sizeof line - len, " pr");
        if (memcmp(&le.key.rp, &key.rp, sizeof key.rp)) len += snprintf(line + len, sizeof line - len, " rp");
        if (memcmp(&le.key.hw, &key.hw, sizeof key.hw)) len += snprintf(line + len, sizeof line - len, " hw");
        fprintf(stderr, "%s\n", line);
      }
    }
    if (hit) {
      graph_hit = true;
      exec = hit->exec;
      launches_per_chunk = hit->launches;
      exec_cached = true;
    } else {
      int64_t before = c.launches;
      cudaError_t e = cudaStreamBeginCapture(c.stream, cudaStreamCaptureModeThreadLocal);
      if (e == cudaSuccess) {
        int rc = IIFE_OK;
        for (int k = 0; k < chunk && rc == IIFE_OK; ++k) rc = enqueue_iteration(k);
        if (fused3 && rc == IIFE_OK) IIFE_LAUNCH(k_cg_chunk_end, 1, 1, 0, H->dev_seq, chunk);
        e = cudaStreamEndCapture(c.stream, &graph);
        if (rc == IIFE_OK && e == cudaSuccess && graph) e = cudaGraphInstantiate(&exec, graph, 0);
        if (rc != IIFE_OK || e != cudaSuccess || !exec) {
          cudaGetLastError();
          if (graph) cudaGraphDestroy(graph);
          graph = nullptr;
          exec = nullptr;
        }
      } else {
        cudaGetLastError();
      }
      launches_per_chunk = c.launches - before;
      c.launches = before;  // captured, not launched yet
      if (exec && cache_on) {  // the cache owns it from here (round-robin replacement)
        CgGraphEntry &e2 = g_cg_graph[g_cg_graph_next];
        g_cg_graph_next = (g_cg_graph_next + 1) % CG_GRAPH_CACHE;
        cg_graph_entry_free(e2);
        e2.key = key;
        e2.graph = graph;
        e2.exec = exec;
        e2.launches = launches_per_chunk;
        e2.valid = true;
        exec_cached = true;
      }
    }
  }
  KSP_DBG(graph_hit ? "graph: cached chunk" : "graph: capture+inst");
  if (!poll_early) {
    int prc = poll_flags(w, hf);
    KSP_DBG("init poll");
    if (prc != IIFE_OK || hf->fl[F_REASON] != 0) {  // converged (or failed) at iteration 0: nothing to run
      if (!exec_cached) {
        if (exec) cudaGraphExecDestroy(exec);
        if (graph) cudaGraphDestroy(graph);
      }
      return prc;
    }
  }
  // Chunks are enqueued two deep: the flag readback of chunk k is awaited while chunk k+1 already
  // runs, so host scheduling jitter between chunks never idles the GPU (after convergence the chunk
  // in flight is a row of no-op kernels).
  int rc = IIFE_OK;
  int64_t enq = 0;
  cudaEvent_t ev[2] = {nullptr, nullptr};
  cudaEventCreateWithFlags(&ev[0], cudaEventDisableTiming);
  cudaEventCreateWithFlags(&ev[1], cudaEventDisableTiming);
  auto enqueue_chunk = [&](int slot) -> int {
    if (exec) {
      cudaError_t e = cudaGraphLaunch(exec, c.stream);
      if (e != cudaSuccess) return set_err(IIFE_ERR_CUDA, "cudaGraphLaunch: %s", cudaGetErrorString(e));
      c.launches += launches_per_chunk;
    } else {
      for (int k = 0; k < chunk; ++k) IIFE_TRY(enqueue_iteration(k));
      if (fused3) IIFE_LAUNCH(k_cg_chunk_end, 1, 1, 0, H->dev_seq, chunk);
      IIFE_CUDA(cudaGetLastError());
    }
    IIFE_CUDA(cudaMemcpyAsync(hf[slot].fl, w.fl, sizeof(int) * F_COUNT, cudaMemcpyDeviceToHost, c.stream));
    IIFE_CUDA(cudaEventRecord(ev[slot], c.stream));
    enq += chunk;
    return IIFE_OK;
  };
  if (merged) {  // p of the first iteration (the merged kernel produces the later ones)
    P2PRed pr0 = pr;
    pr0.k_off = 0;
    IIFE_LAUNCH(k_cg_p_push, g, VEC_THREADS, 0, (const double *)r.p, dinv, p.p, n, w.sc, w.fl, w.hist, (long long)w.hist_len, pr0, rpush);
  }
  rc = enqueue_chunk(0);
  if (rc == IIFE_OK) rc = enqueue_chunk(1);
  for (int k = 0; rc == IIFE_OK; ++k) {
    int slot = k & 1;
    cudaError_t e = cudaEventSynchronize(ev[slot]);
    if (e != cudaSuccess) { rc = set_err(IIFE_ERR_CUDA, "CG chunk: %s", cudaGetErrorString(e)); break; }
    KSP_DBG("chunk done");
    if (hf[slot].fl[F_REASON] != 0) break;
    if (enq > max_it + 2 * (int64_t)chunk) { rc = set_err(IIFE_ERR_STATE, "CG driver ran past max_it without a reason"); break; }
    rc = enqueue_chunk(slot);
  }
  cudaStreamSynchronize(c.stream);
  if (trace.p && rc == IIFE_OK) {
    // averages over iterations 20 .. its-5 of this solve (nanoseconds between the stamps; see p2p_dev.cuh)
    std::vector<unsigned long long> t((size_t)CG_TRACE_ITERS * CG_TRACE_SLOTS);
    cudaMemcpy(t.data(), trace.p, t.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost);
    const int its = hf[0].fl[F_ITS] > hf[1].fl[F_ITS] ? hf[0].fl[F_ITS] : hf[1].fl[F_ITS];
    const int lo = 20, hi = std::min(its - 5, CG_TRACE_ITERS - 2);
    if (hi > lo) {
      auto at = [&](int it, int slot) { return (double)t[(size_t)it * CG_TRACE_SLOTS + slot]; };
      const char *names[9] = {"p: wait (z,r)", "p: update + push", "gap p -> spmv", "spmv: to ghost wait done", "spmv: rest + dot push",
                              "gap spmv -> update", "update: wait delta", "update: x, r + push", "gap update -> p"};
      double sum[9] = {0};
      for (int it = lo; it < hi; ++it) {
        sum[0] += at(it, TR_P_WAITED) - at(it, TR_P_IN);
        sum[1] += at(it, TR_P_OUT) - at(it, TR_P_WAITED);
        sum[2] += at(it, TR_S_IN) - at(it, TR_P_OUT);
        sum[3] += at(it, TR_S_WAITED) - at(it, TR_S_IN);
        sum[4] += at(it, TR_S_OUT) - at(it, TR_S_WAITED);
        sum[5] += at(it, TR_U_IN) - at(it, TR_S_OUT);
        sum[6] += at(it, TR_U_WAITED) - at(it, TR_U_IN);
        sum[7] += at(it, TR_U_OUT) - at(it, TR_U_WAITED);
        sum[8] += at(it + 1, TR_P_IN) - at(it, TR_U_OUT);
      }
      double tot = 0;
      for (double v : sum) tot += v;
      char line[1024];  // one write per rank: the ranks share stderr
      int len = snprintf(line, sizeof line, "[cg trace rank %d] %.1f us/iteration over iterations %d..%d:", H ? H->me : 0,
                         tot / (hi - lo) * 1e-3, lo, hi);
      for (int k = 0; k < 9 && len < (int)sizeof line; ++k)
        len += snprintf(line + len, sizeof line - len, " %s %.1f;", names[k], sum[k] / (hi - lo) * 1e-3);
      fprintf(stderr, "%s\n", line);
    }
  }
  cudaEventDestroy(ev[0]);
  cudaEventDestroy(ev[1]);
  if (!exec_cached) {
    if (exec) cudaGraphExecDestroy(exec);
    if (graph) cudaGraphDestroy(graph);
  }
  return rc;
}

// iife_ksp_solve_hessenberg (row N4: estimateConditionNumber, reference common.py:483-507) asks the FGMRES
// driver to keep the triangular factor of the last cycle's Hessenberg matrix on the host
struct HessenbergKeep {
  bool want = false;
  int m = 0, k = 0;
  std::vector<double> H;  // (m+1) x m column-major, rotations applied (upper triangular in its first k columns)
};
static HessenbergKeep g_hess;

// ------------------------------------------------------------------------------------------------
// FGMRES driver (device pointers)
// ------------------------------------------------------------------------------------------------
static int fgmres_solve(Mat *A, Halo *H, const double *dinv, const double *b, double *x, int64_t max_it, int restart,
                        KspWork &w, HostFlags *hf) {
  Ctx &c = ctx();
  const int64_t n = A->n_rows;
  // row-partitioned (NCCL layer): z_j and x carry ghost entries for the SpMV, every reduction is followed by
  // an allreduce and its scalar step runs in a one-thread kernel afterwards
  const bool dist = (H != nullptr) && c.nranks > 1;
  const int64_t n_ext = H ? H->n_owned + H->n_ghost : n;
  Tmp<double> xe;
  if (H) IIFE_TRY(xe.alloc((size_t)n_ext));
  int m = restart;
  if (m < 1) m = 30;
  if ((int64_t)m > max_it && max_it > 0) m = (int)max_it;
  if (m > 10000) m = 10000;
  const int g = vec_grid(n);
  // small device arrays
  Tmp<double> small;
  size_t hs = (size_t)(m + 1) * m;
  IIFE_TRY(small.alloc(hs + 4 * (size_t)(m + 1)));
  IIFE_CUDA(cudaMemsetAsync(small.p, 0, (hs + 4 * (size_t)(m + 1)) * sizeof(double), c.stream));
  GmresSmall gs;
  gs.H = small.p;
  gs.cs = small.p + hs;
  gs.sn = gs.cs + (m + 1);
  gs.rs = gs.sn + (m + 1);
  gs.y = gs.rs + (m + 1);
  gs.m = m;
  // basis vectors allocated lazily; pointer tables on the device
  std::vector<double *> V, Z;
  Tmp<double *> vtab, ztab;
  Tmp<double> mpart;
  IIFE_TRY(vtab.alloc((size_t)m + 1));
  IIFE_TRY(ztab.alloc((size_t)m + 1));
  IIFE_TRY(mpart.alloc((size_t)(m + 1) * g));
  int rc = IIFE_OK;
  // grow the basis up to V[upto_v], Z[upto_z]; called only at points where the stream is idle
  // grow the basis up to V[upto_v], Z[upto_z] with ONE slab allocation per call; called only at points
  // where the stream is idle (cycle start / chunk polls)
  std::vector<std::pair<double *, size_t>> slabs;
  auto ensure_vecs = [&](int upto_v, int upto_z) -> int {
    int need_v = upto_v + 1 - (int)V.size(), need_z = upto_z + 1 - (int)Z.size();
    if (need_v < 0) need_v = 0;
    if (need_z < 0) need_z = 0;
    if (need_v + need_z == 0) return IIFE_OK;
    size_t n_pad = ((size_t)n_ext + 31) & ~(size_t)31;  // keep every vector 256-byte aligned
    size_t count = (size_t)(need_v + need_z) * n_pad;
    double *slab = nullptr;
    IIFE_TRY(dev_alloc_t(&slab, count));
    slabs.emplace_back(slab, count);
    double *cur = slab;
    for (int k = 0; k < need_v; ++k, cur += n_pad) V.push_back(cur);
    for (int k = 0; k < need_z; ++k, cur += n_pad) Z.push_back(cur);
    IIFE_CUDA(cudaStreamSynchronize(c.stream));
    IIFE_CUDA(cudaMemcpy(vtab.p, V.data(), V.size() * sizeof(double *), cudaMemcpyHostToDevice));
    IIFE_CUDA(cudaMemcpy(ztab.p, Z.data(), Z.size() * sizeof(double *), cudaMemcpyHostToDevice));
    return IIFE_OK;
  };
  auto cleanup = [&]() {
    cudaStreamSynchronize(c.stream);
    for (auto &sl : slabs) dev_free_t(sl.first, sl.second);
  };
  int chunk = env_int("IIFE_KSP_CHUNK", 16);
  if (chunk < 1) chunk = 1;
  bool first_cycle = true;
  int64_t enq_total = 0;
  bool done = false;
  while (rc == IIFE_OK && !done) {
    // cycle start: V0 = b - A x
    if ((rc = ensure_vecs(chunk < m ? chunk : m, (chunk < m ? chunk : m) - 1)) != IIFE_OK) break;
    IIFE_LAUNCH(k_copy_gated, g, VEC_THREADS, 0, b, V[0], n, (const int *)w.fl);
    if (H) {
      IIFE_LAUNCH(k_copy_gated, g, VEC_THREADS, 0, (const double *)x, xe.p, n, (const int *)nullptr);
      if (dist && (rc = halo_exchange(H, xe.p)) != IIFE_OK) break;
      if ((rc = spmv_launch(A, -1.0, xe.p, 1.0, V[0])) != IIFE_OK) break;
    } else if ((rc = spmv_launch(A, -1.0, x, 1.0, V[0])) != IIFE_OK) break;
    if (dist) {
      IIFE_LAUNCH(k_gm_cycle_start<true>, g, VEC_THREADS, 0, V[0], b, n, w.sc, w.fl, gs, w.partials, w.counters, w.hist,
                  (long long)w.hist_len, first_cycle ? 1 : 0);
      if ((rc = allreduce_sum(w.sc + S_RAW, 2)) != IIFE_OK) break;
      IIFE_LAUNCH(k_gm_cycle_scalars, 1, 1, 0, w.sc, w.fl, gs, first_cycle ? 1 : 0, w.hist, (long long)w.hist_len);
    } else {
      IIFE_LAUNCH(k_gm_cycle_start<false>, g, VEC_THREADS, 0, V[0], b, n, w.sc, w.fl, gs, w.partials, w.counters, w.hist,
                  (long long)w.hist_len, first_cycle ? 1 : 0);
    }
    first_cycle = false;
    if ((rc = poll_flags(w, hf)) != IIFE_OK) break;
    if (hf->fl[F_REASON] != 0) { done = true; break; }
    for (int j = 0; j < m && rc == IIFE_OK; ++j) {
      IIFE_LAUNCH(k_gm_scale_pc, g, VEC_THREADS, 0, V[j], Z[j], dinv, n, w.sc, w.fl, j);
      // w = A z_j into V[j+1]  (harmless after convergence: V[j+1] is not read any more)
      if (dist && (rc = halo_exchange(H, Z[j])) != IIFE_OK) break;
      if ((rc = spmv_launch(A, 1.0, Z[j], 0.0, V[j + 1])) != IIFE_OK) break;
      IIFE_LAUNCH(k_gm_dots, g, VEC_THREADS, 0, V[j + 1], (double *const *)vtab.p, n, j, gs, w.fl, mpart.p,
                  w.counters + 1, j);
      if (dist) {
        if ((rc = allreduce_sum(gs.H + (size_t)j * (gs.m + 1), j + 1)) != IIFE_OK) break;
        IIFE_LAUNCH(k_gm_update<true>, g, VEC_THREADS, (size_t)(j + 1) * sizeof(double), V[j + 1], (double *const *)vtab.p,
                    n, j, gs, w.sc, w.fl, w.partials, w.counters, w.hist, (long long)w.hist_len);
        if ((rc = allreduce_sum(w.sc + S_RAW, 1)) != IIFE_OK) break;
        IIFE_LAUNCH(k_gm_update_scalars, 1, 1, 0, gs, w.sc, w.fl, j, w.hist, (long long)w.hist_len);
      } else {
        IIFE_LAUNCH(k_gm_update<false>, g, VEC_THREADS, (size_t)(j + 1) * sizeof(double), V[j + 1], (double *const *)vtab.p,
                    n, j, gs, w.sc, w.fl, w.partials, w.counters, w.hist, (long long)w.hist_len);
      }
      ++enq_total;
      if ((j + 1) % chunk == 0 || j + 1 == m) {
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) { rc = set_err(IIFE_ERR_CUDA, "FGMRES launch: %s", cudaGetErrorString(e)); break; }
        if ((rc = poll_flags(w, hf)) != IIFE_OK) break;
        if (hf->fl[F_REASON] != 0) break;
        if (j + 1 < m) {
          int upto = j + 1 + chunk < m ? j + 1 + chunk : m;
          if ((rc = ensure_vecs(upto, upto - 1)) != IIFE_OK) break;
        }
      }
    }
    if (rc != IIFE_OK) break;
    if (g_hess.want) {  // the stream is idle here (last poll); F_LOC_IT = inner iterations of this cycle
      g_hess.m = m;
      g_hess.k = hf->fl[F_LOC_IT];
      g_hess.H.resize(hs);
      cudaError_t he = cudaMemcpy(g_hess.H.data(), gs.H, hs * sizeof(double), cudaMemcpyDeviceToHost);
      if (he != cudaSuccess) { rc = set_err(IIFE_ERR_CUDA, "Hessenberg readback: %s", cudaGetErrorString(he)); break; }
    }
    // build the solution from however many inner iterations were completed in this cycle
    IIFE_LAUNCH(k_gm_solve_y, 1, 32, 0, gs, w.fl);
    IIFE_LAUNCH(k_gm_build_x, g, VEC_THREADS, (size_t)(m + 1) * sizeof(double), x, (double *const *)ztab.p, n, gs, w.fl,
                w.counters + 2);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { rc = set_err(IIFE_ERR_CUDA, "FGMRES build: %s", cudaGetErrorString(e)); break; }
    if ((rc = poll_flags(w, hf)) != IIFE_OK) break;
    if (hf->fl[F_REASON] != 0) done = true;
    if (enq_total > max_it + m) { rc = set_err(IIFE_ERR_STATE, "FGMRES driver ran past max_it without a reason"); break; }
  }
  cleanup();
  return rc;
}

// ------------------------------------------------------------------------------------------------
// GCR (PETSc's KSPGCR, restart 30, reference common.py:559-560: method='gcr'), single GPU.  Per step k of a cycle:
//   s_k = D^-1 r;  v_k = A s_k;  c_i = (v_k, v_i), i < k (classical Gram-Schmidt: all dots first);
//   v_k -= sum c_i v_i;  s_k -= sum c_i s_i;  nrm = ||v_k||;  v_k /= nrm;  s_k /= nrm;
//   x += (r, v_k) s_k;  r -= (r, v_k) v_k;  test ||r|| (unpreconditioned norm, reference norm ||b||)
// The residual is recomputed from x at the start of every cycle.  Same device-side bookkeeping as FGMRES: F_LOC_IT is
// the step inside the cycle, every kernel is gated by it and by the reason flag, the host polls once per chunk.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(VEC_THREADS)
k_gcr_pc(const double *__restrict__ r, const double *__restrict__ dinv, double *__restrict__ s, int64_t n,
         const int *__restrict__ fl, int k) {
  if (fl[F_REASON] != 0 || fl[F_LOC_IT] != k) return;
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) s[i] = (dinv ? dinv[i] : 1.0) * r[i];
}

// v_k -= sum_{i<k} c_i v_i, s_k -= sum_{i<k} c_i s_i; (r, v_k) and (v_k, v_k); the last CTA turns them into the scale
// 1/||v_k|| (S_SCALE) and the step length (r, v_k)/||v_k|| (S_TT)
__global__ void __launch_bounds__(VEC_THREADS)
k_gcr_update(double *__restrict__ v, double *__restrict__ sv, const double *__restrict__ r, double *const *__restrict__ V,
             double *const *__restrict__ SV, int64_t n, int k, GmresSmall gs, double *sc, const int *__restrict__ fl,
             double *partials, unsigned int *counter) {
  if (fl[F_REASON] != 0 || fl[F_LOC_IT] != k) return;
  extern __shared__ double csh[];  // k coefficients
  __shared__ double red[32];
  __shared__ double out[2];
  __shared__ bool last;
  const double *ck = k > 0 ? gs.H + (size_t)(k - 1) * (gs.m + 1) : nullptr;
  for (int i = threadIdx.x; i < k; i += blockDim.x) csh[i] = ck[i];
  __syncthreads();
  double acc[2] = {0.0, 0.0};
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    double a = v[i], b = sv[i];
    for (int q = 0; q < k; ++q) {
      const double c = csh[q];
      a = fma(-c, V[q][i], a);
      b = fma(-c, SV[q][i], b);
    }
    v[i] = a;
    sv[i] = b;
    acc[0] = fma(r[i], a, acc[0]);
    acc[1] = fma(a, a, acc[1]);
  }
  if (grid_reduce<2>(acc, partials, counter, out, red, &last) && threadIdx.x == 0) {
    const double nrm = sqrt(out[1]);
    sc[S_SCALE] = 1.0 / nrm;  // nrm == 0 (breakdown) gives inf/NaN, which the convergence test reports as NANORINF
    sc[S_TT] = out[0] / nrm;
  }
}

// normalise v_k, s_k; x += tt s_k; r -= tt v_k; ||r|| -> iteration count, history, convergence test
__global__ void __launch_bounds__(VEC_THREADS)
k_gcr_apply(double *__restrict__ v, double *__restrict__ sv, double *__restrict__ x, double *__restrict__ r, int64_t n, int k,
            double *sc, int *fl, double *partials, unsigned int *counter, double *hist, long long hist_len) {
  if (fl[F_REASON] != 0 || fl[F_LOC_IT] != k) return;
  __shared__ double red[32];
  __shared__ double out[1];
  __shared__ bool last;
  const double scale = sc[S_SCALE], tt = sc[S_TT];
  double acc[1] = {0.0};
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const double vi = v[i] * scale, si = sv[i] * scale;
    v[i] = vi;
    sv[i] = si;
    x[i] = fma(tt, si, x[i]);
    const double ri = fma(-tt, vi, r[i]);
    r[i] = ri;
    acc[0] = fma(ri, ri, acc[0]);
  }
  if (grid_reduce<1>(acc, partials, counter, out, red, &last) && threadIdx.x == 0) {
    const double res = sqrt(out[0]);
    const int its = fl[F_ITS] + 1;
    fl[F_ITS] = its;
    fl[F_LOC_IT] = k + 1;
    sc[S_RES] = res;
    log_hist(hist, hist_len, its, res);
    converged_default(sc, fl, its, res);
    if (fl[F_REASON] == 0 && its >= fl[F_MAXIT]) fl[F_REASON] = IIFE_KSP_DIVERGED_ITS;
  }
}

static int gcr_solve(Mat *A, const double *dinv, const double *b, double *x, int64_t max_it, int restart, KspWork &w,
                     HostFlags *hf) {
  Ctx &c = ctx();
  const int64_t n = A->n_rows;
  int m = restart > 0 ? restart : 30;
  if (m > 1000) m = 1000;
  const int g = vec_grid(n);
  Tmp<double> small, r, mpart;
  const size_t hs = (size_t)(m + 1) * m;
  IIFE_TRY(small.alloc(hs + 4 * (size_t)(m + 1)));
  IIFE_CUDA(cudaMemsetAsync(small.p, 0, (hs + 4 * (size_t)(m + 1)) * sizeof(double), c.stream));
  GmresSmall gs;
  gs.H = small.p;
  gs.cs = small.p + hs;
  gs.sn = gs.cs + (m + 1);
  gs.rs = gs.sn + (m + 1);
  gs.y = gs.rs + (m + 1);
  gs.m = m;
  IIFE_TRY(r.alloc((size_t)n));
  IIFE_TRY(mpart.alloc((size_t)(m + 1) * g));
  const size_t n_pad = ((size_t)n + 31) & ~(size_t)31;
  Tmp<double> slab;  // v_0..v_{m-1}, s_0..s_{m-1}
  IIFE_TRY(slab.alloc(2 * (size_t)m * n_pad));
  std::vector<double *> V((size_t)m), SV((size_t)m);
  for (int k = 0; k < m; ++k) {
    V[(size_t)k] = slab.p + (size_t)k * n_pad;
    SV[(size_t)k] = slab.p + (size_t)(m + k) * n_pad;
  }
  Tmp<double *> vtab, stab;
  IIFE_TRY(vtab.alloc((size_t)m));
  IIFE_TRY(stab.alloc((size_t)m));
  IIFE_CUDA(cudaMemcpyAsync(vtab.p, V.data(), (size_t)m * sizeof(double *), cudaMemcpyHostToDevice, c.stream));
  IIFE_CUDA(cudaMemcpyAsync(stab.p, SV.data(), (size_t)m * sizeof(double *), cudaMemcpyHostToDevice, c.stream));
  IIFE_CUDA(cudaStreamSynchronize(c.stream));  // the host tables go out of use
  int chunk = env_int("IIFE_KSP_CHUNK", 10);
  if (chunk < 1) chunk = 1;
  bool first_cycle = true;
  int64_t enq_total = 0;
  for (;;) {
    // cycle start: r = b - A x, ||r|| (and ||b|| the first time), test
    IIFE_LAUNCH(k_copy_gated, g, VEC_THREADS, 0, b, r.p, n, (const int *)w.fl);
    IIFE_TRY(spmv_launch(A, -1.0, x, 1.0, r.p));
    IIFE_LAUNCH(k_gm_cycle_start<false>, g, VEC_THREADS, 0, r.p, b, n, w.sc, w.fl, gs, w.partials, w.counters, w.hist,
                (long long)w.hist_len, first_cycle ? 1 : 0);
    first_cycle = false;
    IIFE_TRY(poll_flags(w, hf));
    if (hf->fl[F_REASON] != 0) return IIFE_OK;
    for (int k = 0; k < m; ++k) {
      IIFE_LAUNCH(k_gcr_pc, g, VEC_THREADS, 0, r.p, dinv, SV[(size_t)k], n, w.fl, k);
      IIFE_TRY(spmv_launch(A, 1.0, SV[(size_t)k], 0.0, V[(size_t)k]));  // harmless after convergence: v_k is not read any more
      if (k > 0)
        IIFE_LAUNCH(k_gm_dots, g, VEC_THREADS, 0, V[(size_t)k], (double *const *)vtab.p, n, k - 1, gs, w.fl, mpart.p,
                    w.counters + 1, k);
      IIFE_LAUNCH(k_gcr_update, g, VEC_THREADS, (size_t)(k + 1) * sizeof(double), V[(size_t)k], SV[(size_t)k], r.p,
                  (double *const *)vtab.p, (double *const *)stab.p, n, k, gs, w.sc, w.fl, w.partials, w.counters);
      IIFE_LAUNCH(k_gcr_apply, g, VEC_THREADS, 0, V[(size_t)k], SV[(size_t)k], x, r.p, n, k, w.sc, w.fl, w.partials,
                  w.counters + 2, w.hist, (long long)w.hist_len);
      ++enq_total;
      if ((k + 1) % chunk == 0 || k + 1 == m) {
        IIFE_CHECK_LAUNCH();
        IIFE_TRY(poll_flags(w, hf));
        if (hf->fl[F_REASON] != 0) return IIFE_OK;
      }
    }
    if (enq_total > max_it + m) return set_err(IIFE_ERR_STATE, "GCR driver ran past max_it without a reason");
  }
}

__global__ void k_ksp_setup(double *sc, int *fl, double rtol, double atol, double dtol, int maxit) {
  for (int k = 0; k < S_COUNT; ++k) sc[k] = 0.0;
  for (int k = 0; k < F_COUNT; ++k) fl[k] = 0;
  sc[S_RTOL] = rtol;
  sc[S_ATOL] = atol;
  sc[S_DTOL] = dtol;
  fl[F_MAXIT] = maxit;
}

}  // namespace iife

using namespace iife;

static int ksp_solve_common(Mat *A, Halo *H, int ksp_type, int pc_type, double rtol, double atol, double dtol,
                            int64_t max_it, int restart, const double *b, double *x, int mem, iife_ksp_result *res,
                            double *hist, int64_t hist_len) {
  Ctx &c = ctx();
  if (!A || !b || !x) return set_err(IIFE_ERR_ARG, "NULL argument");
  if (!H && A->n_rows != A->n_cols) return set_err(IIFE_ERR_ARG, "KSP needs a square operator, got %lld x %lld", (long long)A->n_rows, (long long)A->n_cols);
  if (H && (A->n_rows != H->n_owned || A->n_cols != H->n_owned + H->n_ghost))
    return set_err(IIFE_ERR_ARG, "local operator %lld x %lld does not match the halo (%lld owned + %lld ghost)", (long long)A->n_rows,
                   (long long)A->n_cols, (long long)H->n_owned, (long long)H->n_ghost);
  if (H && mem != IIFE_MEM_DEVICE) return set_err(IIFE_ERR_ARG, "the row-partitioned solver takes device vectors");
  if (ksp_type != IIFE_KSP_CG && ksp_type != IIFE_KSP_FGMRES && ksp_type != IIFE_KSP_GCR)
    return set_err(IIFE_ERR_ARG, "unknown ksp_type %d", ksp_type);
  if (ksp_type == IIFE_KSP_GCR && H) return set_err(IIFE_ERR_UNSUPPORTED, "GCR is not available on the row-partitioned path");
  if (pc_type != IIFE_PC_NONE && pc_type != IIFE_PC_JACOBI) return set_err(IIFE_ERR_ARG, "unknown pc_type %d", pc_type);
  if (max_it < 0) max_it = 0;
  if (max_it > 0x7ffffff0LL) max_it = 0x7ffffff0LL;
  if (dtol <= 0.0) dtol = 1e4;
  const int64_t n = A->n_rows;
  const double *dinv = nullptr;
  // SELL-32 copy of the operator for the iteration (CSR if rejected); its fill pass also yields the Jacobi diagonal
  IIFE_TRY(mat_ensure_sell(A, pc_type == IIFE_PC_JACOBI));
  if (pc_type == IIFE_PC_JACOBI) {
    IIFE_TRY(mat_ensure_dinv(A));  // no-op when the fill pass wrote it
    dinv = A->dinv;
  }
  KspWork w;
  g_arena.cursor = 0;  // the arena serves this solve's requests in order
  Ws<double> sc, partials, dhist, dx, db;
  Ws<int> fl;
  Ws<unsigned int> counters;
  IIFE_TRY(sc.alloc(S_COUNT));
  IIFE_TRY(fl.alloc(F_COUNT));
  IIFE_TRY(partials.alloc(4 * MAX_PARTIALS));
  IIFE_TRY(counters.alloc(4));
  IIFE_CUDA(cudaMemsetAsync(counters.p, 0, 4 * sizeof(unsigned int), c.stream));
  w.sc = sc.p;
  w.fl = fl.p;
  w.partials = partials.p;
  w.counters = counters.p;
  if (hist && hist_len > 0) {
    IIFE_TRY(dhist.alloc((size_t)hist_len));
    IIFE_CUDA(cudaMemsetAsync(dhist.p, 0, (size_t)hist_len * sizeof(double), c.stream));
    w.hist = dhist.p;
    w.hist_len = hist_len;
  }
  IIFE_LAUNCH(k_ksp_setup, 1, 1, 0, w.sc, w.fl, rtol, atol, dtol, (int)max_it);
  IIFE_CHECK_LAUNCH();
  const double *bd = b;
  double *xd = x;
  if (mem == IIFE_MEM_HOST) {
    IIFE_TRY(dx.alloc((size_t)n));
    IIFE_TRY(db.alloc((size_t)n));
    IIFE_CUDA(cudaMemcpyAsync(dx.p, x, (size_t)n * sizeof(double), cudaMemcpyHostToDevice, c.stream));
    IIFE_CUDA(cudaMemcpyAsync(db.p, b, (size_t)n * sizeof(double), cudaMemcpyHostToDevice, c.stream));
    bd = db.p;
    xd = dx.p;
  }
  static HostFlags *hf_cached = nullptr;  // pinned, two slots, kept for the life of the process
  if (!hf_cached) IIFE_CUDA(cudaMallocHost((void **)&hf_cached, 2 * sizeof(HostFlags)));
  HostFlags *hf = hf_cached;
  int rc;
  if (n == 0 && !H) {
    rc = IIFE_OK;
    for (int k = 0; k < F_COUNT; ++k) hf->fl[k] = 0;
    hf->fl[F_REASON] = IIFE_KSP_CONVERGED_ATOL;
  } else if (ksp_type == IIFE_KSP_CG) {
    rc = cg_solve(A, H, dinv, bd, xd, max_it, w, hf);
  } else if (ksp_type == IIFE_KSP_GCR) {
    rc = gcr_solve(A, dinv, bd, xd, max_it, restart, w, hf);
  } else {
    rc = fgmres_solve(A, H, dinv, bd, xd, max_it, restart, w, hf);
  }
  if (rc == IIFE_OK) {
    double hsc[S_COUNT] = {0};
    if (n > 0 || H) {
      cudaMemcpyAsync(hsc, w.sc, sizeof(hsc), cudaMemcpyDeviceToHost, c.stream);
      cudaMemcpyAsync(hf->fl, w.fl, sizeof(int) * F_COUNT, cudaMemcpyDeviceToHost, c.stream);
    }
    if (mem == IIFE_MEM_HOST) cudaMemcpyAsync(x, xd, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost, c.stream);
    if (w.hist) cudaMemcpyAsync(hist, w.hist, (size_t)hist_len * sizeof(double), cudaMemcpyDeviceToHost, c.stream);
    cudaError_t e = cudaStreamSynchronize(c.stream);
    if (e != cudaSuccess) rc = set_err(IIFE_ERR_CUDA, "KSP finish: %s", cudaGetErrorString(e));
    if (res) {
      res->iterations = hf->fl[F_ITS];
      res->reason = hf->fl[F_REASON];
      res->_pad = 0;
      res->rnorm = ksp_type == IIFE_KSP_CG ? hsc[S_DP] : hsc[S_RES];
      res->rnorm0 = hsc[S_RHO0];
    }
  }
  return rc;
}

extern "C" int iife_ksp_solve(iife_mat A_, int ksp_type, int pc_type, double rtol, double atol, double dtol,
                              int64_t max_it, int restart, const double *b, double *x, int mem, iife_halo halo,
                              iife_ksp_result *res, double *hist, int64_t hist_len) {
  IIFE_NEED_INIT();
  return ksp_solve_common((Mat *)A_, (Halo *)halo, ksp_type, pc_type, rtol, atol, dtol, max_it, restart, b, x, mem, res,
                          hist, hist_len);
}

extern "C" int iife_ksp_solve_hessenberg(iife_mat A_, int pc_type, double rtol, double atol, double dtol, int64_t max_it,
                                         int restart, const double *b, double *x, int mem, iife_ksp_result *res,
                                         double *R, int64_t r_capacity, int64_t *k_out) {
  IIFE_NEED_INIT();
  if (!k_out) return set_err(IIFE_ERR_ARG, "k_out is NULL");
  *k_out = 0;
  g_hess.want = true;
  g_hess.k = 0;
  int rc = ksp_solve_common((Mat *)A_, nullptr, IIFE_KSP_FGMRES, pc_type, rtol, atol, dtol, max_it, restart, b, x, mem, res,
                            nullptr, 0);
  g_hess.want = false;
  if (rc != IIFE_OK) return rc;
  const int64_t k = g_hess.k;
  *k_out = k;
  if (k > 0 && R) {
    if (r_capacity < k * k) return set_err(IIFE_ERR_ARG, "R holds %lld doubles, %lld needed", (long long)r_capacity, (long long)(k * k));
    const int m1 = g_hess.m + 1;
    for (int64_t c = 0; c < k; ++c)
      for (int64_t i = 0; i < k; ++i) R[c * k + i] = (i <= c) ? g_hess.H[(size_t)c * m1 + (size_t)i] : 0.0;
  }
  return IIFE_OK;
}

extern "C" int iife_ksp_solve_dist(iife_mat A_local, iife_halo H, int ksp_type, int pc_type, double rtol, double atol,
                                   double dtol, int64_t max_it, int restart, const double *b_dev, double *x_dev,
                                   iife_ksp_result *res, double *hist, int64_t hist_len) {
  IIFE_NEED_INIT();
  if (!H) return set_err(IIFE_ERR_ARG, "halo is NULL");
  return ksp_solve_common((Mat *)A_local, (Halo *)H, ksp_type, pc_type, rtol, atol, dtol, max_it, restart, b_dev, x_dev,
                          IIFE_MEM_DEVICE, res, hist, hist_len);
}
