// p2p.cu — NVLink peer-memory collectives of the row-partitioned CG (one process per GPU).
//
// With NCCL, every CG iteration pays three collective launches (halo send/recv, two allreduces of 1-2
// doubles): ~100 us of latency per iteration against ~70 us of SpMV + vector work per GPU at 8 GPUs
// (bench r01: 39 ms of 45 ms per step).  Here the two exchanges are kernels of this library that store
// straight into the peers' memory (cudaIpc mappings over NVLink/NVSwitch):
//   * k_halo_xchg   every send entry of p is stored into the neighbour's ghost slot of ITS p vector; the
//                   last CTA raises a sequence flag in each neighbour's mailbox and waits for the flags of
//                   the ranks this rank receives from.
//   * k_p2p_allreduce  each rank stores its partial sums into every rank's mailbox, raises a flag, waits
//                   for all flags and adds the partials in rank order — every rank gets the bit-identical
//                   sum (needed: the convergence decision must agree on all ranks).
// No NCCL call is left inside the iteration, so a chunk of iterations is one CUDA graph.  Ordering between
// iterations needs no extra barrier: a rank's halo stores of iteration i+1 are issued after it passed the
// allreduce of iteration i, which every peer only enters after its SpMV has consumed the ghosts of i.
// Spins are bounded (about 2 s of clock64) and raise an error flag instead of hanging the GPU.
#include "common.cuh"
#include "p2p_dev.cuh"
#include <string.h>
#include <algorithm>

namespace iife {

__global__ void __launch_bounds__(256)
k_halo_xchg(const double *__restrict__ x, const int *__restrict__ send_idx, const unsigned char *__restrict__ send_peer,
            const int *__restrict__ send_off, long long total_send, PeerTable pt, Mailbox *mbox, int me, int nranks,
            unsigned int send_mask, unsigned int recv_mask, unsigned long long *seq_ptr, unsigned int *counter, int *err,
            const int *__restrict__ reason) {
  if (reason && *reason != 0) return;
  __shared__ bool last;
  const unsigned long long seq = *seq_ptr + 1;
  long long stride = (long long)gridDim.x * blockDim.x;
  for (long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x; k < total_send; k += stride) {
    int q = send_peer[k];
    pt.xbuf[q][pt.dst_start[q] + (k - send_off[q])] = x[send_idx[k]];
  }
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned int t = atomicAdd(counter, 1u);
    last = (t == gridDim.x - 1);
  }
  __syncthreads();
  if (!last) return;
  __threadfence_system();
  int q = threadIdx.x;
  if (q < nranks && ((send_mask >> q) & 1u)) st_flag(&pt.mbox[q]->halo_flag[me], seq);
  if (q < nranks && ((recv_mask >> q) & 1u)) spin_until(&mbox->halo_flag[q], seq, err);
  __syncthreads();
  if (threadIdx.x == 0) {
    *seq_ptr = seq;
    *counter = 0u;
  }
  __threadfence_system();
}

__global__ void k_p2p_allreduce(double *vals, int n, PeerTable pt, Mailbox *mbox, int me, int nranks,
                                unsigned long long *seq_ptr, int *err, const int *__restrict__ reason) {
  if (reason && *reason != 0) return;
  const unsigned long long seq = *seq_ptr + 1;
  const int par = (int)(seq & 1ull);
  int q = threadIdx.x;
  if (q < nranks) {
    for (int i = 0; i < n; ++i) pt.mbox[q]->ar_vals[par][me][i] = vals[i];
    __threadfence_system();
    st_flag(&pt.mbox[q]->ar_flag[par][me], seq);
    spin_until(&mbox->ar_flag[par][q], seq, err);
  }
  __syncthreads();
  __threadfence_system();
  if (threadIdx.x == 0) {
    for (int i = 0; i < n; ++i) {
      double s = 0.0;
      for (int r = 0; r < nranks; ++r) s += ((volatile double *)mbox->ar_vals[par][r])[i];
      vals[i] = s;
    }
    *seq_ptr = seq;
  }
}

static PeerTable make_table(const Halo *H) {
  PeerTable pt;
  for (int q = 0; q < P2P_MAX_RANKS; ++q) {
    pt.xbuf[q] = H->peer_xbuf[q];
    pt.mbox[q] = H->peer_mbox[q];
    pt.dst_start[q] = H->dst_start[q];
  }
  return pt;
}

int p2p_halo_exchange(Halo *H, const int *reason_flag) {
  Ctx &c = ctx();
  int64_t g = (H->total_send + 255) / 256;
  if (g > c.sm_count) g = c.sm_count;
  if (g < 1) g = 1;
  IIFE_LAUNCH(k_halo_xchg, (int)g, 256, 0, H->xbuf, H->send_idx, H->send_peer, H->send_off_dev, (long long)H->total_send,
              make_table(H), H->mbox, H->me, H->nranks, H->send_mask, H->recv_mask, H->dev_seq, H->p2p_counter, H->p2p_err,
              reason_flag);
  IIFE_CHECK_LAUNCH();
  return IIFE_OK;
}

int p2p_allreduce(Halo *H, double *vals, int n, const int *reason_flag) {
  if (n > 4) return set_err(IIFE_ERR_ARG, "p2p_allreduce handles at most 4 values");
  IIFE_LAUNCH(k_p2p_allreduce, 1, 32, 0, vals, n, make_table(H), H->mbox, H->me, H->nranks, H->dev_seq + 1, H->p2p_err,
              reason_flag);
  IIFE_CHECK_LAUNCH();
  return IIFE_OK;
}

}  // namespace iife

using namespace iife;

extern "C" {

// Allocates the peer-visible vector and mailbox of a halo and returns their IPC handles (2 x 64 bytes).
int iife_halo_p2p_export(iife_halo H_, void *handles128) {
  IIFE_NEED_INIT();
  Halo *H = (Halo *)H_;
  if (!H || !handles128) return set_err(IIFE_ERR_ARG, "NULL argument");
  if (H->nranks > P2P_MAX_RANKS) return set_err(IIFE_ERR_UNSUPPORTED, "peer-memory path supports at most %d ranks", P2P_MAX_RANKS);
  Ctx &c = ctx();
  if (!H->xbuf) {
    // IPC handles need whole cudaMalloc allocations: bypass the caching allocator
    size_t nx = (size_t)(H->n_owned + H->n_ghost) * sizeof(double);
    if (nx < 256) nx = 256;
    IIFE_CUDA(cudaMalloc((void **)&H->xbuf, nx));
    IIFE_CUDA(cudaMemset(H->xbuf, 0, nx));
    IIFE_CUDA(cudaMalloc((void **)&H->mbox, sizeof(Mailbox)));
    IIFE_CUDA(cudaMemset(H->mbox, 0, sizeof(Mailbox)));
    IIFE_CUDA(cudaMalloc((void **)&H->dev_seq, 4 * sizeof(unsigned long long)));
    IIFE_CUDA(cudaMemset(H->dev_seq, 0, 4 * sizeof(unsigned long long)));
    IIFE_CUDA(cudaMalloc((void **)&H->p2p_counter, sizeof(unsigned int)));
    IIFE_CUDA(cudaMemset(H->p2p_counter, 0, sizeof(unsigned int)));
    IIFE_CUDA(cudaMalloc((void **)&H->p2p_err, sizeof(int)));
    IIFE_CUDA(cudaMemset(H->p2p_err, 0, sizeof(int)));
  }
  cudaIpcMemHandle_t hx, hm;
  IIFE_CUDA(cudaIpcGetMemHandle(&hx, H->xbuf));
  IIFE_CUDA(cudaIpcGetMemHandle(&hm, H->mbox));
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  memcpy(handles128, &hx, 64);
  memcpy((char *)handles128 + 64, &hm, 64);
  H->me = c.rank;
  return IIFE_OK;
}

// all_handles: nranks x 128 bytes (rank order); dst_start[q]: offset inside rank q's vector where this
// rank's send block to q starts.  After this call the solver uses the peer-memory path for this halo.
int iife_halo_p2p_attach(iife_halo H_, const void *all_handles, const int64_t *dst_start) {
  IIFE_NEED_INIT();
  Halo *H = (Halo *)H_;
  if (!H || !all_handles || !dst_start) return set_err(IIFE_ERR_ARG, "NULL argument");
  if (!H->xbuf) return set_err(IIFE_ERR_STATE, "call iife_halo_p2p_export first");
  Ctx &c = ctx();
  H->send_mask = H->recv_mask = 0;
  for (int q = 0; q < H->nranks; ++q) {
    H->dst_start[q] = dst_start[q];
    if (q == c.rank) {
      H->peer_xbuf[q] = H->xbuf;
      H->peer_mbox[q] = H->mbox;
      continue;
    }
    cudaIpcMemHandle_t hx, hm;
    memcpy(&hx, (const char *)all_handles + (size_t)q * 128, 64);
    memcpy(&hm, (const char *)all_handles + (size_t)q * 128 + 64, 64);
    void *px = nullptr, *pm = nullptr;
    IIFE_CUDA(cudaIpcOpenMemHandle(&px, hx, cudaIpcMemLazyEnablePeerAccess));
    IIFE_CUDA(cudaIpcOpenMemHandle(&pm, hm, cudaIpcMemLazyEnablePeerAccess));
    H->peer_xbuf[q] = (double *)px;
    H->peer_mbox[q] = (Mailbox *)pm;
    if (H->send_counts[q] > 0) H->send_mask |= 1u << q;
    if (H->recv_counts[q] > 0) H->recv_mask |= 1u << q;
  }
  // per-entry destination rank + device copy of the send offsets
  std::vector<unsigned char> peer((size_t)H->total_send);
  std::vector<int> off(H->nranks + 1);
  for (int q = 0; q < H->nranks; ++q) {
    off[q] = (int)H->send_off[q];
    for (int64_t k = H->send_off[q]; k < H->send_off[q + 1]; ++k) peer[(size_t)k] = (unsigned char)q;
  }
  off[H->nranks] = (int)H->send_off[H->nranks];
  IIFE_CUDA(cudaMalloc((void **)&H->send_peer, peer.size() ? peer.size() : 1));
  IIFE_CUDA(cudaMalloc((void **)&H->send_off_dev, off.size() * sizeof(int)));
  if (!peer.empty()) IIFE_CUDA(cudaMemcpy(H->send_peer, peer.data(), peer.size(), cudaMemcpyHostToDevice));
  IIFE_CUDA(cudaMemcpy(H->send_off_dev, off.data(), off.size() * sizeof(int), cudaMemcpyHostToDevice));
  // push plan by owned row (three-kernel CG iteration): invert send_idx
  {
    std::vector<int> idx((size_t)H->total_send);
    if (H->total_send) IIFE_CUDA(cudaMemcpy(idx.data(), H->send_idx, idx.size() * sizeof(int), cudaMemcpyDeviceToHost));
    struct Ent { int row; unsigned char peer; long long dst; };
    std::vector<Ent> ent((size_t)H->total_send);
    for (int64_t k = 0; k < H->total_send; ++k) {
      const int q = peer[(size_t)k];
      ent[(size_t)k] = {idx[(size_t)k], (unsigned char)q, (long long)H->dst_start[q] + (k - H->send_off[q])};
    }
    std::stable_sort(ent.begin(), ent.end(), [](const Ent &a, const Ent &b) { return a.row < b.row; });
    std::vector<int> brow, bptr;
    std::vector<unsigned char> bpeer(ent.size());
    std::vector<long long> bdst(ent.size());
    // per 32 owned rows: (bit mask of the boundary rows, number of boundary rows before the word) -> index into bptr by
    // one popcount, no search
    std::vector<unsigned int> bmask(2 * ((size_t)(H->n_owned + 31) / 32 + 1), 0u);
    for (size_t k = 0; k < ent.size(); ++k) {
      if (k == 0 || ent[k].row != ent[k - 1].row) {
        brow.push_back(ent[k].row);
        bptr.push_back((int)k);
        bmask[2 * ((size_t)ent[k].row >> 5)] |= 1u << (ent[k].row & 31);
      }
      bpeer[k] = ent[k].peer;
      bdst[k] = ent[k].dst;
    }
    bptr.push_back((int)ent.size());
    for (size_t w = 0, before = 0; 2 * w < bmask.size(); ++w) {
      bmask[2 * w + 1] = (unsigned int)before;
      before += (size_t)__builtin_popcount(bmask[2 * w]);
    }
    H->n_brow = (int)brow.size();
    IIFE_CUDA(cudaMalloc((void **)&H->brow, (brow.size() + 1) * sizeof(int)));
    IIFE_CUDA(cudaMalloc((void **)&H->bptr, bptr.size() * sizeof(int)));
    IIFE_CUDA(cudaMalloc((void **)&H->bpeer, bpeer.size() + 1));
    IIFE_CUDA(cudaMalloc((void **)&H->bdst, (bdst.size() + 1) * sizeof(long long)));
    IIFE_CUDA(cudaMalloc((void **)&H->bmask, bmask.size() * sizeof(unsigned int)));
    if (!brow.empty()) IIFE_CUDA(cudaMemcpy(H->brow, brow.data(), brow.size() * sizeof(int), cudaMemcpyHostToDevice));
    IIFE_CUDA(cudaMemcpy(H->bptr, bptr.data(), bptr.size() * sizeof(int), cudaMemcpyHostToDevice));
    if (!bpeer.empty()) IIFE_CUDA(cudaMemcpy(H->bpeer, bpeer.data(), bpeer.size(), cudaMemcpyHostToDevice));
    if (!bdst.empty()) IIFE_CUDA(cudaMemcpy(H->bdst, bdst.data(), bdst.size() * sizeof(long long), cudaMemcpyHostToDevice));
    IIFE_CUDA(cudaMemcpy(H->bmask, bmask.data(), bmask.size() * sizeof(unsigned int), cudaMemcpyHostToDevice));
  }
  H->p2p = true;
  return IIFE_OK;
}

int iife_halo_p2p_error(iife_halo H_, int *err) {
  IIFE_NEED_INIT();
  Halo *H = (Halo *)H_;
  if (!H || !err) return set_err(IIFE_ERR_ARG, "NULL argument");
  *err = 0;
  if (H->p2p_err) IIFE_CUDA(cudaMemcpy(err, H->p2p_err, sizeof(int), cudaMemcpyDeviceToHost));
  return IIFE_OK;
}

}  // extern "C"
