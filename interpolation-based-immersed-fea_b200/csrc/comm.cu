// comm.cu — multi-GPU plumbing of the row-partitioned path (one process per GPU): NCCL communicator
// bootstrap, ghost-entry (halo) exchange for the SpMV, fp64 sum-allreduce for the dots, and a
// variable-size block exchange used for ghost rows of A_f / M in the distributed PtAP.
//
// This is the B200 counterpart of the reference's only parallel mechanism, PETSc's MPIAIJ row-block
// decomposition with VecScatter ghost updates and MPI_Allreduce dots (reference la_utils.py:116-125,
// common.py:673-677).  NCCL is resolved at run time with dlopen so that the
// library loads (and its host-only entry points work) on machines without NCCL or a GPU; inside a
// torch process the already-loaded torch-bundled libnccl.so.2 is picked up.
#include "common.cuh"
#include <dlfcn.h>
#include <string.h>

namespace iife {

// minimal NCCL ABI (stable across 2.x)
typedef struct ncclComm *ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
typedef int ncclResult_t;
enum { ncclInt8 = 0, ncclChar = 0, ncclFloat64 = 8 };
enum { ncclSum = 0 };

struct NcclApi {
  void *handle = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllReduce)(const void *, void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Send)(const void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Recv)(void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  const char *(*GetErrorString)(ncclResult_t) = nullptr;
};
static NcclApi g_nccl;

static int nccl_load() {
  if (g_nccl.handle) return IIFE_OK;
  const char *names[] = {"libnccl.so.2", "libnccl.so"};
  void *h = nullptr;
  for (const char *nm : names) {
    h = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
    if (h) break;
  }
  if (!h) return set_err(IIFE_ERR_COMM, "cannot dlopen libnccl.so.2: %s", dlerror());
#define LOAD(field, sym)                                                    \
  *(void **)(&g_nccl.field) = dlsym(h, sym);                                \
  if (!g_nccl.field) return set_err(IIFE_ERR_COMM, "NCCL symbol %s missing", sym);
  LOAD(GetUniqueId, "ncclGetUniqueId")
  LOAD(CommInitRank, "ncclCommInitRank")
  LOAD(CommDestroy, "ncclCommDestroy")
  LOAD(AllReduce, "ncclAllReduce")
  LOAD(Send, "ncclSend")
  LOAD(Recv, "ncclRecv")
  LOAD(GroupStart, "ncclGroupStart")
  LOAD(GroupEnd, "ncclGroupEnd")
  LOAD(GetErrorString, "ncclGetErrorString")
#undef LOAD
  g_nccl.handle = h;
  return IIFE_OK;
}

#define IIFE_NCCL(expr)                                                                                  \
  do {                                                                                                   \
    ncclResult_t _r = (expr);                                                                            \
    if (_r != 0) return set_err(IIFE_ERR_COMM, "%s:%d: %s: %s", __FILE__, __LINE__, #expr, g_nccl.GetErrorString(_r)); \
  } while (0)

__global__ void k_pack(const double *__restrict__ x, const int *__restrict__ idx, double *__restrict__ buf, int64_t n) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) buf[i] = x[idx[i]];
}

int halo_exchange(Halo *H, double *x_dev) {
  Ctx &c = ctx();
  if (H->nranks == 1) return IIFE_OK;
  if (!c.nccl_comm) return set_err(IIFE_ERR_COMM, "communicator not initialised");
  if (H->total_send > 0) {
    int64_t g = (H->total_send + 255) / 256;
    if (g > c.sm_count * 8) g = c.sm_count * 8;
    IIFE_LAUNCH(k_pack, (int)g, 256, 0, x_dev, H->send_idx, H->send_buf, H->total_send);
    IIFE_CHECK_LAUNCH();
  }
  ncclComm_t comm = (ncclComm_t)c.nccl_comm;
  IIFE_NCCL(g_nccl.GroupStart());
  for (int p = 0; p < H->nranks; ++p) {
    if (p == c.rank) continue;
    if (H->send_counts[p] > 0)
      IIFE_NCCL(g_nccl.Send(H->send_buf + H->send_off[p], (size_t)H->send_counts[p], ncclFloat64, p, comm, c.stream));
    if (H->recv_counts[p] > 0)
      IIFE_NCCL(g_nccl.Recv(x_dev + H->n_owned + H->recv_off[p], (size_t)H->recv_counts[p], ncclFloat64, p, comm, c.stream));
  }
  IIFE_NCCL(g_nccl.GroupEnd());
  return IIFE_OK;
}

int allreduce_sum(double *buf, int64_t n) {
  Ctx &c = ctx();
  if (c.nranks == 1 || n == 0) return IIFE_OK;
  if (!c.nccl_comm) return set_err(IIFE_ERR_COMM, "communicator not initialised");
  IIFE_NCCL(g_nccl.AllReduce(buf, buf, (size_t)n, ncclFloat64, ncclSum, (ncclComm_t)c.nccl_comm, c.stream));
  return IIFE_OK;
}

}  // namespace iife

using namespace iife;

extern "C" {

int iife_comm_unique_id(void *id128) {
  if (!id128) return set_err(IIFE_ERR_ARG, "id128 is NULL");
  IIFE_TRY(nccl_load());
  ncclUniqueId id;
  IIFE_NCCL(g_nccl.GetUniqueId(&id));
  memcpy(id128, &id, 128);
  return IIFE_OK;
}

int iife_comm_init(int rank, int nranks, const void *id128) {
  IIFE_NEED_INIT();
  Ctx &c = ctx();
  if (nranks < 1 || rank < 0 || rank >= nranks) return set_err(IIFE_ERR_ARG, "bad rank %d / nranks %d", rank, nranks);
  if (c.nccl_comm) return set_err(IIFE_ERR_STATE, "communicator already initialised");
  if (nranks == 1) {
    c.rank = 0;
    c.nranks = 1;
    return IIFE_OK;
  }
  if (!id128) return set_err(IIFE_ERR_ARG, "id128 is NULL");
  IIFE_TRY(nccl_load());
  ncclUniqueId id;
  memcpy(&id, id128, 128);
  ncclComm_t comm = nullptr;
  IIFE_NCCL(g_nccl.CommInitRank(&comm, nranks, id, rank));
  c.nccl_comm = comm;
  c.rank = rank;
  c.nranks = nranks;
  return IIFE_OK;
}

int iife_comm_finalize(void) {
  Ctx &c = ctx();
  if (c.nccl_comm) {
    if (c.init) cudaStreamSynchronize(c.stream);
    g_nccl.CommDestroy((ncclComm_t)c.nccl_comm);
    c.nccl_comm = nullptr;
  }
  c.rank = 0;
  c.nranks = 1;
  return IIFE_OK;
}

int iife_comm_info(int *rank, int *nranks) {
  if (rank) *rank = ctx().rank;
  if (nranks) *nranks = ctx().nranks;
  return IIFE_OK;
}

int iife_halo_create(int64_t n_owned, int64_t n_ghost, const int64_t *send_counts, const int32_t *send_idx,
                     const int64_t *recv_counts, iife_halo *out) {
  IIFE_NEED_INIT();
  Ctx &c = ctx();
  if (!out) return set_err(IIFE_ERR_ARG, "out is NULL");
  *out = nullptr;
  if (n_owned < 0 || n_ghost < 0) return set_err(IIFE_ERR_ARG, "negative sizes");
  Halo *H = new Halo();
  H->n_owned = n_owned;
  H->n_ghost = n_ghost;
  H->nranks = c.nranks;
  H->send_counts.assign(c.nranks, 0);
  H->recv_counts.assign(c.nranks, 0);
  H->send_off.assign(c.nranks + 1, 0);
  H->recv_off.assign(c.nranks + 1, 0);
  for (int p = 0; p < c.nranks; ++p) {
    H->send_counts[p] = send_counts ? send_counts[p] : 0;
    H->recv_counts[p] = recv_counts ? recv_counts[p] : 0;
    if (H->send_counts[p] < 0 || H->recv_counts[p] < 0 || (p == c.rank && (H->send_counts[p] || H->recv_counts[p]))) {
      delete H;
      return set_err(IIFE_ERR_ARG, "bad halo counts for peer %d", p);
    }
    H->send_off[p + 1] = H->send_off[p] + H->send_counts[p];
    H->recv_off[p + 1] = H->recv_off[p] + H->recv_counts[p];
  }
  if (H->recv_off[c.nranks] != n_ghost) {
    delete H;
    return set_err(IIFE_ERR_ARG, "recv counts sum to %lld, n_ghost is %lld", (long long)H->recv_off[c.nranks], (long long)n_ghost);
  }
  H->total_send = H->send_off[c.nranks];
  int rc = dev_alloc_t(&H->send_idx, (size_t)H->total_send);
  if (rc == IIFE_OK) rc = dev_alloc_t(&H->send_buf, (size_t)H->total_send);
  if (rc == IIFE_OK && H->total_send > 0) {
    if (!send_idx) rc = set_err(IIFE_ERR_ARG, "send_idx is NULL");
    else {
      for (int64_t k = 0; k < H->total_send; ++k)
        if (send_idx[k] < 0 || send_idx[k] >= n_owned) { rc = set_err(IIFE_ERR_ARG, "send_idx[%lld] out of the owned range", (long long)k); break; }
      if (rc == IIFE_OK) {
        cudaError_t e = cudaMemcpy(H->send_idx, send_idx, (size_t)H->total_send * sizeof(int), cudaMemcpyHostToDevice);
        if (e != cudaSuccess) rc = set_err(IIFE_ERR_CUDA, "halo upload: %s", cudaGetErrorString(e));
      }
    }
  }
  if (rc != IIFE_OK) {
    if (H->send_idx) dev_free_t(H->send_idx, (size_t)H->total_send);
    if (H->send_buf) dev_free_t(H->send_buf, (size_t)H->total_send);
    delete H;
    return rc;
  }
  *out = (iife_halo)H;
  return IIFE_OK;
}

int iife_halo_exchange(iife_halo H_, double *x_dev) {
  IIFE_NEED_INIT();
  if (!H_ || !x_dev) return set_err(IIFE_ERR_ARG, "NULL argument");
  return halo_exchange((Halo *)H_, x_dev);
}

int iife_halo_destroy(iife_halo H_) {
  Halo *H = (Halo *)H_;
  if (!H) return IIFE_OK;
  if (ctx().init) cudaStreamSynchronize(ctx().stream);
  dev_free_t(H->send_idx, (size_t)H->total_send);
  dev_free_t(H->send_buf, (size_t)H->total_send);
  // peer-memory path (p2p.cu): plain cudaMalloc allocations and IPC mappings
  for (int q = 0; q < P2P_MAX_RANKS; ++q) {
    if (H->peer_xbuf[q] && H->peer_xbuf[q] != H->xbuf) cudaIpcCloseMemHandle(H->peer_xbuf[q]);
    if (H->peer_mbox[q] && H->peer_mbox[q] != H->mbox) cudaIpcCloseMemHandle(H->peer_mbox[q]);
  }
  cudaFree(H->send_peer);
  cudaFree(H->send_off_dev);
  cudaFree(H->brow);
  cudaFree(H->bptr);
  cudaFree(H->bpeer);
  cudaFree(H->bdst);
  cudaFree(H->bmask);
  cudaFree(H->xbuf);
  cudaFree(H->mbox);
  cudaFree(H->dev_seq);
  cudaFree(H->p2p_counter);
  cudaFree(H->p2p_err);
  cudaGetLastError();
  delete H;
  return IIFE_OK;
}

int iife_spmv_dist(iife_mat A_, iife_halo H_, double *x_dev, double *y_dev) {
  IIFE_NEED_INIT();
  Mat *A = (Mat *)A_;
  Halo *H = (Halo *)H_;
  if (!A || !H || !x_dev || !y_dev) return set_err(IIFE_ERR_ARG, "NULL argument");
  // rows may differ from the owned column block (M is n_f_local x n_b): only the column space must match
  if (A->n_cols != H->n_owned + H->n_ghost)
    return set_err(IIFE_ERR_ARG, "operator %lld x %lld does not match the halo (%lld owned + %lld ghost)", (long long)A->n_rows,
                   (long long)A->n_cols, (long long)H->n_owned, (long long)H->n_ghost);
  IIFE_TRY(halo_exchange(H, x_dev));
  return spmv_launch(A, 1.0, x_dev, 0.0, y_dev);
}

int iife_allreduce_sum(double *buf_dev, int64_t n) {
  IIFE_NEED_INIT();
  if (!buf_dev && n) return set_err(IIFE_ERR_ARG, "NULL buffer");
  return allreduce_sum(buf_dev, n);
}

int iife_alltoallv_bytes(const void *send_dev, const int64_t *send_displs, void *recv_dev, const int64_t *recv_displs) {
  IIFE_NEED_INIT();
  Ctx &c = ctx();
  if (!send_displs || !recv_displs) return set_err(IIFE_ERR_ARG, "NULL displacements");
  if (c.nranks == 1) {
    int64_t nb = send_displs[1] - send_displs[0];
    if (nb != recv_displs[1] - recv_displs[0]) return set_err(IIFE_ERR_ARG, "self block size mismatch");
    if (nb) IIFE_CUDA(cudaMemcpyAsync((char *)recv_dev + recv_displs[0], (const char *)send_dev + send_displs[0], (size_t)nb, cudaMemcpyDeviceToDevice, c.stream));
    return IIFE_OK;
  }
  if (!c.nccl_comm) return set_err(IIFE_ERR_COMM, "communicator not initialised");
  ncclComm_t comm = (ncclComm_t)c.nccl_comm;
  // own block: local copy
  {
    int64_t nb = send_displs[c.rank + 1] - send_displs[c.rank];
    if (nb != recv_displs[c.rank + 1] - recv_displs[c.rank]) return set_err(IIFE_ERR_ARG, "self block size mismatch");
    if (nb) IIFE_CUDA(cudaMemcpyAsync((char *)recv_dev + recv_displs[c.rank], (const char *)send_dev + send_displs[c.rank], (size_t)nb, cudaMemcpyDeviceToDevice, c.stream));
  }
  IIFE_NCCL(g_nccl.GroupStart());
  for (int p = 0; p < c.nranks; ++p) {
    if (p == c.rank) continue;
    int64_t ns = send_displs[p + 1] - send_displs[p], nr = recv_displs[p + 1] - recv_displs[p];
    if (ns > 0) IIFE_NCCL(g_nccl.Send((const char *)send_dev + send_displs[p], (size_t)ns, ncclInt8, p, comm, c.stream));
    if (nr > 0) IIFE_NCCL(g_nccl.Recv((char *)recv_dev + recv_displs[p], (size_t)nr, ncclInt8, p, comm, c.stream));
  }
  IIFE_NCCL(g_nccl.GroupEnd());
  return IIFE_OK;
}

}  // extern "C"
