// ctx.cu — process context (one GPU per process), error strings, device allocation accounting and
// the device-wide exclusive scan used by the symbolic phases.
#include "common.cuh"
#include <algorithm>
#include <map>
#include <unordered_map>

namespace iife {

static Ctx g_ctx;
Ctx &ctx() { return g_ctx; }

static thread_local char g_err[1024] = "";
static std::unordered_map<void *, size_t> g_block_sizes;

int set_err(int code, const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}
const char *get_err() { return g_err; }

// ------------------------------------------------------------------ caching device allocator
// cudaMalloc / cudaFree cost tens of microseconds to milliseconds each and cudaFree synchronises the
// device, which would put a host round trip into every call of the hot path (work vectors of the KSP,
// temporaries of the symbolic phase, the FGMRES basis).  Freed blocks are therefore kept in a
// size-indexed free list and reused; everything runs on ONE stream, so reuse is stream-ordered and
// needs no event.  Blocks go back to the driver at iife_finalize or when cudaMalloc runs out.
static std::multimap<size_t, void *> g_free_blocks;
static int64_t g_cached_bytes = 0;

static size_t round_size(size_t bytes) {
  if (bytes < 512) return 512;
  if (bytes < (1u << 20)) return (bytes + 511) & ~(size_t)511;
  return (bytes + ((size_t)2 << 20) - 1) & ~(((size_t)2 << 20) - 1);
}

void dev_release_cached() {
  for (auto &kv : g_free_blocks) cudaFree(kv.second);
  g_free_blocks.clear();
  g_cached_bytes = 0;
}

int dev_alloc(void **p, size_t bytes) {
  *p = nullptr;
  size_t want = round_size(bytes);
  auto it = g_free_blocks.lower_bound(want);
  // a cached block may be up to 25 % larger than the request; small requests take up to 4x (at most 8 MB more): the
  // arrays of a small plan then find their blocks again after a plan of slightly different sizes was freed, instead of
  // going to cudaMalloc (0.1-1 ms each: the spread of the cold times of configs 1-4)
  const size_t slack = std::max(want / 4, std::min(want * 3, (size_t)8 << 20));
  if (it != g_free_blocks.end() && it->first <= want + slack) {
    *p = it->second;
    g_cached_bytes -= (int64_t)it->first;
    g_ctx.dev_bytes += (int64_t)it->first;
    g_block_sizes[*p] = it->first;
    g_free_blocks.erase(it);
    return IIFE_OK;
  }
  cudaError_t e = cudaMalloc(p, want);
  if (e != cudaSuccess) {
    cudaGetLastError();
    cudaStreamSynchronize(g_ctx.stream);
    dev_release_cached();
    e = cudaMalloc(p, want);
  }
  if (e != cudaSuccess) {
    cudaGetLastError();
    *p = nullptr;
    return set_err(IIFE_ERR_NOMEM, "cudaMalloc of %zu bytes failed: %s (library holds %lld bytes)", want,
                   cudaGetErrorString(e), (long long)g_ctx.dev_bytes);
  }
  g_ctx.dev_bytes += (int64_t)want;
  g_block_sizes[*p] = want;
  return IIFE_OK;
}

int dev_free(void *p, size_t bytes) {
  (void)bytes;
  if (!p) return IIFE_OK;
  auto it = g_block_sizes.find(p);
  if (it == g_block_sizes.end()) return set_err(IIFE_ERR_STATE, "dev_free of an unknown pointer");
  size_t sz = it->second;
  g_block_sizes.erase(it);
  g_ctx.dev_bytes -= (int64_t)sz;
  g_free_blocks.emplace(sz, p);
  g_cached_bytes += (int64_t)sz;
  return IIFE_OK;
}

// ------------------------------------------------------------------ exclusive scan
// Three-phase scan: per-block sums (int64) -> scan of block sums (recursive) -> per-block rescan.
// Each block of 256 threads covers SCAN_ITEMS*256 elements.
constexpr int SCAN_THREADS = 256;
constexpr int SCAN_ITEMS = 8;
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;

__global__ void scan_block_sums(const int *__restrict__ in, int64_t n, long long *__restrict__ bsum) {
  __shared__ long long sh[SCAN_THREADS / 32];
  int64_t base = (int64_t)blockIdx.x * SCAN_TILE;
  long long s = 0;
#pragma unroll
  for (int k = 0; k < SCAN_ITEMS; ++k) {
    int64_t i = base + (int64_t)k * SCAN_THREADS + threadIdx.x;
    if (i < n) s += in[i];
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    long long t = 0;
    for (int w = 0; w < SCAN_THREADS / 32; ++w) t += sh[w];
    bsum[blockIdx.x] = t;
  }
}

// single-block exclusive scan of up to a few thousand int64 block sums, in place; total -> bsum[nb]
__global__ void scan_small_i64(long long *__restrict__ a, int nb) {
  __shared__ long long carry;
  __shared__ long long sh[1024];
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (int base = 0; base < nb; base += 1024) {
    int i = base + threadIdx.x;
    long long v = i < nb ? a[i] : 0;
    sh[threadIdx.x] = v;
    __syncthreads();
    for (int o = 1; o < 1024; o <<= 1) {
      long long t = threadIdx.x >= o ? sh[threadIdx.x - o] : 0;
      __syncthreads();
      sh[threadIdx.x] += t;
      __syncthreads();
    }
    long long incl = sh[threadIdx.x];
    if (i < nb) a[i] = carry + incl - v;
    __syncthreads();
    if (threadIdx.x == 1023) carry += incl;
    __syncthreads();
  }
  if (threadIdx.x == 0) a[nb] = carry;
}

template <class OutT>
__global__ void scan_apply(const int *__restrict__ in, OutT *__restrict__ out, int64_t n,
                           const long long *__restrict__ boff) {
  // thread t owns SCAN_ITEMS consecutive elements of the tile -> serial scan + warp/block scan of sums
  __shared__ long long wsum[SCAN_THREADS / 32];
  int64_t base = (int64_t)blockIdx.x * SCAN_TILE + (int64_t)threadIdx.x * SCAN_ITEMS;
  int v[SCAN_ITEMS];
  long long s = 0;
#pragma unroll
  for (int k = 0; k < SCAN_ITEMS; ++k) {
    int64_t i = base + k;
    v[k] = i < n ? in[i] : 0;
    s += v[k];
  }
  // inclusive warp scan of s
  long long incl = s;
  int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    long long t = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += t;
  }
  if (lane == 31) wsum[w] = incl;
  __syncthreads();
  long long woff = 0;
  for (int k = 0; k < w; ++k) woff += wsum[k];
  long long run = boff[blockIdx.x] + woff + incl - s;
#pragma unroll
  for (int k = 0; k < SCAN_ITEMS; ++k) {
    int64_t i = base + k;
    if (i < n) out[i] = (OutT)run;
    run += v[k];
  }
  if (blockIdx.x == gridDim.x - 1 && threadIdx.x == SCAN_THREADS - 1) {
    // total goes to out[n]
    out[n] = (OutT)run;
  }
}

__global__ void scan_empty(int *out) { out[0] = 0; }

int exclusive_scan_i32(const int *in, int *out, int64_t n, int64_t *total64) {
  Ctx &c = ctx();
  if (n == 0) {
    IIFE_LAUNCH(scan_empty, 1, 1, 0, out);
    IIFE_CHECK_LAUNCH();
    if (total64) *total64 = 0;
    return IIFE_OK;
  }
  int64_t nb64 = (n + SCAN_TILE - 1) / SCAN_TILE;
  if (nb64 > 0x7fffffff) return set_err(IIFE_ERR_UNSUPPORTED, "scan of %lld elements too large", (long long)n);
  int nb = (int)nb64;
  Tmp<long long> bsum;
  IIFE_TRY(bsum.alloc((size_t)nb + 1));
  IIFE_LAUNCH(scan_block_sums, nb, SCAN_THREADS, 0, in, n, bsum.p);
  IIFE_LAUNCH(scan_small_i64, 1, 1024, 0, bsum.p, nb);
  IIFE_LAUNCH(scan_apply<int>, nb, SCAN_THREADS, 0, in, out, n, bsum.p);
  IIFE_CHECK_LAUNCH();
  long long total = 0;
  IIFE_CUDA(cudaMemcpyAsync(&total, bsum.p + nb, sizeof(long long), cudaMemcpyDeviceToHost, c.stream));
  IIFE_CUDA(cudaStreamSynchronize(c.stream));
  if (total64) *total64 = total;
  if (total >= 0x7fffffffLL)
    return set_err(IIFE_ERR_UNSUPPORTED, "scan total %lld does not fit int32 indices", total);
  return IIFE_OK;
}

__global__ void scan_empty64(long long *out) { out[0] = 0; }

// same scan with 64-bit offsets (slot-plan offsets exceed 2^31 at the 50 M-DOF size)
int exclusive_scan_i32_i64(const int *in, long long *out, int64_t n, int64_t *total64) {
  Ctx &c = ctx();
  if (n == 0) {
    IIFE_LAUNCH(scan_empty64, 1, 1, 0, out);
    IIFE_CHECK_LAUNCH();
    if (total64) *total64 = 0;
    return IIFE_OK;
  }
  int64_t nb64 = (n + SCAN_TILE - 1) / SCAN_TILE;
  if (nb64 > 0x7fffffff) return set_err(IIFE_ERR_UNSUPPORTED, "scan of %lld elements too large", (long long)n);
  int nb = (int)nb64;
  Tmp<long long> bsum;
  IIFE_TRY(bsum.alloc((size_t)nb + 1));
  IIFE_LAUNCH(scan_block_sums, nb, SCAN_THREADS, 0, in, n, bsum.p);
  IIFE_LAUNCH(scan_small_i64, 1, 1024, 0, bsum.p, nb);
  IIFE_LAUNCH(scan_apply<long long>, nb, SCAN_THREADS, 0, in, out, n, bsum.p);
  IIFE_CHECK_LAUNCH();
  long long total = 0;
  IIFE_CUDA(cudaMemcpyAsync(&total, bsum.p + nb, sizeof(long long), cudaMemcpyDeviceToHost, c.stream));
  IIFE_CUDA(cudaStreamSynchronize(c.stream));
  if (total64) *total64 = total;
  return IIFE_OK;
}

}  // namespace iife

using namespace iife;

extern "C" {

int iife_version(void) { return IIFE_VERSION; }
const char *iife_last_error(void) { return get_err(); }

int iife_device_count(int *n) {
  if (!n) return set_err(IIFE_ERR_ARG, "n is NULL");
  int c = 0;
  cudaError_t e = cudaGetDeviceCount(&c);
  if (e != cudaSuccess) {
    cudaGetLastError();
    *n = 0;
    return set_err(IIFE_ERR_NO_DEVICE, "cudaGetDeviceCount: %s", cudaGetErrorString(e));
  }
  *n = c;
  return IIFE_OK;
}

int iife_init(int device) {
  Ctx &c = ctx();
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n == 0) {
    cudaGetLastError();
    return set_err(IIFE_ERR_NO_DEVICE, "no CUDA device available (%s); libiife has no CPU fallback",
                   e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
  }
  if (device < 0 || device >= n) return set_err(IIFE_ERR_ARG, "device %d out of range [0,%d)", device, n);
  if (c.init && c.device == device) return IIFE_OK;
  if (c.init) return set_err(IIFE_ERR_STATE, "already initialised on device %d", c.device);
  IIFE_CUDA(cudaSetDevice(device));
  cudaDeviceProp prop;
  IIFE_CUDA(cudaGetDeviceProperties(&prop, device));
  if (prop.major < 10)
    return set_err(IIFE_ERR_NO_DEVICE, "device %d is sm_%d%d; libiife is built for sm_100a only", device, prop.major,
                   prop.minor);
  c.sm_count = prop.multiProcessorCount;
  c.max_smem_optin = (int)prop.sharedMemPerBlockOptin;
  IIFE_CUDA(cudaStreamCreateWithFlags(&c.own_stream, cudaStreamNonBlocking));
  c.stream = c.own_stream;
  c.device = device;
  c.init = true;
  return IIFE_OK;
}

int iife_plan_cache_clear(void);


int iife_finalize(void) {
  Ctx &c = ctx();
  if (!c.init) return IIFE_OK;
  iife_plan_cache_clear();
  cudaStreamSynchronize(c.stream);
  ksp_release_cached_graphs();
  dev_release_cached();
  if (c.own_stream) cudaStreamDestroy(c.own_stream);
  c.own_stream = nullptr;
  c.stream = nullptr;
  c.init = false;
  c.device = -1;
  return IIFE_OK;
}

int iife_set_stream(void *s) {
  IIFE_NEED_INIT();
  Ctx &c = ctx();
  IIFE_CUDA(cudaStreamSynchronize(c.stream));
  c.stream = s ? (cudaStream_t)s : c.own_stream;
  return IIFE_OK;
}

int iife_sync(void) {
  IIFE_NEED_INIT();
  IIFE_CUDA(cudaStreamSynchronize(ctx().stream));
  return IIFE_OK;
}

int iife_device_bytes(int64_t *bytes) {
  if (!bytes) return set_err(IIFE_ERR_ARG, "bytes is NULL");
  *bytes = ctx().dev_bytes;
  return IIFE_OK;
}

int iife_launch_count(int64_t *n, int reset) {
  if (n) *n = ctx().launches;
  if (reset) ctx().launches = 0;
  return IIFE_OK;
}

}  // extern "C"
