// Graph handle management: adopt / export CSR Laplacians, row-block partition.
#include "common.cuh"

#include <atomic>
#include <thread>

#include <string.h>
#include <stdlib.h>

namespace meld {

static thread_local char g_err[1024] = "";
long long g_launches = 0;
long long g_syncs = 0;

void set_error(const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

static thread_local cudaStream_t g_stream = nullptr;
cudaStream_t &current_stream() { return g_stream; }

bool use_pool() {
  static int v = -1;
  if (v < 0) v = getenv("MELD_B200_NO_POOL") ? 0 : 1;
  return v == 1;
}

void use_stream(cudaStream_t s) {
  g_stream = s;
  static thread_local int configured_dev = -1;
  int dev = -1;
  if (cudaGetDevice(&dev) == cudaSuccess && dev != configured_dev) {
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
      unsigned long long keep = ~0ull;  // never hand memory back to the OS between calls
      cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
    }
    configured_dev = dev;
  }
  cudaGetLastError();
}

// ---- build-scoped arena (see common.cuh) ------------------------------------------------------------------
namespace {
struct Arena {
  char *base = nullptr;
  size_t cap = 0, off = 0, need = 0, high = 0;
  std::atomic<bool> active{false};  // claimed with a compare-exchange: ctypes drops the GIL around every call
  int dev = -1;
  std::thread::id owner;  // only the thread whose build opened the arena allocates from it
} g_arena;
constexpr size_t kArenaAlign = 256;
}  // namespace

bool arena_begin(cudaStream_t s) {
  Arena &a = g_arena;
  if (getenv("MELD_B200_NO_ARENA")) return false;
  bool expected = false;
  if (!a.active.compare_exchange_strong(expected, true)) return false;  // nested / concurrent build: stays on the pool
  a.owner = std::this_thread::get_id();
  int dev = -1;
  cudaGetDevice(&dev);
  if (a.base && dev != a.dev) {  // the process switched devices: start over
    cudaSetDevice(a.dev);
    cudaFree(a.base);
    cudaSetDevice(dev);
    a.base = nullptr;
    a.cap = a.high = 0;
  }
  a.dev = dev;
  if (a.high > a.cap) {  // grow to the high-water mark of earlier builds (+6 %), outside any timed kernel work
    cudaStreamSynchronize(s);
    if (a.base) cudaFree(a.base);
    a.base = nullptr;
    const size_t want = a.high + a.high / 16 + (64u << 20);
    if (cudaMalloc((void **)&a.base, want) == cudaSuccess) {
      a.cap = want;
    } else {
      cudaGetLastError();
      a.cap = 0;
    }
  }
  a.off = a.need = 0;
  return true;
}

void arena_end() {
  Arena &a = g_arena;
  if (!a.active.load() || a.owner != std::this_thread::get_id()) return;  // only the build that claimed it
  if (a.need > a.high) a.high = a.need;
  a.off = 0;
  a.owner = std::thread::id();  // no thread matches between two builds
  a.active.store(false);
}

void *arena_alloc(size_t bytes) {
  Arena &a = g_arena;
  if (!a.active.load() || a.owner != std::this_thread::get_id()) return nullptr;
  const size_t sz = (bytes + kArenaAlign - 1) / kArenaAlign * kArenaAlign;
  a.need += sz;
  if (!a.base || a.off + sz > a.cap) return nullptr;
  void *p = a.base + a.off;
  a.off += sz;
  return p;
}

bool arena_owns(const void *p) {
  const Arena &a = g_arena;
  return a.base && (const char *)p >= a.base && (const char *)p < a.base + a.cap;
}

void arena_release() {
  Arena &a = g_arena;
  bool expected = false;
  if (!a.active.compare_exchange_strong(expected, true)) return;  // a build is running: nothing to free now
  if (a.base) {
    int dev = -1;
    cudaGetDevice(&dev);
    if (a.dev >= 0 && dev != a.dev) cudaSetDevice(a.dev);
    cudaDeviceSynchronize();
    cudaFree(a.base);
    if (a.dev >= 0 && dev != a.dev) cudaSetDevice(dev);
  }
  a.base = nullptr;
  a.cap = a.off = 0;  // `high` stays: the next build re-grows the arena in one step
  a.active.store(false);
}

int sm_count() {
  static int cached = 0;
  if (cached) return cached;
  int dev = 0, n = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 0;
  if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return 0;
  cached = n;
  return n;
}

// ---- tuning knobs (bench / tests only; defaults are the shipped configuration) ------
Tuning g_tuning;

Tuning &tuning() { return g_tuning; }

// blk[b] = first row whose first entry is at or after b * chunk.
__global__ void partition_rows_kernel(const int32_t *__restrict__ row_ptr, int64_t n_rows, int32_t chunk,
                                      int32_t n_blk, int32_t *__restrict__ blk) {
  int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b > n_blk) return;
  if (b == n_blk) {
    blk[b] = (int32_t)n_rows;
    return;
  }
  int64_t target = (int64_t)b * chunk;
  int64_t lo = 0, hi = n_rows;  // lower_bound over row_ptr[0..n_rows)
  while (lo < hi) {
    int64_t mid = (lo + hi) >> 1;
    if ((int64_t)row_ptr[mid] < target)
      lo = mid + 1;
    else
      hi = mid;
  }
  blk[b] = (int32_t)lo;
}

int graph_finalize(meld_b200_graph *g, cudaStream_t stream) {
  const int32_t chunk = tuning().blk_chunk;
  MELD_REQUIRE(chunk >= 32, "graph_finalize: bad tuning (blk_chunk=%d)", chunk);
  g->blk_chunk = chunk;
  g->n_blk = (int32_t)(g->nnz > 0 ? ceil_div(g->nnz, chunk) : 1);
  MELD_CHECK(g->blk.alloc((size_t)g->n_blk + 1));
  partition_rows_kernel<<<(unsigned)ceil_div(g->n_blk + 1, 256), 256, 0, stream>>>(g->row_ptr.p, g->n_rows, chunk,
                                                                                  g->n_blk, g->blk.p);
  MELD_LAUNCH_CHECK();
  MELD_SYNC(stream);  // builders return synchronised: the caller may free what it passed in
  g->stats[5] = 0;
  g->stats[6] = 0;
  g->stats[7] = g->n_blk;
  return 0;
}

__global__ void indptr64_to_32_kernel(const int64_t *__restrict__ in, int64_t n, int32_t *__restrict__ out) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = (int32_t)in[i];
}
__global__ void indptr32_to_64_kernel(const int32_t *__restrict__ in, int64_t n, int64_t *__restrict__ out) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = (int64_t)in[i];
}

}  // namespace meld

using namespace meld;

extern "C" {

int meld_b200_version(void) { return 200; /* 0.2.0 */ }

int meld_b200_release_workspace(void) {
  meld::arena_release();
  return 0;
}

int64_t meld_b200_launch_count(void) { return (int64_t)g_launches; }

int64_t meld_b200_sync_count(void) { return (int64_t)g_syncs; }

const char *meld_b200_last_error(void) { return g_err; }

int meld_b200_device_info(int *sm_count_host, int *cc_major_host, int *cc_minor_host) {
  int dev = 0;
  MELD_CUDA(cudaGetDevice(&dev));
  int sms = 0, maj = 0, min = 0;
  MELD_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  MELD_CUDA(cudaDeviceGetAttribute(&maj, cudaDevAttrComputeCapabilityMajor, dev));
  MELD_CUDA(cudaDeviceGetAttribute(&min, cudaDevAttrComputeCapabilityMinor, dev));
  if (sm_count_host) *sm_count_host = sms;
  if (cc_major_host) *cc_major_host = maj;
  if (cc_minor_host) *cc_minor_host = min;
  return 0;
}

int meld_b200_set_tuning(const char *key, int value) {
  MELD_REQUIRE(key != nullptr, "set_tuning: NULL key");
  Tuning &t = tuning();
  if (!strcmp(key, "blk_chunk")) t.blk_chunk = value;
  else if (!strcmp(key, "p1_segments")) t.p1_segments = value;
  else if (!strcmp(key, "reorder")) t.reorder = value;
  else if (!strcmp(key, "ctas_per_sm")) t.ctas_per_sm = value;
  else if (!strcmp(key, "group")) t.group = value;
  else if (!strcmp(key, "tc_multicast")) t.tc_multicast = value;
  else if (!strcmp(key, "prune")) t.prune = value;
  else if (!strcmp(key, "clusters")) t.clusters = value;
  else if (!strcmp(key, "kmeans_iters")) t.kmeans_iters = value;
  else if (!strcmp(key, "km_var_pct")) t.km_var_pct = value;
  else if (!strcmp(key, "reorder_min_n")) t.reorder_min_n = value;
  else if (!strcmp(key, "prune_window")) t.prune_window = value;
  else if (!strcmp(key, "cluster_cells")) t.cluster_cells = value;
  else if (!strcmp(key, "reg_topk")) t.reg_topk = value;
  else if (!strcmp(key, "prune_proj")) t.prune_proj = value;
  else if (!strcmp(key, "merge_rows")) t.merge_rows = value;
  else if (!strcmp(key, "tl_sort")) t.tl_sort = value;
  else if (!strcmp(key, "tl_chunks")) t.tl_chunks = value;
  else if (!strcmp(key, "flat_threads")) t.flat_threads = value;
  else if (!strcmp(key, "flat_group")) t.flat_group = value;
  else if (!strcmp(key, "flat_sched")) t.flat_sched = value;
  else if (!strcmp(key, "flat_pipe")) t.flat_pipe = value;
  else if (!strcmp(key, "flat_gen")) t.flat_gen = value;
  else if (!strcmp(key, "pad_width")) t.pad_width = value;
  else if (!strcmp(key, "flat_hint")) t.flat_hint = value;
  else if (!strcmp(key, "flat_layout")) t.flat_layout = value;
  else if (!strcmp(key, "tl_interleave")) t.tl_interleave = value;
  else {
    set_error("set_tuning: unknown key '%s'", key);
    return MELD_B200_ERR_INVALID;
  }
  return 0;
}

int meld_b200_graph_from_csr(int64_t n_rows, int64_t n_cols, int64_t row0, int64_t nnz, const int64_t *indptr,
                             const int32_t *indices, const double *data, void *stream_, meld_b200_graph_t **out) {
  cudaStream_t stream = (cudaStream_t)stream_;
  meld::use_stream(stream);
  MELD_REQUIRE(out != nullptr, "graph_from_csr: graph_out is NULL");
  *out = nullptr;
  MELD_REQUIRE(n_rows >= 0 && n_cols > 0 && row0 >= 0 && row0 + n_rows <= n_cols,
               "graph_from_csr: bad shape n_rows=%lld n_cols=%lld row0=%lld", (long long)n_rows, (long long)n_cols,
               (long long)row0);
  MELD_REQUIRE(nnz >= 0 && nnz < (int64_t)2147483647 - kCsrPad, "graph_from_csr: nnz=%lld out of range",
               (long long)nnz);
  MELD_REQUIRE(n_cols < (int64_t)2147483647, "graph_from_csr: n_cols too large for int32 columns");
  MELD_REQUIRE(indptr && (nnz == 0 || (indices && data)), "graph_from_csr: NULL array");
  meld_b200_graph *g = new (std::nothrow) meld_b200_graph();
  if (!g) {
    set_error("graph_from_csr: host allocation failed");
    return MELD_B200_ERR_NOMEM;
  }
  g->n_rows = n_rows;
  g->n_cols = n_cols;
  g->row0 = row0;
  g->nnz = nnz;
  int rc = 0;
  auto fail = [&](int code) {
    delete g;
    return code;
  };
  if ((rc = g->row_ptr.alloc((size_t)n_rows + 1 + kCsrPad))) return fail(rc);
  if ((rc = g->col.alloc((size_t)nnz + kCsrPad))) return fail(rc);
  if ((rc = g->val.alloc((size_t)nnz + kCsrPad))) return fail(rc);
  indptr64_to_32_kernel<<<(unsigned)ceil_div(n_rows + 1, 256), 256, 0, stream>>>(indptr, n_rows + 1, g->row_ptr.p);
  cudaError_t e = cudaGetLastError();
  if (e == cudaSuccess) e = cudaMemsetAsync(g->row_ptr.p + n_rows + 1, 0, kCsrPad * sizeof(int32_t), stream);
  if (e == cudaSuccess) e = cudaMemsetAsync(g->col.p + nnz, 0, kCsrPad * sizeof(int32_t), stream);
  if (e == cudaSuccess) e = cudaMemsetAsync(g->val.p + nnz, 0, kCsrPad * sizeof(double), stream);
  if (e == cudaSuccess && nnz > 0)
    e = cudaMemcpyAsync(g->col.p, indices, (size_t)nnz * sizeof(int32_t), cudaMemcpyDeviceToDevice, stream);
  if (e == cudaSuccess && nnz > 0)
    e = cudaMemcpyAsync(g->val.p, data, (size_t)nnz * sizeof(double), cudaMemcpyDeviceToDevice, stream);
  if (e != cudaSuccess) {
    set_error("graph_from_csr: %s", cudaGetErrorString(e));
    return fail(MELD_B200_ERR_CUDA);
  }
  if ((rc = graph_finalize(g, stream))) return fail(rc);
  *out = g;
  return 0;
}

int meld_b200_graph_info(const meld_b200_graph_t *g, int64_t *n_rows, int64_t *n_cols, int64_t *row0, int64_t *nnz) {
  MELD_REQUIRE(g != nullptr, "graph_info: NULL graph");
  if (n_rows) *n_rows = g->n_rows;
  if (n_cols) *n_cols = g->n_cols;
  if (row0) *row0 = g->row0;
  if (nnz) *nnz = g->nnz;
  return 0;
}

int meld_b200_graph_export_csr(const meld_b200_graph_t *g, int64_t *indptr, int32_t *indices, double *data,
                               void *stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  meld::use_stream(stream);
  MELD_REQUIRE(g && indptr && indices && data, "graph_export_csr: NULL argument");
  indptr32_to_64_kernel<<<(unsigned)ceil_div(g->n_rows + 1, 256), 256, 0, stream>>>(g->row_ptr.p, g->n_rows + 1,
                                                                                  indptr);
  MELD_LAUNCH_CHECK();
  if (g->nnz > 0) {
    MELD_CUDA(cudaMemcpyAsync(indices, g->col.p, (size_t)g->nnz * sizeof(int32_t), cudaMemcpyDeviceToDevice, stream));
    MELD_CUDA(cudaMemcpyAsync(data, g->val.p, (size_t)g->nnz * sizeof(double), cudaMemcpyDeviceToDevice, stream));
  }
  return 0;
}

int meld_b200_graph_build_stats(const meld_b200_graph_t *g, int64_t *stats8_host) {
  MELD_REQUIRE(g && stats8_host, "graph_build_stats: NULL argument");
  memcpy(stats8_host, g->stats, sizeof(g->stats));
  return 0;
}

int meld_b200_graph_destroy(meld_b200_graph_t *g) {
  // buffers are returned to the pool in stream order on the stream of this thread's last call; make sure
  // nothing that may still run on another stream reads them
  if (g) {
    cudaDeviceSynchronize();
    meld::use_stream(nullptr);  // free on the legacy default stream
  }
  delete g;
  return 0;
}

}  // extern "C"
