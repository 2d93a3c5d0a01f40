// Shared internals of libmeld_b200.so (not part of the C-ABI).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>
#include <new>

#include "../../include/meld_b200.h"

namespace meld {

// ---- error plumbing --------------------------------------------------------------
void set_error(const char *fmt, ...);

#define MELD_CUDA(call)                                                                          \
  do {                                                                                           \
    cudaError_t err__ = (call);                                                                  \
    if (err__ != cudaSuccess) {                                                                  \
      meld::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(err__)); \
      return (err__ == cudaErrorMemoryAllocation) ? MELD_B200_ERR_NOMEM : MELD_B200_ERR_CUDA;    \
    }                                                                                            \
  } while (0)

#define MELD_CHECK(expr)        \
  do {                          \
    int rc__ = (expr);          \
    if (rc__ != 0) return rc__; \
  } while (0)

#define MELD_REQUIRE(cond, ...)        \
  do {                                 \
    if (!(cond)) {                     \
      meld::set_error(__VA_ARGS__);    \
      return MELD_B200_ERR_INVALID;    \
    }                                  \
  } while (0)

// Every kernel launch of this library goes through MELD_LAUNCH_CHECK, which also counts it
// (meld_b200_launch_count; bench.py reports the count as "gpu_launches").
extern long long g_launches;
#define MELD_LAUNCH_CHECK()          \
  do {                               \
    ++meld::g_launches;              \
    MELD_CUDA(cudaGetLastError());   \
  } while (0)

// Host waits on the stream are counted too (meld_b200_sync_count; bench.py reports them per step).
extern long long g_syncs;
#define MELD_SYNC(s)                          \
  do {                                        \
    ++meld::g_syncs;                          \
    MELD_CUDA(cudaStreamSynchronize(s));      \
  } while (0)

static inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }
static inline int64_t round_up(int64_t a, int64_t b) { return ceil_div(a, b) * b; }

int sm_count();

// Launch configuration of the Chebyshev SpMM kernels (see cheby.cu) and of the graph build (knn.cu, knn_tc.cu).  The
// defaults are the shipped configuration; meld_b200_set_tuning changes them for bench sweeps and tests.
struct Tuning {
  int blk_chunk = 768;    // target CSR entries per row block (the nonzero-balanced row ranges of flat_sched = 1)
  int ctas_per_sm = 1;    // persistent CTAs per SM (round-1 flat kernels)
  int group = 0;          // lanes per row (0 = choose from mean nnz/row)
  int flat_threads = 1024;  // threads per CTA of the flat kernel (1024: <= 64 registers, 768: <= 80)
  int flat_pipe = 2;      // software-pipelined flat kernel (8 lanes per row): 0 never, 1 for P <= 4, 2 for P = 1 only
                          // (measured: 154 us either way at p = 4 -- L1-tag bound; Lanczos p = 1 gains 15 %)
  int flat_sched = 1;     // flat kernel: 1 = every CTA owns a contiguous, nonzero-balanced row range; 0 = round-robin
  int flat_group = 0;     // lanes per row of the flat kernel (0 = `group`)
  int pad_width = 0;      // 1: signals of 3 / 5..7 columns run zero-padded to 4 / 8 (measured slower, see cheby.cu)
  int flat_gen = 1;       // 1: second-generation flat kernel (cheby_flat2_kernel), 0: the round-1 flat kernels
  int flat_hint = -1;     // cheby_flat2_kernel HINT (0..3, see cheby.cu); -1: chosen from the signal width
  int flat_layout = -1;   // cheby_flat2_kernel LAYOUT (0: 4 consecutive entries per lane, 1: lane-consecutive); -1: auto
  int reorder = 1;        // knn_graph_build orders cells along a Morton curve of the leading dims
  int tc_multicast = 2;   // candidate search: CTA cluster size (1, 2, 4) sharing B tiles by TMA multicast
  int p1_segments = 0;    // candidate search pass 1 scans this many column segments per row (own first; 0 = all)
  int prune = 1;          // candidate search skips (row tile, column tile) pairs that bounding balls prove too far apart
  int clusters = 64;      // k-means clusters of the internal cell order (0: Morton order only)
  int kmeans_iters = 1;   // Lloyd iterations after seeding
  int km_var_pct = 99;    // k-means / projection features: leading ones up to this % of the total variance (0: up to 128)
  int reorder_min_n = 4096;  // cells are re-ordered (clusters + Morton curve) from this size on
  int merge_rows = 1;     // graph assembly places mirrored entries by rank instead of a segmented sort of every row
  int prune_proj = 1;     // tile pruning also uses the projection bound between k-means clusters
  int reg_topk = 1;       // pass 1 keeps its top-k lists in registers (k1 <= 32) instead of shared memory
  int tl_chunks = 2;      // most units per row tile in a list-driven search pass
  int tl_sort = 0;        // 1: pass-1 tile lists are ordered closest tile first (measured: no gain once insertion is cheap)
  int tl_interleave = 0;  // bit 0 / bit 1: pass 1 / pass 2 chunks of a tile list interleave instead of being contiguous
  int cluster_cells = 1024;  // fewest cells per k-means cluster (fewer clusters for small inputs)
  int prune_window = 2;   // the window pass of the pruned search scans own tile +- this many column tiles
};
Tuning &tuning();

// ---- device buffer with explicit ownership -----------------------------------------
// Allocations are stream-ordered (cudaMallocAsync from the device's default pool, which is told to keep
// its memory): a build allocates ~20 temporaries of up to a GB, and plain cudaMalloc / cudaFree cost
// milliseconds each -- tens of milliseconds once NCCL has enabled peer access, because every allocation
// is then mapped into every peer.  The stream is the one of the API call in progress on this thread.
cudaStream_t &current_stream();
void use_stream(cudaStream_t s);  // also configures the pool on first use
bool use_pool();                  // MELD_B200_NO_POOL=1 falls back to cudaMalloc / cudaFree (diagnosis)

// Build-scoped temporaries come from ONE library-owned arena instead of the pool: a build allocates ~60 device
// buffers (4-5 GB at 500k cells), and although the pool keeps its memory, cudaMallocAsync / cudaFreeAsync of
// such blocks were seen to stall the calling thread by 30-700 ms at random on busy hosts -- an order of magnitude
// more than the ~45 ms the build's kernels take.  arena_begin() at the top of a build makes DevBuf::alloc a
// pointer bump; arena_end() resets it (a build ends synchronised, so nothing is still in use).  The first build
// of a process runs on the pool and records the high-water mark; the arena is sized from it once.  Buffers
// that outlive the build (everything inside a graph / candidates handle) are `persistent` and stay on the pool.
void *arena_alloc(size_t bytes);   // nullptr: arena not active or full (caller falls back to the pool)
bool arena_owns(const void *p);
bool arena_begin(cudaStream_t s);  // false: another build holds the arena (this one stays on the pool)
void arena_end();
void arena_release();              // frees the block (meld_b200_release_workspace)
struct ArenaScope {
  bool mine;
  explicit ArenaScope(cudaStream_t s) : mine(arena_begin(s)) {}
  ~ArenaScope() {
    if (mine) arena_end();
  }
  ArenaScope(const ArenaScope &) = delete;
  ArenaScope &operator=(const ArenaScope &) = delete;
};

template <typename T>
struct DevBuf {
  T *p = nullptr;
  size_t n = 0;
  bool persistent = false;
  cudaStream_t s = nullptr;  // stream the block was allocated on; it is freed on the same one
  int alloc(size_t count) {
    release();
    if (count == 0) count = 1;
    if (!persistent) {
      p = static_cast<T *>(arena_alloc(count * sizeof(T)));
      if (p) {
        n = count;
        return 0;
      }
    }
    s = current_stream();
    cudaError_t e = use_pool() ? cudaMallocAsync((void **)&p, count * sizeof(T), s)
                               : cudaMalloc((void **)&p, count * sizeof(T));
    if (e != cudaSuccess) {
      p = nullptr;
      set_error("cudaMallocAsync(%zu bytes) failed: %s", count * sizeof(T), cudaGetErrorString(e));
      cudaGetLastError();
      return MELD_B200_ERR_NOMEM;
    }
    n = count;
    return 0;
  }
  void release() {
    if (p && !arena_owns(p)) {
      if (use_pool())
        cudaFreeAsync(p, s);  // ordered after the kernels queued on the stream the block lives on, whichever
                              // thread or stream the caller is on by now (graph_destroy also synchronises)
      else
        cudaFree(p);
    }
    p = nullptr;
    n = 0;
  }
  ~DevBuf() { release(); }
  DevBuf() = default;
  explicit DevBuf(bool persistent_) : persistent(persistent_) {}
  DevBuf(const DevBuf &) = delete;
  DevBuf &operator=(const DevBuf &) = delete;
};

// Elements of padding appended to col/val so the kernels' 16- / 32-byte vector loads of a row's entries may start
// before / end after the row's own entries.
constexpr int kCsrPad = 16;

}  // namespace meld

// The opaque handle of include/meld_b200.h.
struct meld_b200_graph {
  int64_t n_rows = 0, n_cols = 0, row0 = 0, nnz = 0;
  // CSR of L (values f64, columns int32, row pointers int32: nnz < 2^31 per GPU).
  meld::DevBuf<int32_t> row_ptr{true};  // n_rows + 1
  meld::DevBuf<int32_t> col{true};      // nnz + kCsrPad
  meld::DevBuf<double> val{true};       // nnz + kCsrPad
  // Row-block partition (blocks of ~blk_chunk entries): block b = rows [blk[b], blk[b+1]); the SpMM kernels cut their
  // per-CTA row ranges at block boundaries so every CTA streams the same number of nonzeros.
  meld::DevBuf<int32_t> blk{true};  // n_blk + 1
  int32_t n_blk = 0;
  int32_t blk_chunk = 0;  // target nnz per block (C)
  // Cell order used internally (graph row a = caller's cell perm[a]); null = identity.
  meld::DevBuf<int32_t> perm{true};
  // Row slices of a partitioned operator: bit k of halo[i] = the k-th peer (ranks in order, this one skipped)
  // references local row i as a column of its own rows, i.e. needs this row of every exchanged vector.
  meld::DevBuf<uint8_t> halo{true};
  // Chebyshev / Lanczos workspace, grown on demand.
  meld::DevBuf<double> work{true};
  // Un-symmetrised kNN kernel kept for export (compact CSR, slot order), optional.
  meld::DevBuf<int64_t> knn_ptr{true};   // n + 1
  meld::DevBuf<int32_t> knn_cnt{true};   // n + 1
  meld::DevBuf<int32_t> knn_col{true};   // knn_nnz
  meld::DevBuf<double> knn_val{true};    // knn_nnz
  int64_t knn_nnz = -1;
  int64_t stats[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  // ms of search pass 1 / pass 2, flops issued by pass 2 / by pass 1, flops of an unpruned pass, reserved
  double times[8] = {0, 0, 0, 0, 0, 0, 0, 0};
};

// Peer-memory context of a row-partitioned filter (include/meld_b200.h: meld_b200_dist_t): one cudaMalloc'd
// block per rank -- [flags | error word | done counter | two full-length signal buffers] -- exported through
// CUDA IPC and mapped by every other rank of the box, so a rank's kernels store straight into its peers' HBM
// over NVLink and publish completion through the peers' flag words.
struct meld_b200_dist {
  int rank = 0, world = 1;
  int64_t n = 0;       // rows of the full operator
  int p_max = 0;       // widest signal the buffers hold
  size_t buf_len = 0;  // doubles per signal buffer (padded)
  size_t bytes = 0;
  char *base = nullptr;              // this rank's block
  char *peer_base[8] = {nullptr};    // mapped blocks of all ranks ([rank] = base)
  unsigned long long epoch = 0;      // last phase published (identical on every rank: same call sequence)
  bool connected = false;
  // [0, 64) flag words, one per rank | 512 error word | 576 done counter | [640, 896) 4 x 8 scalar slots (the
  // rank-ordered partial sums of the distributed Lanczos, ring indexed by epoch) | 1024.. the two signal buffers
  static constexpr size_t kFlagsOff = 0, kErrOff = 512, kCtrOff = 576, kScalOff = 640, kBufOff = 1024;
  double *scal(int r) const { return reinterpret_cast<double *>(peer_base[r] + kScalOff); }
  unsigned long long *flags(int r) const { return reinterpret_cast<unsigned long long *>(peer_base[r] + kFlagsOff); }
  int *err() const { return reinterpret_cast<int *>(base + kErrOff); }
  unsigned int *ctr() const { return reinterpret_cast<unsigned int *>(base + kCtrOff); }
  double *buf(int r, int which) const {
    return reinterpret_cast<double *>(peer_base[r] + kBufOff) + (size_t)which * buf_len;
  }
};

namespace meld {
// Build the row-block partition of a graph whose row_ptr is final.
int graph_finalize(meld_b200_graph *g, cudaStream_t stream);
}  // namespace meld
