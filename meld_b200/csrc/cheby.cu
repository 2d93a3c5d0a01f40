// Chebyshev three-term recurrence on the CSR Laplacian (the HBM-roofline kernel),
// the Lanczos lmax estimate that reuses it, and the tiny signal helpers.
//
// Reference behaviour restated (upstream PyGSP 0.5.1, reached from meld/filter.py:39-59):
//   cheby_op: T0 = S; T1 = (L S - a2 S)/a1; R = c0/2 T0 + c1 T1;
//             T_k = (2/a1)(L - a2 I) T_{k-1} - T_{k-2}; R += c_k T_k      (a1 = a2 = lmax/2)
//   estimate_lmax: 1.01 * largest eigenvalue of L.
//
// One launch per recurrence term; T_k is written over T_{k-2} (row i of T_{k-2} is only ever read by row i).
// Kernels: no shared-memory staging, no warp roles; 32 warps per SM walk their own rows, gather the neighbours' signal
// rows into registers with 256-bit loads and stream values / columns with non-allocating vector loads.
//   cheby_flat2_kernel   (default, 8 lanes per row)  variants by template: cache hints, entry layout, DOT (fused
//                        Lanczos dot product), PEER (row-partitioned multi-GPU term: waits for the peers' flags,
//                        stores T_k into the peers' buffers over NVLink, publishes its own flag).  Bound by the L1TEX
//                        data pipe: one slot per gathered 32-byte sector, whatever level serves it.
//   cheby_flat_kernel / cheby_flat_pipe_kernel  (round 1)  4 / 16 / 32 lanes per row for very short (<= 12 entries) or
//                        very long (> 256) rows; also the cross-checks of tests/test_gpu_filter.py.
// The first design of round 1 (TMA-staged row blocks with per-block column dictionaries and cp.async gather warps,
// 286 us per launch at 500k cells against 152) was removed in round 2: it lost at every graph shape measured.
// Also here: the Lanczos lmax estimate (single GPU and row-partitioned), the shared-basis sweep
// (meld_b200_cheby_sweep), the row-partitioned filter (meld_b200_cheby_filter_dist) and the small signal helpers.
#include "common.cuh"

#include <math.h>
#include <stdlib.h>
#include <vector>

namespace meld {

constexpr int kMaxPeers = 7;  // ranks of one NVSwitch box besides this one

struct StepArgs {
  const int32_t *row_ptr;
  const int32_t *col;
  const double *val;
  const int32_t *blk;
  int32_t n_blk;
  int64_t row0;
  const double *Tcur;
  const double *Told;
  double *Tnew;
  double *R;
  double alpha, shift, gamma, c, c_cur;
  int r_acc;
  // flat kernel extras
  double *dot_partials;        // DOT variant (p = 1): per-CTA partial sums of T_cur[row] * y[row]
  // row-partitioned multi-GPU step (PEER variant): T_new is also stored into every peer's copy of the buffer
  // (P2P stores over NVLink) and completion is published through flags in peer memory
  double *peer_tnew[kMaxPeers];         // peers' T_new slices (already offset to this rank's first row)
  unsigned long long *peer_flags[kMaxPeers];  // peers' flag slot for THIS rank
  unsigned long long *my_flags;         // this rank's flag array (one slot per rank), written by the peers
  unsigned int *done_ctr;               // CTAs of this launch that have finished their stores
  int *err_flag;                        // set when a flag wait times out
  int n_peers, world, rank;
  int store_r;                          // last term: T_new / peer_tnew receive the finished R rows instead of T_k
  int peer_store;                       // 0: T_new stays local (distributed Lanczos: only the scalar is published)
  const uint8_t *halo;                  // per local row: which peers reference it (nullptr: all of them)
  double *peer_scal[kMaxPeers];         // peers' scalar slots (4 x 8 doubles, ring indexed by epoch & 3)
  double *my_scal;                      // this rank's scalar slots
  const double *scal_partials;          // per-CTA partial sums whose total the last CTA publishes (or nullptr)
  int n_scal_partials;
  unsigned long long wait_epoch;        // > 0: wait until every peer's flag >= wait_epoch before gathering
  unsigned long long post_epoch;        // > 0: value published to the peers when this launch's stores are done
};

// Gather one P-wide signal row straight from global memory (direct path): one 256-bit request per
// 32-byte row -- every distinct 128-byte line touched by a warp-level load costs L1 wavefronts.
__device__ __forceinline__ void ldg256(const double *p, double &a, double &b, double &c, double &d) {
  asm volatile("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(a), "=d"(b), "=d"(c), "=d"(d) : "l"(p));
}
template <int P>
__device__ __forceinline__ void gather_row(const double *__restrict__ T, int32_t c, double (&x)[P]) {
  const double *t = T + (size_t)c * P;
  if constexpr (P % 4 == 0) {
#pragma unroll
    for (int k = 0; k < P; k += 4) ldg256(t + k, x[k], x[k + 1], x[k + 2], x[k + 3]);
  } else if constexpr (P % 2 == 0) {
#pragma unroll
    for (int k = 0; k < P; k += 2) {
      double2 v = __ldg(reinterpret_cast<const double2 *>(t + k));
      x[k] = v.x;
      x[k + 1] = v.y;
    }
  } else {
#pragma unroll
    for (int k = 0; k < P; ++k) x[k] = __ldg(t + k);
  }
}
__device__ __forceinline__ double ld_stream(const double *p) {  // read-once operand: do not keep in L1
  double v;
  asm volatile("ld.global.cs.f64 %0, [%1];" : "=d"(v) : "l"(p));
  return v;
}
__device__ __forceinline__ void st_stream(double *p, double v) {
  asm volatile("st.global.cs.f64 [%0], %1;" ::"l"(p), "d"(v) : "memory");
}


// block-wide sum (fixed order => deterministic); sh holds 33 doubles
__device__ __forceinline__ double block_sum_fwd(double v, double *sh) {
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  __syncthreads();
  if (l == 0) sh[w] = v;
  __syncthreads();
  double t = (threadIdx.x < (blockDim.x >> 5)) ? sh[threadIdx.x] : 0.0;
  if (w == 0) {
    for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
    if (l == 0) sh[32] = t;
  }
  __syncthreads();
  return sh[32];
}

// ---- the flat kernels ---------------------------------------------------------------------------------------
// Measured on B200 (tools/microbench/gather_bench.cu): random 32-byte rows out of an L2-resident table arrive
// at 0.55 / 0.70 / 0.79 rows per cycle per SM with 1024 / 2048 / 4096 LDG.256 in flight per SM, LDGSTS and
// 32-byte TMA bulk copies are far slower (0.2 and 0.16 at best) -- the gather is bound by memory-level
// parallelism, not by L1 wavefronts.  So this kernel spends nothing on staging: no shared memory, no
// producer / gather / compute roles, no barriers; every one of the (up to) 32 warps per SM walks its own rows
// (G lanes per row, U independent gathers per lane in flight) and streams values and columns straight from
// global memory with evict-first loads.  Row groups are dealt to warps round-robin, so the warps of the grid
// stream one contiguous window of the matrix at any time.
__device__ __forceinline__ int ld_stream_s32(const int32_t *p) {
  int v;
  asm volatile("ld.global.cs.s32 %0, [%1];" : "=r"(v) : "l"(p));
  return v;
}

__device__ __forceinline__ void ld_stream_v4(const double *p, double (&v)[4]) {
  asm volatile("ld.global.cs.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(v[0]), "=d"(v[1]), "=d"(v[2]), "=d"(v[3]) : "l"(p));
}
__device__ __forceinline__ void ld_stream_c4(const int32_t *p, int (&c)[4]) {
  asm volatile("ld.global.cs.v4.s32 {%0,%1,%2,%3}, [%4];" : "=r"(c[0]), "=r"(c[1]), "=r"(c[2]), "=r"(c[3]) : "l"(p));
}

template <int P>
struct FlatUnroll {
  static constexpr int value = P <= 4 ? 4 : 2;  // gathers in flight per lane
};

template <int P, int G, int TB>
__global__ void __launch_bounds__(TB, 1) cheby_flat_kernel(const StepArgs a, const int64_t n_rows) {
  constexpr int U = FlatUnroll<P>::value;
  constexpr int RPW = 32 / G;
  const int lane = threadIdx.x & 31;
  const int gid = lane / G, gl = lane % G;
  // Row groups go to warps round-robin over the whole grid (a.blk == nullptr), or every CTA owns one contiguous
  // row range holding 1/gridDim of the nonzeros (cut at the row blocks of graph_finalize) and deals its
  // groups to its own warps: balanced by nonzeros, and a CTA's L1 keeps seeing the same neighbourhood.
  int64_t rb0 = ((int64_t)blockIdx.x * (TB / 32) + (threadIdx.x >> 5)) * RPW, r_end = n_rows;
  int64_t rstep = (int64_t)gridDim.x * (TB / 32) * RPW;
  if (a.blk != nullptr) {
    const int64_t b0 = ((int64_t)blockIdx.x * a.n_blk) / gridDim.x, b1 = ((int64_t)(blockIdx.x + 1) * a.n_blk) / gridDim.x;
    rb0 = __ldg(a.blk + b0) + (int64_t)(threadIdx.x >> 5) * RPW;
    r_end = __ldg(a.blk + b1);
    rstep = (int64_t)(TB / 32) * RPW;
  }
  for (int64_t rb = rb0; rb < r_end; rb += rstep) {  // warp-uniform
    const int64_t r = rb + gid;
    const bool act = r < r_end;
    int eb = 0, ee = 0;
    if (act) {
      eb = __ldg(a.row_ptr + r);
      ee = __ldg(a.row_ptr + r + 1);
    }
    const bool epi = act && gl < P;
    const size_t li = (size_t)r * P + gl;
    double tc = 0.0, told = 0.0, rold = 0.0;
    if (epi) {  // requested first so their latency overlaps the row product
      tc = __ldg(a.Tcur + (size_t)(a.row0 + r) * P + gl);
      if (a.gamma != 0.0) told = ld_stream(a.Told + li);
      if (a.R != nullptr && a.r_acc) rold = ld_stream(a.R + li);
    }
    double acc[P];
#pragma unroll
    for (int k = 0; k < P; ++k) acc[k] = 0.0;
    // Each lane takes U consecutive entries per iteration, starting at a multiple of 4 entries: values and
    // columns arrive as one aligned 256-bit / 128-bit load per lane (entries before the row's first or past
    // its last are masked; the arrays are padded by kCsrPad entries).
    for (int a0 = (eb & ~3) + gl * 4; a0 < ee; a0 += 4 * G) {
      double v[4];
      int c[4];
      ld_stream_v4(a.val + a0, v);
      ld_stream_c4(a.col + a0, c);
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int eu = a0 + u;
        if (eu < eb || eu >= ee) {
          v[u] = 0.0;
          c[u] = -1;
        }
      }
      if constexpr (U >= 4) {
        double x[4][P];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          if (c[u] >= 0) {
            gather_row<P>(a.Tcur, c[u], x[u]);
          } else {
#pragma unroll
            for (int k = 0; k < P; ++k) x[u][k] = 0.0;
          }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
#pragma unroll
          for (int k = 0; k < P; ++k) acc[k] = fma(v[u], x[u][k], acc[k]);
        }
      } else {  // wide signals: two gathers in flight at a time
#pragma unroll
        for (int h = 0; h < 4; h += 2) {
          double x[2][P];
#pragma unroll
          for (int u = 0; u < 2; ++u) {
            if (c[h + u] >= 0) {
              gather_row<P>(a.Tcur, c[h + u], x[u]);
            } else {
#pragma unroll
              for (int k = 0; k < P; ++k) x[u][k] = 0.0;
            }
          }
#pragma unroll
          for (int u = 0; u < 2; ++u) {
#pragma unroll
            for (int k = 0; k < P; ++k) acc[k] = fma(v[h + u], x[u][k], acc[k]);
          }
        }
      }
    }
    __syncwarp();
#pragma unroll
    for (int o = G / 2; o > 0; o >>= 1) {
#pragma unroll
      for (int k = 0; k < P; ++k) acc[k] += __shfl_xor_sync(0xffffffffu, acc[k], o);
    }
    if (epi) {
      double y = acc[0];
#pragma unroll
      for (int k = 1; k < P; ++k)
        if (gl == k) y = acc[k];
      double tn = a.alpha * (y - a.shift * tc);
      if (a.gamma != 0.0) tn -= a.gamma * told;
      if (a.Tnew) a.Tnew[li] = tn;  // gathered by the next step: keep cacheable
      if (a.R) {
        double rv = a.c * tn + a.c_cur * tc;
        if (a.r_acc) rv += rold;
        st_stream(a.R + li, rv);
      }
    }
  }
}

// Software-pipelined flat kernel (P <= 4).  In cheby_flat_kernel a warp's work is a serial chain per row group:
// row pointers -> (values, columns) -> gathers -> next 4 entries ..., three dependent memory latencies during
// which that warp has no gathers in flight.  Here the row pointers of the NEXT group are requested at the top
// of the current one, and the columns of the next step (next 4 entries of the row, or the first 4 of the next
// group) are requested BEFORE the current step's gathers, so a warp goes from gathers to gathers; the values
// of a step are requested together with its gathers (only the FMAs need them).
template <int P, int G, int TB>
__global__ void __launch_bounds__(TB, 1) cheby_flat_pipe_kernel(const StepArgs a, const int64_t n_rows) {
  static_assert(P <= 4, "pipelined flat kernel: P <= 4");
  constexpr int RPW = 32 / G;
  const int lane = threadIdx.x & 31;
  const int gid = lane / G, gl = lane % G;
  int64_t rb0 = ((int64_t)blockIdx.x * (TB / 32) + (threadIdx.x >> 5)) * RPW, r_end = n_rows;
  int64_t rstep = (int64_t)gridDim.x * (TB / 32) * RPW;
  if (a.blk != nullptr) {
    const int64_t b0 = ((int64_t)blockIdx.x * a.n_blk) / gridDim.x, b1 = ((int64_t)(blockIdx.x + 1) * a.n_blk) / gridDim.x;
    rb0 = __ldg(a.blk + b0) + (int64_t)(threadIdx.x >> 5) * RPW;
    r_end = __ldg(a.blk + b1);
    rstep = (int64_t)(TB / 32) * RPW;
  }
  if (rb0 >= r_end) return;  // warp-uniform
  // prologue: row pointers and first columns of the first group
  int eb_n = 0, ee_n = 0;
  if (rb0 + gid < r_end) {
    eb_n = __ldg(a.row_ptr + rb0 + gid);
    ee_n = __ldg(a.row_ptr + rb0 + gid + 1);
  }
  int cn[4] = {-1, -1, -1, -1};
  {
    const int a0 = (eb_n & ~3) + gl * 4;
    if (a0 < ee_n) ld_stream_c4(a.col + a0, cn);
  }
  for (int64_t rb = rb0; rb < r_end; rb += rstep) {  // warp-uniform
    const int64_t r = rb + gid;
    const bool act = r < r_end;
    const int eb = eb_n, ee = ee_n;
    {  // row pointers of the next group
      const int64_t rn = r + rstep;
      eb_n = ee_n = 0;
      if (rn < r_end) {
        eb_n = __ldg(a.row_ptr + rn);
        ee_n = __ldg(a.row_ptr + rn + 1);
      }
    }
    const bool epi = act && gl < P;
    const size_t li = (size_t)r * P + gl;
    double tc = 0.0, told = 0.0, rold = 0.0;
    if (epi) {
      tc = __ldg(a.Tcur + (size_t)(a.row0 + r) * P + gl);
      if (a.gamma != 0.0) told = ld_stream(a.Told + li);
      if (a.R != nullptr && a.r_acc) rold = ld_stream(a.R + li);
    }
    double acc[P];
#pragma unroll
    for (int k = 0; k < P; ++k) acc[k] = 0.0;
    int a0 = (eb & ~3) + gl * 4;
    while (true) {  // warp-uniform trip count: the longest row of the group
      int c[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int eu = a0 + u;
        c[u] = (eu >= eb && eu < ee) ? cn[u] : -1;
      }
      // 1. the columns of the next step
      const int a1 = a0 + 4 * G;
      const bool more = a1 < ee;
      const bool any_more = __any_sync(0xffffffffu, more);
#pragma unroll
      for (int u = 0; u < 4; ++u) cn[u] = -1;
      if (any_more) {
        if (more) ld_stream_c4(a.col + a1, cn);
      } else {
        const int an = (eb_n & ~3) + gl * 4;  // first step of the next group (its row pointers were requested above)
        if (an < ee_n) ld_stream_c4(a.col + an, cn);
      }
      // 2. this step's gathers and values
      double x[4][P];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        if (c[u] >= 0) {
          gather_row<P>(a.Tcur, c[u], x[u]);
        } else {
#pragma unroll
          for (int k = 0; k < P; ++k) x[u][k] = 0.0;
        }
      }
      double v[4] = {0.0, 0.0, 0.0, 0.0};
      if (a0 < ee) ld_stream_v4(a.val + a0, v);
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        if (c[u] < 0) v[u] = 0.0;
#pragma unroll
        for (int k = 0; k < P; ++k) acc[k] = fma(v[u], x[u][k], acc[k]);
      }
      if (!any_more) break;
      a0 = a1;
    }
#pragma unroll
    for (int o = G / 2; o > 0; o >>= 1) {
#pragma unroll
      for (int k = 0; k < P; ++k) acc[k] += __shfl_xor_sync(0xffffffffu, acc[k], o);
    }
    if (epi) {
      double y = acc[0];
#pragma unroll
      for (int k = 1; k < P; ++k)
        if (gl == k) y = acc[k];
      double tn = a.alpha * (y - a.shift * tc);
      if (a.gamma != 0.0) tn -= a.gamma * told;
      if (a.Tnew) a.Tnew[li] = tn;
      if (a.R) {
        double rv = a.c * tn + a.c_cur * tc;
        if (a.r_acc) rv += rold;
        st_stream(a.R + li, rv);
      }
    }
  }
}


// ---- flat kernel, second generation (8 lanes per row) --------------------------------------------------------
// Same walk as cheby_flat_kernel with compile-time variants:
//   HINT    cache policy of the operands.  The kernel is bound by the L1TEX pipe (77 % busy at 0.40 of the HBM
//           roofline, profiles/r01b_ncu_cheby_flat_c4.txt), so what the read-once matrix stream does to the L1
//           (allocation + fill of 317 MB per launch next to the 16 MB of signal rows worth keeping) matters:
//           0  values / columns evict-first (ld.cs), gathers default            (= cheby_flat_kernel)
//           1  values / columns L1::no_allocate, gathers default
//           2  values / columns and gathers L1::no_allocate
//           3  values / columns L1::no_allocate, gathers L1::evict_last
//   LAYOUT  0: a lane takes 4 consecutive entries (one 256-bit + one 128-bit load per lane)
//           1: the 8 lanes of a row take 8 consecutive entries per step (columns of neighbouring lanes are
//              neighbours in the sorted row, so gathers of one instruction can share 128-byte lines)
//   DOT     p = 1 only: per-CTA partial sums of T_cur[row] * y[row] (the Lanczos alpha) -> a.dot_partials
//   PEER    row-partitioned multi-GPU step: wait for the peers' flags of the previous term, store T_new into
//           every peer's buffer as well (P2P over NVLink), publish this term's flag when all CTAs are done
template <int HINT>
__device__ __forceinline__ void ld_val4_h(const double *p, double (&v)[4]) {
  if constexpr (HINT == 0)
    asm volatile("ld.global.cs.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(v[0]), "=d"(v[1]), "=d"(v[2]), "=d"(v[3]) : "l"(p));
  else
    asm volatile("ld.global.nc.L1::no_allocate.v4.f64 {%0,%1,%2,%3}, [%4];"
                 : "=d"(v[0]), "=d"(v[1]), "=d"(v[2]), "=d"(v[3])
                 : "l"(p));
}
template <int HINT>
__device__ __forceinline__ void ld_col4_h(const int32_t *p, int (&c)[4]) {
  if constexpr (HINT == 0)
    asm volatile("ld.global.cs.v4.s32 {%0,%1,%2,%3}, [%4];" : "=r"(c[0]), "=r"(c[1]), "=r"(c[2]), "=r"(c[3]) : "l"(p));
  else
    asm volatile("ld.global.nc.L1::no_allocate.v4.s32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(c[0]), "=r"(c[1]), "=r"(c[2]), "=r"(c[3])
                 : "l"(p));
}
template <int HINT>
__device__ __forceinline__ double ld_val1_h(const double *p) {
  double v;
  if constexpr (HINT == 0)
    asm volatile("ld.global.cs.f64 %0, [%1];" : "=d"(v) : "l"(p));
  else
    asm volatile("ld.global.nc.L1::no_allocate.f64 %0, [%1];" : "=d"(v) : "l"(p));
  return v;
}
template <int HINT>
__device__ __forceinline__ int ld_col1_h(const int32_t *p) {
  int v;
  if constexpr (HINT == 0)
    asm volatile("ld.global.cs.s32 %0, [%1];" : "=r"(v) : "l"(p));
  else
    asm volatile("ld.global.nc.L1::no_allocate.s32 %0, [%1];" : "=r"(v) : "l"(p));
  return v;
}
template <int HINT>
__device__ __forceinline__ void ldg256_h(const double *p, double &a, double &b, double &c, double &d) {
  if constexpr (HINT == 2)
    asm volatile("ld.global.nc.L1::no_allocate.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(a), "=d"(b), "=d"(c), "=d"(d) : "l"(p));
  else if constexpr (HINT == 3)
    asm volatile("ld.global.nc.L1::evict_last.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(a), "=d"(b), "=d"(c), "=d"(d) : "l"(p));
  else
    asm volatile("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(a), "=d"(b), "=d"(c), "=d"(d) : "l"(p));
}
template <int P, int HINT>
__device__ __forceinline__ void gather_row_h(const double *__restrict__ T, int32_t c, double (&x)[P]) {
  const double *t = T + (size_t)c * P;
  if constexpr (P % 4 == 0) {
#pragma unroll
    for (int k = 0; k < P; k += 4) ldg256_h<HINT>(t + k, x[k], x[k + 1], x[k + 2], x[k + 3]);
  } else {
    gather_row<P>(T, c, x);
  }
}

__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long *p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_sys(unsigned long long *p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long global_timer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

constexpr unsigned long long kFlagTimeoutNs = 4000000000ull;  // 4 s: a lost peer must not hang the GPU

// Thread 0 of the CTA waits until every rank's slot of my_flags has reached `epoch` (its stores of that
// phase are visible here), then releases the CTA.  A timeout sets *err_flag and lets the kernel run on.
__device__ __forceinline__ void peer_wait(const unsigned long long *my_flags, int world, int rank,
                                          unsigned long long epoch, int *err_flag) {
  if (epoch == 0) return;
  if (threadIdx.x == 0 && *reinterpret_cast<volatile int *>(err_flag) == 0) {  // after a time-out: never wait again
    const unsigned long long t0 = global_timer_ns();
    for (int w = 0; w < world; ++w) {
      if (w == rank) continue;
      while (ld_acquire_sys(my_flags + w) < epoch) {
        if (global_timer_ns() - t0 > kFlagTimeoutNs) {
          atomicExch(err_flag, 1);
          break;
        }
        __nanosleep(64);
      }
    }
  }
  __syncthreads();
}

// Called by every thread after its last P2P store.  The last CTA of the launch publishes `epoch` to the
// peers (fence + counter: all CTAs' stores are ordered before the flag stores).
__device__ __forceinline__ void peer_post(const StepArgs &a) {
  if (a.post_epoch == 0) return;
  // CTA barrier, then ONE system-scope fence per CTA (cumulative over the stores the barrier made visible to this
  // thread -- the pattern of a cooperative grid barrier); 1024 fences per CTA each wait for the NVLink acknowledgements
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence_system();
    const unsigned int prev = atomicAdd(a.done_ctr, 1u);
    if (prev == gridDim.x - 1) {
      *a.done_ctr = 0;  // the next launch starts after this one has ended
      __threadfence_system();
      if (a.scal_partials != nullptr) {
        // this rank's partial sum (CTA partials added in a fixed order) goes into slot [epoch & 3][rank] of every
        // rank; all ranks then add the slots in rank order and obtain the same global sum bit for bit
        double sum = 0.0;
        for (int i = 0; i < a.n_scal_partials; ++i) sum += __ldcg(a.scal_partials + i);
        const int slot = (int)(a.post_epoch & 3ull) * 8 + a.rank;
        a.my_scal[slot] = sum;
        for (int w = 0; w < a.n_peers; ++w) a.peer_scal[w][slot] = sum;
        __threadfence_system();
      }
      for (int w = 0; w < a.n_peers; ++w) st_release_sys(a.peer_flags[w], a.post_epoch);
    }
  }
}

// global sum published in phase `epoch`: the ranks' slots in rank order (call after peer_wait(epoch))
__device__ __forceinline__ double peer_scalar_sum(const double *my_scal, int world, unsigned long long epoch) {
  double s = 0.0;
  const int base = (int)(epoch & 3ull) * 8;
  for (int r = 0; r < world; ++r) s += __ldcg(my_scal + base + r);
  return s;
}

template <int P, int TB, int HINT, int LAYOUT, bool DOT, bool PEER>
__global__ void __launch_bounds__(TB, 1) cheby_flat2_kernel(const StepArgs a, const int64_t n_rows) {
  static_assert(!DOT || P == 1, "the fused dot product is the Lanczos (p = 1) path");
  constexpr int G = 8;
  constexpr int RPW = 32 / G;
  const int lane = threadIdx.x & 31;
  const int gid = lane / G, gl = lane % G;
  if constexpr (PEER) peer_wait(a.my_flags, a.world, a.rank, a.wait_epoch, a.err_flag);
  int64_t rb0 = ((int64_t)blockIdx.x * (TB / 32) + (threadIdx.x >> 5)) * RPW, r_end = n_rows;
  int64_t rstep = (int64_t)gridDim.x * (TB / 32) * RPW;
  if (a.blk != nullptr) {
    const int64_t b0 = ((int64_t)blockIdx.x * a.n_blk) / gridDim.x, b1 = ((int64_t)(blockIdx.x + 1) * a.n_blk) / gridDim.x;
    rb0 = __ldg(a.blk + b0) + (int64_t)(threadIdx.x >> 5) * RPW;
    r_end = __ldg(a.blk + b1);
    rstep = (int64_t)(TB / 32) * RPW;
  }
  double dot = 0.0;
  for (int64_t rb = rb0; rb < r_end; rb += rstep) {  // warp-uniform
    const int64_t r = rb + gid;
    const bool act = r < r_end;
    int eb = 0, ee = 0;
    if (act) {
      eb = __ldg(a.row_ptr + r);
      ee = __ldg(a.row_ptr + r + 1);
    }
    const bool epi = act && gl < P;
    const size_t li = (size_t)r * P + gl;
    double tc = 0.0, told = 0.0, rold = 0.0;
    if (epi) {  // requested first so their latency overlaps the row product
      tc = __ldg(a.Tcur + (size_t)(a.row0 + r) * P + gl);
      if (a.gamma != 0.0) told = ld_stream(a.Told + li);
      if (a.R != nullptr && a.r_acc) rold = ld_stream(a.R + li);
    }
    double acc[P];
#pragma unroll
    for (int k = 0; k < P; ++k) acc[k] = 0.0;
    if constexpr (LAYOUT == 0) {
      for (int a0 = (eb & ~3) + gl * 4; a0 < ee; a0 += 4 * G) {
        double v[4];
        int c[4];
        ld_val4_h<HINT>(a.val + a0, v);
        ld_col4_h<HINT>(a.col + a0, c);
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int eu = a0 + u;
          if (eu < eb || eu >= ee) {
            v[u] = 0.0;
            c[u] = -1;
          }
        }
        constexpr int U = P <= 4 ? 4 : 2;  // gathers in flight per lane
#pragma unroll
        for (int h = 0; h < 4; h += U) {
          double x[U][P];
#pragma unroll
          for (int u = 0; u < U; ++u) {
            if (c[h + u] >= 0) {
              gather_row_h<P, HINT>(a.Tcur, c[h + u], x[u]);
            } else {
#pragma unroll
              for (int k = 0; k < P; ++k) x[u][k] = 0.0;
            }
          }
#pragma unroll
          for (int u = 0; u < U; ++u) {
#pragma unroll
            for (int k = 0; k < P; ++k) acc[k] = fma(v[h + u], x[u][k], acc[k]);
          }
        }
      }
    } else {
      constexpr int U = P <= 4 ? 4 : 2;
      for (int a0 = eb + gl; a0 < ee; a0 += U * G) {
        double v[U];
        int c[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const int eu = a0 + u * G;
          if (eu < ee) {
            v[u] = ld_val1_h<HINT>(a.val + eu);
            c[u] = ld_col1_h<HINT>(a.col + eu);
          } else {
            v[u] = 0.0;
            c[u] = -1;
          }
        }
        double x[U][P];
#pragma unroll
        for (int u = 0; u < U; ++u) {
          if (c[u] >= 0) {
            gather_row_h<P, HINT>(a.Tcur, c[u], x[u]);
          } else {
#pragma unroll
            for (int k = 0; k < P; ++k) x[u][k] = 0.0;
          }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
#pragma unroll
          for (int k = 0; k < P; ++k) acc[k] = fma(v[u], x[u][k], acc[k]);
        }
      }
    }
    __syncwarp();
#pragma unroll
    for (int o = G / 2; o > 0; o >>= 1) {
#pragma unroll
      for (int k = 0; k < P; ++k) acc[k] += __shfl_xor_sync(0xffffffffu, acc[k], o);
    }
    if (epi) {
      double y = acc[0];
#pragma unroll
      for (int k = 1; k < P; ++k)
        if (gl == k) y = acc[k];
      double tn = a.alpha * (y - a.shift * tc);
      if (a.gamma != 0.0) tn -= a.gamma * told;
      if constexpr (DOT) dot = fma(tc, tn, dot);
      double outv = tn;
      if (a.R) {
        double rv = a.c * tn + a.c_cur * tc;
        if (a.r_acc) rv += rold;
        st_stream(a.R + li, rv);
        if constexpr (PEER) {
          if (a.store_r) outv = rv;
        }
      }
      if (a.Tnew) a.Tnew[li] = outv;  // gathered by the next step: keep cacheable
      if constexpr (PEER) {
        if (a.Tnew && a.peer_store) {
          // only the peers whose rows reference this row as a column need it (the finished R goes to everyone)
          const unsigned hm = (a.halo != nullptr && !a.store_r) ? (unsigned)__ldg(a.halo + r) : 0xffu;
#pragma unroll 1
          for (int w = 0; w < a.n_peers; ++w)
            if ((hm >> w) & 1u) a.peer_tnew[w][li] = outv;
        }
      }
    }
  }
  if constexpr (DOT) {
    __shared__ double dot_sh[33];
    const double s = block_sum_fwd(dot, dot_sh);
    if (threadIdx.x == 0) a.dot_partials[blockIdx.x] = s;
  }
  if constexpr (PEER) peer_post(a);
}

typedef void (*FlatKernel)(const StepArgs, const int64_t);

template <int P, int TB>
static FlatKernel pick_flat_group(int G) {
  if constexpr (P <= 4) {
    if (G == 4) return cheby_flat_kernel<P, 4, TB>;
  }
  if (G <= 8) return cheby_flat_kernel<P, 8, TB>;
  if (G <= 16) return cheby_flat_kernel<P, 16, TB>;
  return cheby_flat_kernel<P, 32, TB>;
}

template <int TB>
static FlatKernel pick_flat(int P, int G) {
  switch (P) {
    case 1: return pick_flat_group<1, TB>(G);
    case 2: return pick_flat_group<2, TB>(G);
    case 3: return pick_flat_group<3, TB>(G);
    case 4: return pick_flat_group<4, TB>(G);
    case 5: return pick_flat_group<5, TB>(G);
    case 6: return pick_flat_group<6, TB>(G);
    case 7: return pick_flat_group<7, TB>(G);
    case 8: return pick_flat_group<8, TB>(G);
    default: return nullptr;
  }
}

// ---- dispatch of the second-generation flat kernel -------------------------------------------------------------
template <int P, int HINT>
static FlatKernel pick_flat2_layout(int layout) {
  if (layout == 1) return cheby_flat2_kernel<P, 1024, HINT, 1, false, false>;
  return cheby_flat2_kernel<P, 1024, HINT, 0, false, false>;
}
template <int P>
static FlatKernel pick_flat2_hint(int hint, int layout) {
  switch (hint) {
    case 1: return pick_flat2_layout<P, 1>(layout);
    case 2: return pick_flat2_layout<P, 2>(layout);
    case 3: return pick_flat2_layout<P, 3>(layout);
    default: return pick_flat2_layout<P, 0>(layout);
  }
}
static FlatKernel pick_flat2(int P, int hint, int layout) {
  switch (P) {
    case 1: return pick_flat2_hint<1>(hint, layout);
    case 2: return pick_flat2_hint<2>(hint, layout);
    case 3: return pick_flat2_hint<3>(hint, layout);
    case 4: return pick_flat2_hint<4>(hint, layout);
    case 5: return pick_flat2_hint<5>(hint, layout);
    case 6: return pick_flat2_hint<6>(hint, layout);
    case 7: return pick_flat2_hint<7>(hint, layout);
    case 8: return pick_flat2_hint<8>(hint, layout);
    default: return nullptr;
  }
}
template <int HINT>
static FlatKernel pick_flat2_peer_h(int P) {
  switch (P) {
    case 1: return cheby_flat2_kernel<1, 1024, HINT, 1, false, true>;
    case 2: return cheby_flat2_kernel<2, 1024, HINT, 1, false, true>;
    case 3: return cheby_flat2_kernel<3, 1024, HINT, 1, false, true>;
    case 4: return cheby_flat2_kernel<4, 1024, HINT, 1, false, true>;
    case 5: return cheby_flat2_kernel<5, 1024, HINT, 1, false, true>;
    case 6: return cheby_flat2_kernel<6, 1024, HINT, 1, false, true>;
    case 7: return cheby_flat2_kernel<7, 1024, HINT, 1, false, true>;
    case 8: return cheby_flat2_kernel<8, 1024, HINT, 1, false, true>;
    default: return nullptr;
  }
}
static FlatKernel pick_flat2_peer(int P, int hint) {
  return hint >= 2 ? pick_flat2_peer_h<2>(P) : pick_flat2_peer_h<1>(P);
}
static FlatKernel pick_flat2_dot(int hint, bool peer) {
  (void)hint;
  return peer ? cheby_flat2_kernel<1, 1024, 1, 1, true, true> : cheby_flat2_kernel<1, 1024, 1, 1, true, false>;
}

// Measured on the 500k-cell graph of config 4 (tools/r02_probe.py, profiles/r02_probe_spmm_variants.txt): the cache
// policy of the matrix stream does not move the p <= 4 time (the kernel is bound by L1TEX data-pipe wavefronts, one
// per gathered row, not by what the L1 holds); lane-consecutive entries gain 3 % at p = 4 and 15 % at p = 1;
// at p = 8 (two wavefronts per gathered row) not allocating the gathers in L1 gains 12 %.
static int auto_hint(int P, int hint) { return hint >= 0 ? hint : (P > 4 ? 2 : 1); }
static int auto_layout(int P, int layout) { return layout >= 0 ? layout : (P > 4 ? 0 : 1); }

static int choose_group(const meld_b200_graph *g, int P) {
  int G = tuning().group;
  if (G != 4 && G != 8 && G != 16 && G != 32) {
    const double avg = g->n_rows > 0 ? (double)g->nnz / (double)g->n_rows : 0.0;
    G = avg <= 12.0 ? 4 : (avg <= 80.0 ? 8 : (avg <= 256.0 ? 16 : 32));
  }
  if (G < P) G = 8;
  return G;
}

static int launch_step(const meld_b200_graph *g, StepArgs a, int P, cudaStream_t stream, int *grid_out = nullptr) {
  const Tuning &t = tuning();
  const int G = choose_group(g, P);
  const bool want_peer = a.n_peers > 0 || a.wait_epoch != 0 || a.post_epoch != 0;
  const bool want_dot = a.dot_partials != nullptr;
  // very short (<= 12 entries) or very long (> 256) rows keep the round-1 kernels with 4 / 32 lanes per row
  const bool g8 = (t.flat_group > 0 ? t.flat_group : G) == 8 || (t.flat_group <= 0 && G == 16);
  if ((t.flat_gen == 1 && g8) || want_peer || want_dot) {  // second-generation flat kernel (8 lanes per row)
    const int hint = auto_hint(P, t.flat_hint), layout = auto_layout(P, t.flat_layout);
    FlatKernel fk = want_dot ? pick_flat2_dot(hint, want_peer)
                             : (want_peer ? pick_flat2_peer(P, hint) : pick_flat2(P, hint, layout));
    MELD_REQUIRE(fk != nullptr, "cheby_step: p=%d outside 1..8", P);
    MELD_REQUIRE(!want_dot || P == 1, "cheby_step: the fused dot product needs p = 1");
    a.row_ptr = g->row_ptr.p;
    a.col = g->col.p;
    a.val = g->val.p;
    a.row0 = g->row0;
    a.blk = (t.flat_sched == 1 && g->blk.p != nullptr && g->n_blk >= 4 * sm_count()) ? g->blk.p : nullptr;
    a.n_blk = g->n_blk;
    int grid = sm_count();
    const int64_t groups = ceil_div(g->n_rows, 4);
    if ((int64_t)grid * 32 > groups) grid = (int)ceil_div(groups, 32);
    if (grid < 1) grid = 1;
    if (a.scal_partials != nullptr) a.n_scal_partials = grid;  // one partial per CTA of this launch
    fk<<<grid, 1024, 0, stream>>>(a, g->n_rows);
    MELD_LAUNCH_CHECK();
    if (grid_out) *grid_out = grid;
    return 0;
  }
  {  // round-1 flat kernels (4 / 16 / 32 lanes per row, 768-thread and pipelined variants)
    int Gf = t.flat_group > 0 ? t.flat_group : G;
    if (Gf < P) Gf = 8;
    const bool wide = t.flat_threads != 768;
    FlatKernel fk = wide ? pick_flat<1024>(P, Gf) : pick_flat<768>(P, Gf);
    if ((t.flat_pipe == 1 || (t.flat_pipe == 2 && P == 1)) && P <= 4 && Gf == 8) {
      switch (P) {
        case 1: fk = wide ? cheby_flat_pipe_kernel<1, 8, 1024> : cheby_flat_pipe_kernel<1, 8, 768>; break;
        case 2: fk = wide ? cheby_flat_pipe_kernel<2, 8, 1024> : cheby_flat_pipe_kernel<2, 8, 768>; break;
        case 3: fk = wide ? cheby_flat_pipe_kernel<3, 8, 1024> : cheby_flat_pipe_kernel<3, 8, 768>; break;
        default: fk = wide ? cheby_flat_pipe_kernel<4, 8, 1024> : cheby_flat_pipe_kernel<4, 8, 768>; break;
      }
    }
    MELD_REQUIRE(fk != nullptr, "cheby_step: p=%d outside 1..8", P);
    a.row_ptr = g->row_ptr.p;
    a.col = g->col.p;
    a.val = g->val.p;
    a.row0 = g->row0;
    a.blk = (t.flat_sched == 1 && g->blk.p != nullptr && g->n_blk >= 4 * sm_count()) ? g->blk.p : nullptr;
    a.n_blk = g->n_blk;
    int grid = sm_count() * (t.ctas_per_sm > 0 ? t.ctas_per_sm : 1);
    const int threads = wide ? 1024 : 768;
    const int64_t groups = ceil_div(g->n_rows, 32 / (Gf <= 4 ? 4 : (Gf <= 8 ? 8 : (Gf <= 16 ? 16 : 32))));
    if ((int64_t)grid * (threads / 32) > groups) grid = (int)ceil_div(groups, threads / 32);
    if (grid < 1) grid = 1;
    fk<<<grid, threads, 0, stream>>>(a, g->n_rows);
    MELD_LAUNCH_CHECK();
    return 0;
  }
}

static int grid_for(int64_t n, int threads);

static int ensure_work(meld_b200_graph *g, size_t count) {
  if (g->work.n >= count) return 0;
  return g->work.alloc(count);
}

// Internal row width of a p-column signal.  Only 32-byte aligned rows can be gathered with 256-bit loads, so padding
// p = 3 / 5 / 6 / 7 with zero columns to 4 / 8 doubles turns three to seven narrow loads per gathered row into one or
// two wide ones.  MEASURED (round 2, bench.py): it is SLOWER -- config 5 (2M cells, p = 6 -> 8) 796 -> 894 us per
// launch, config 2 (50k cells, p = 3 -> 4) 5.0 -> 6.0 ms per filter: the gather cost follows the 32-byte sectors a row
// occupies (two for 48 as for 64 bytes), not the number of load instructions, and the padded columns add a third to
// the epilogue's vector traffic.  Off by default (tuning key pad_width = 1 enables it for experiments).
static inline int padded_width(int p) {
  if (!tuning().pad_width) return p;
  return p <= 2 ? p : (p <= 4 ? 4 : 8);
}

// rows of an (n, p) array through a permutation.  The CALLER's array has p columns, the internal (graph-order) one pw
// >= p (zero padded): gather: internal[a] = caller[perm[a]]; scatter: caller[perm[a]] = internal[a].
__global__ void permute_rows_kernel(const double *__restrict__ in, const int32_t *__restrict__ perm, int64_t n, int p,
                                    int pw, int scatter, double *__restrict__ out) {
  const int64_t total = n * pw;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
    const int64_t i = t / pw;
    const int j = (int)(t - i * pw);
    const int64_t o = perm ? (int64_t)perm[i] : i;
    if (scatter) {
      if (j < p) out[o * p + j] = in[t];
    } else {
      out[t] = j < p ? in[o * p + j] : 0.0;
    }
  }
}

// ---- Lanczos helpers ----------------------------------------------------------------
// Three-term Lanczos on UNNORMALISED vectors w_j = beta_j v_j, two launches per step:
//   1. y = L w_j with the fused partial sums of w_j . y        (cheby_flat2_kernel<1, ..., DOT>)
//   2. alpha_j = (w_j . y) / beta_j^2 ;  w_{j+1} = (y - alpha_j w_j) / beta_j - (beta_j / beta_{j-1}) w_{j-1}
//      written over w_{j-1}; partial sums of |w_{j+1}|^2 = beta_{j+1}^2                  (lanczos_axpy_kernel)
// All reductions are two-stage with a fixed number of partials (deterministic).  Scalars stay on the device;
// the host only downloads (alpha, beta) every few steps to test the Ritz residual.
constexpr int kRedBlocks = 256;  // partial sums per reduction (fixed => deterministic)
constexpr int kRedThreads = 256;

__device__ __forceinline__ double block_sum(double v, double *sh) { return block_sum_fwd(v, sh); }

__device__ __forceinline__ double sum_partials(const double *partials, double *sh) {
  double v = (threadIdx.x < kRedBlocks) ? partials[threadIdx.x] : 0.0;
  return block_sum(v, sh);
}

__global__ void lanczos_init_kernel(double *v, int64_t n, double *partials) {
  __shared__ double sh[33];
  double s = 0.0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    uint64_t z = (uint64_t)i * 0x9E3779B97F4A7C15ull + 0xD1B54A32D192ED03ull;  // splitmix64
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    z ^= z >> 31;
    const double x = (double)(z >> 11) * (1.0 / 9007199254740992.0) - 0.5;
    v[i] = x;
    s += x * x;
  }
  s = block_sum(s, sh);
  if (threadIdx.x == 0) partials[blockIdx.x] = s;
}

// pa: partials of w_j . y; pb_cur / pb_prev: partials of |w_j|^2 / |w_{j-1}|^2; pb_next receives |w_{j+1}|^2.
// w_prev is overwritten with w_{j+1}.
__global__ void __launch_bounds__(kRedThreads) lanczos_axpy_kernel(const double *__restrict__ y,
                                                                   const double *__restrict__ w_cur, double *w_prev,
                                                                   int64_t n, const double *pa, const double *pb_cur,
                                                                   const double *pb_prev, int j, double *alpha_arr,
                                                                   double *beta_arr, double *pb_next) {
  __shared__ double sh[33];
  const double dot = sum_partials(pa, sh);
  const double b2 = sum_partials(pb_cur, sh);
  const double bp2 = j > 0 ? sum_partials(pb_prev, sh) : 1.0;
  const double beta = sqrt(b2), beta_prev = sqrt(bp2);
  const double alpha = b2 > 0.0 ? dot / b2 : 0.0;
  const double inv = beta > 0.0 ? 1.0 / beta : 0.0;
  const double ratio = (j > 0 && beta_prev > 0.0) ? beta / beta_prev : 0.0;
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    alpha_arr[j] = alpha;
    beta_arr[j] = beta;
  }
  double s = 0.0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const double x = (y[i] - alpha * w_cur[i]) * inv - ratio * w_prev[i];
    w_prev[i] = x;
    s = fma(x, x, s);
  }
  s = block_sum(s, sh);
  if (threadIdx.x == 0) pb_next[blockIdx.x] = s;
}

// Row-partitioned Lanczos step 2 (see meld_b200_estimate_lmax_dist): the dot product w_j . (L w_j) and |w_j|^2 are
// the rank-ordered sums of the slots the ranks published in phases `dot_epoch` / `norm_epoch`; this rank updates
// its rows of w_{j+1} (over w_{j-1}) in its own and in every peer's vector, and its last CTA publishes |w_{j+1}|^2
// of its rows.  pb0: partials of the full |w_0|^2 (j == 0: every rank generated the whole start vector).
__global__ void __launch_bounds__(kRedThreads) lanczos_axpy_peer_kernel(const StepArgs a, const double *__restrict__ y_loc,
                                                                        const double *__restrict__ w_cur, double *w_prev,
                                                                        int64_t row0, int64_t nloc,
                                                                        unsigned long long dot_epoch,
                                                                        unsigned long long norm_epoch, const double *pb0,
                                                                        int j, double *alpha_arr, double *beta_arr,
                                                                        double *pb_out) {
  __shared__ double sh[33];
  peer_wait(a.my_flags, a.world, a.rank, a.wait_epoch, a.err_flag);
  const double dot = peer_scalar_sum(a.my_scal, a.world, dot_epoch);
  const double b2 = j == 0 ? sum_partials(pb0, sh) : peer_scalar_sum(a.my_scal, a.world, norm_epoch);
  const double beta = sqrt(b2);
  const double beta_prev = j > 0 ? __ldcg(beta_arr + j - 1) : 1.0;
  const double alpha = b2 > 0.0 ? dot / b2 : 0.0;
  const double inv = beta > 0.0 ? 1.0 / beta : 0.0;
  const double ratio = (j > 0 && beta_prev > 0.0) ? beta / beta_prev : 0.0;
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    alpha_arr[j] = alpha;
    beta_arr[j] = beta;
  }
  double s = 0.0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nloc; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t gi = row0 + i;
    const double x = (y_loc[i] - alpha * w_cur[gi]) * inv - ratio * w_prev[gi];
    w_prev[gi] = x;
    const unsigned hm = a.halo != nullptr ? (unsigned)__ldg(a.halo + i) : 0xffu;
    for (int w = 0; w < a.n_peers; ++w)
      if ((hm >> w) & 1u) a.peer_tnew[w][i] = x;  // peers' copies of this rank's rows (those that gather them)
    s = fma(x, x, s);
  }
  s = block_sum(s, sh);
  if (threadIdx.x == 0) pb_out[blockIdx.x] = s;
  peer_post(a);
}

// beta_arr[k] = |w_k| from the slots of phase `norm_epoch` (waits until every rank has published)
__global__ void __launch_bounds__(32) lanczos_tail_peer_kernel(const StepArgs a, unsigned long long norm_epoch, int k,
                                                               double *beta_arr) {
  peer_wait(a.my_flags, a.world, a.rank, a.wait_epoch, a.err_flag);
  if (threadIdx.x == 0) beta_arr[k] = sqrt(peer_scalar_sum(a.my_scal, a.world, norm_epoch));
}

// beta_arr[k] = |w_k| for the newest vector (its partials are complete once the axpy of step k-1 has run)
__global__ void __launch_bounds__(kRedThreads) lanczos_tail_kernel(const double *pb, int k, double *beta_arr) {
  __shared__ double sh[33];
  const double b2 = sum_partials(pb, sh);
  if (threadIdx.x == 0) beta_arr[k] = sqrt(b2);
}

// All eigenvalues of the symmetric tridiagonal (diag d[0..k), off-diagonal e[0..k-1)) by the implicit QL
// iteration, together with the LAST component of every normalised eigenvector (what the Ritz residual
// |beta_k s_last| needs).  d is overwritten with the eigenvalues, zl with those components.
static bool tridiag_ql_last_row(std::vector<double> &d, std::vector<double> &e, std::vector<double> &zl, int k) {
  for (int i = 0; i < k; ++i) zl[(size_t)i] = i == k - 1 ? 1.0 : 0.0;
  if (k == 1) return true;
  e[(size_t)k - 1] = 0.0;
  for (int l = 0; l < k; ++l) {
    int iter = 0, m;
    do {
      for (m = l; m < k - 1; ++m) {
        const double dd = fabs(d[(size_t)m]) + fabs(d[(size_t)m + 1]);
        if (fabs(e[(size_t)m]) <= 2.3e-16 * dd) break;
      }
      if (m != l) {
        if (++iter > 60) return false;
        double g = (d[(size_t)l + 1] - d[(size_t)l]) / (2.0 * e[(size_t)l]);
        double r = hypot(g, 1.0);
        g = d[(size_t)m] - d[(size_t)l] + e[(size_t)l] / (g + copysign(r, g));
        double sn = 1.0, cs = 1.0, pp = 0.0;
        int i;
        for (i = m - 1; i >= l; --i) {
          double f = sn * e[(size_t)i];
          const double b = cs * e[(size_t)i];
          r = hypot(f, g);
          e[(size_t)i + 1] = r;
          if (r == 0.0) {
            d[(size_t)i + 1] -= pp;
            e[(size_t)m] = 0.0;
            break;
          }
          sn = f / r;
          cs = g / r;
          g = d[(size_t)i + 1] - pp;
          r = (d[(size_t)i] - g) * sn + 2.0 * cs * b;
          pp = sn * r;
          d[(size_t)i + 1] = g + pp;
          g = cs * r - b;
          f = zl[(size_t)i + 1];
          zl[(size_t)i + 1] = sn * zl[(size_t)i] + cs * f;
          zl[(size_t)i] = cs * zl[(size_t)i] - sn * f;
        }
        if (r == 0.0 && i >= l) continue;
        d[(size_t)l] -= pp;
        e[(size_t)l] = g;
        e[(size_t)m] = 0.0;
      }
    } while (m != l);
  }
  return true;
}

// ---- signal helpers --------------------------------------------------------------------
__global__ void count_codes_kernel(const int32_t *__restrict__ codes, int64_t n, int p, unsigned long long *cnt) {
  __shared__ unsigned int sc[64];
  if (threadIdx.x < 64) sc[threadIdx.x] = 0;
  __syncthreads();
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = codes[i];
    if (c >= 0 && c < p) {
      if (p <= 64)
        atomicAdd(&sc[c], 1u);
      else
        atomicAdd(&cnt[c], 1ull);
    }
  }
  __syncthreads();
  if (p <= 64 && threadIdx.x < p && sc[threadIdx.x]) atomicAdd(&cnt[threadIdx.x], (unsigned long long)sc[threadIdx.x]);
}

__global__ void fill_indicator_kernel(const int32_t *__restrict__ codes, int64_t n, int p, int normalize,
                                      const unsigned long long *__restrict__ cnt, double *__restrict__ S) {
  const int64_t total = n * p;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
    const int64_t i = t / p;
    const int j = (int)(t - i * p);
    double v = 0.0;
    if (codes[i] == j) v = normalize ? 1.0 / (double)cnt[j] : 1.0;
    S[t] = v;
  }
}

__global__ void l1_normalize_rows_kernel(const double *__restrict__ in, int64_t n, int p, double *__restrict__ out) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    double s = 0.0;
    for (int j = 0; j < p; ++j) s += fabs(in[i * p + j]);
    if (s == 0.0) s = 1.0;
    for (int j = 0; j < p; ++j) out[i * p + j] = in[i * p + j] / s;
  }
}

// R_f = sum_k C[f][k] T_k for FC filters at a time (the shared-basis parameter sweep): every thread owns one
// element of the (n, p) signal, walks the stored basis once and keeps FC accumulators; products and sums are
// rounded separately, in the order of PyGSP's loop (r = c0/2 T0 + c1 T1, r = r + c_k T_k).  Output rows go
// back to the caller's cell order through perm.
template <int FC>
__global__ void __launch_bounds__(256) combine_basis_kernel(const double *__restrict__ T, size_t slot_stride,
                                                           int n_terms, const double *__restrict__ C, int f0, int nf,
                                                           const int32_t *__restrict__ perm, int64_t n, int p, int pw,
                                                           double *__restrict__ R) {
  extern __shared__ double cs[];  // FC x n_terms
  for (int t = threadIdx.x; t < FC * n_terms; t += blockDim.x) {
    const int f = t / n_terms, k = t - f * n_terms;
    cs[t] = f < nf ? C[(size_t)(f0 + f) * n_terms + k] : 0.0;
  }
  __syncthreads();
  const int64_t total = n * pw, out_total = n * p;  // the basis is stored pw wide (zero-padded), the output p wide
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    const int64_t i = e / pw;
    const int j = (int)(e - i * pw);
    if (j >= p) continue;
    double acc[FC];
#pragma unroll
    for (int f = 0; f < FC; ++f) acc[f] = 0.0;
    for (int k = 0; k < n_terms; ++k) {
      const double t = T[(size_t)k * slot_stride + e];
#pragma unroll
      for (int f = 0; f < FC; ++f) acc[f] = __dadd_rn(acc[f], __dmul_rn(cs[f * n_terms + k], t));
    }
    const int64_t o = perm ? (int64_t)perm[i] : i;
#pragma unroll
    for (int f = 0; f < FC; ++f)
      if (f < nf) R[(size_t)(f0 + f) * out_total + o * p + j] = acc[f];
  }
}

// Last phase of a row-partitioned filter: wait until every peer has stored its rows of R into this rank's
// buffer, bring the rows back into the caller's cell order, then tell the peers this rank's buffers are free.
__global__ void __launch_bounds__(256) dist_unpermute_kernel(const StepArgs a, const double *__restrict__ X,
                                                            const int32_t *__restrict__ perm, int64_t n, int p, int pw,
                                                            double *__restrict__ out) {
  peer_wait(a.my_flags, a.world, a.rank, a.wait_epoch, a.err_flag);
  const int64_t total = n * pw;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
    const int64_t i = t / pw;
    const int j = (int)(t - i * pw);
    const int64_t o = perm ? (int64_t)perm[i] : i;
    if (j < p) out[o * p + j] = __ldcg(X + t);  // written by peers over NVLink: read through L2
  }
  peer_post(a);
}

static int grid_for(int64_t n, int threads) {
  int64_t b = ceil_div(n > 0 ? n : 1, threads);
  int64_t cap = (int64_t)sm_count() * 8;
  if (cap < 1) cap = 1184;
  return (int)(b < cap ? b : cap);
}

}  // namespace meld

using namespace meld;

extern "C" {

int meld_b200_cheby_step(meld_b200_graph_t *g, const double *T_cur, const double *T_old, double *T_new, double *R,
                         int p, double alpha, double shift, double gamma, double c, double c_cur, int r_accumulate,
                         void *stream_) {
  meld::use_stream((cudaStream_t)stream_);
  MELD_REQUIRE(g && T_cur, "cheby_step: NULL argument");
  MELD_REQUIRE(p >= 1 && p <= 8, "cheby_step: p=%d outside 1..8", p);
  MELD_REQUIRE(gamma == 0.0 || T_old != nullptr, "cheby_step: gamma != 0 needs T_old");
  MELD_REQUIRE((const double *)T_new != T_cur, "cheby_step: T_new may not alias T_cur");
  MELD_REQUIRE(((uintptr_t)T_cur & 31) == 0, "cheby_step: T_cur must be 32-byte aligned");
  StepArgs a{};
  a.Tcur = T_cur;
  a.Told = T_old;
  a.Tnew = T_new;
  a.R = R;
  a.alpha = alpha;
  a.shift = shift;
  a.gamma = gamma;
  a.c = c;
  a.c_cur = c_cur;
  a.r_acc = r_accumulate;
  return launch_step(g, a, p, (cudaStream_t)stream_);
}

int meld_b200_graph_permute_signal(const meld_b200_graph_t *g, const double *in, int p, int to_internal, double *out,
                                   void *stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  meld::use_stream(stream);
  MELD_REQUIRE(g && in && out && p >= 1 && in != out, "graph_permute_signal: bad argument");
  const int64_t n = g->n_cols;
  permute_rows_kernel<<<grid_for(n * p, 256), 256, 0, stream>>>(in, g->perm.p, n, p, p, to_internal ? 0 : 1, out);
  MELD_LAUNCH_CHECK();
  return 0;
}

int meld_b200_cheby_filter(meld_b200_graph_t *g, double lmax, const double *coeffs_host, int n_coeffs, const double *S,
                           int p, double *R, void *stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  meld::use_stream(stream);
  MELD_REQUIRE(g && coeffs_host && S && R, "cheby_filter: NULL argument");
  MELD_REQUIRE(n_coeffs >= 2, "cheby_filter: need at least 2 coefficients (got %d)", n_coeffs);
  MELD_REQUIRE(p >= 1 && p <= 8, "cheby_filter: p=%d outside 1..8 (split the signal into column chunks)", p);
  MELD_REQUIRE(lmax > 0.0 && isfinite(lmax), "cheby_filter: lmax=%g", lmax);
  MELD_REQUIRE(g->row0 == 0 && g->n_rows == g->n_cols, "cheby_filter: needs the full operator (use cheby_step)");
  MELD_REQUIRE(S != R, "cheby_filter: R may not alias S");
  const int64_t n = g->n_rows;
  const int pw = padded_width(p);
  // four padded work arrays (S and R in graph order, two recurrence buffers): even length + 2 so
  // the kernel's 16-byte aligned bulk copies of row slices stay inside the allocation
  const size_t len = ((size_t)n * pw + 2 + 3) & ~(size_t)3;  // multiple of 32 bytes: 256-bit gathers on the direct path
  MELD_CHECK(ensure_work(g, 4 * len));
  double *Sp = g->work.p, *Rp = Sp + len, *Ta = Rp + len, *Tb = Ta + len;
  const int pgrid = grid_for(n * pw, 256);
  permute_rows_kernel<<<pgrid, 256, 0, stream>>>(S, g->perm.p, n, p, pw, /*scatter=*/0, Sp);  // S into graph order
  MELD_LAUNCH_CHECK();
  const double a1 = lmax / 2.0, a2 = lmax / 2.0;
  // k = 1: T1 = (L S - a2 S)/a1 ; R = c0/2 S + c1 T1
  StepArgs a{};
  a.Tcur = Sp;
  a.Told = nullptr;
  a.Tnew = (n_coeffs > 2) ? Ta : nullptr;
  a.R = Rp;
  a.alpha = 1.0 / a1;
  a.shift = a2;
  a.gamma = 0.0;
  a.c = coeffs_host[1];
  a.c_cur = 0.5 * coeffs_host[0];
  a.r_acc = 0;
  MELD_CHECK(launch_step(g, a, pw, stream));
  // k = 2 still reads T0 = S, so T2 goes to the second buffer; from k = 3 on T_k overwrites T_{k-2}.
  const double *cur = Ta, *old = Sp;
  for (int k = 2; k < n_coeffs; ++k) {
    double *nxt = (k == 2) ? Tb : const_cast<double *>(old);
    a.Tcur = cur;
    a.Told = old;
    a.Tnew = (k + 1 < n_coeffs) ? nxt : nullptr;  // the last term only feeds R
    a.alpha = 2.0 / a1;
    a.gamma = 1.0;
    a.c = coeffs_host[k];
    a.c_cur = 0.0;
    a.r_acc = 1;
    MELD_CHECK(launch_step(g, a, pw, stream));
    old = cur;
    cur = nxt;
  }
  permute_rows_kernel<<<pgrid, 256, 0, stream>>>(Rp, g->perm.p, n, p, pw, /*scatter=*/1, R);  // back to caller order
  MELD_LAUNCH_CHECK();
  return 0;
}

static void fill_peer_args(StepArgs &a, const meld_b200_dist *d, int which_buf, int64_t row0, int p) {
  a.n_peers = 0;
  for (int w = 0; w < d->world; ++w) {
    if (w == d->rank) continue;
    a.peer_tnew[a.n_peers] = d->buf(w, which_buf) + (size_t)row0 * p;
    a.peer_flags[a.n_peers] = d->flags(w) + d->rank;
    ++a.n_peers;
  }
  a.peer_store = 1;
  {
    int k = 0;
    for (int w = 0; w < d->world; ++w)
      if (w != d->rank) a.peer_scal[k++] = d->scal(w);
  }
  a.my_scal = d->scal(d->rank);
  a.my_flags = d->flags(d->rank);
  a.done_ctr = d->ctr();
  a.err_flag = d->err();
  a.world = d->world;
  a.rank = d->rank;
}

int meld_b200_cheby_filter_dist(meld_b200_graph_t *gs, meld_b200_dist_t *d, double lmax, const double *coeffs_host,
                                int n_coeffs, const double *S, int p, double *R, void *stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  meld::use_stream(stream);
  MELD_REQUIRE(gs && d && coeffs_host && S && R, "cheby_filter_dist: NULL argument");
  MELD_REQUIRE(d->connected, "cheby_filter_dist: the peer buffers are not connected (meld_b200_dist_connect)");
  MELD_REQUIRE(n_coeffs >= 2, "cheby_filter_dist: need at least 2 coefficients (got %d)", n_coeffs);
  MELD_REQUIRE(p >= 1 && p <= d->p_max, "cheby_filter_dist: p=%d outside 1..%d", p, d->p_max);
  MELD_REQUIRE(lmax > 0.0 && isfinite(lmax), "cheby_filter_dist: lmax=%g", lmax);
  MELD_REQUIRE(gs->n_cols == d->n, "cheby_filter_dist: graph has %lld columns, context %lld", (long long)gs->n_cols,
               (long long)d->n);
  MELD_REQUIRE(S != R, "cheby_filter_dist: R may not alias S");
  const int64_t n = gs->n_cols, nloc = gs->n_rows, row0 = gs->row0;
  const int pw = padded_width(p);  // <= 8 <= the buffers' width
  MELD_REQUIRE(pw <= d->p_max || pw == p, "cheby_filter_dist: padded width %d exceeds the context's %d", pw, d->p_max);
  const size_t lenl = ((size_t)nloc * pw + 2 + 3) & ~(size_t)3;
  MELD_CHECK(ensure_work(gs, lenl));
  double *Rloc = gs->work.p;
  const int pgrid = grid_for(n * pw, 256);
  // T_0 = S in graph order, every rank the whole signal (replicated, 8 n p bytes).  The peers finished storing
  // into this rank's buffers before the previous call returned here (its last phase waited for them).
  int ci = 0, oi = 1;
  permute_rows_kernel<<<pgrid, 256, 0, stream>>>(S, gs->perm.p, n, p, pw, /*scatter=*/0, d->buf(d->rank, ci));
  MELD_LAUNCH_CHECK();
  const double a1 = lmax / 2.0, a2 = lmax / 2.0;
  const int m = n_coeffs - 1;
  for (int k = 1; k <= m; ++k) {
    StepArgs a{};
    a.Tcur = d->buf(d->rank, ci);
    a.Told = k >= 2 ? d->buf(d->rank, oi) + (size_t)row0 * pw : nullptr;
    a.Tnew = d->buf(d->rank, oi) + (size_t)row0 * pw;  // T_k over T_{k-2}; at k = m the finished rows of R
    fill_peer_args(a, d, oi, row0, pw);
    a.halo = gs->halo.p;
    a.store_r = k == m;
    a.R = Rloc;
    a.alpha = (k == 1 ? 1.0 : 2.0) / a1;
    a.shift = a2;
    a.gamma = k >= 2 ? 1.0 : 0.0;
    a.c = coeffs_host[k];
    a.c_cur = k == 1 ? 0.5 * coeffs_host[0] : 0.0;
    a.r_acc = k >= 2;
    a.wait_epoch = d->epoch;  // the peers' stores of term k-1 (or their release of the buffers) are visible
    a.post_epoch = ++d->epoch;
    if (nloc > 0) {
      MELD_CHECK(launch_step(gs, a, pw, stream));
    } else {  // a rank without rows still takes part in every phase
      dist_unpermute_kernel<<<1, 256, 0, stream>>>(a, nullptr, nullptr, 0, p, p, nullptr);
      MELD_LAUNCH_CHECK();
    }
    const int t = ci;
    ci = oi;
    oi = t;
  }
  StepArgs a{};
  fill_peer_args(a, d, 0, 0, p);
  a.wait_epoch = d->epoch;
  a.post_epoch = ++d->epoch;
  dist_unpermute_kernel<<<pgrid, 256, 0, stream>>>(a, d->buf(d->rank, ci), gs->perm.p, n, p, pw, R);
  MELD_LAUNCH_CHECK();
  return 0;
}

int meld_b200_cheby_sweep(meld_b200_graph_t *g, double lmax, const double *coeffs_host, int n_filters, int n_coeffs,
                          const double *S, int p, double *R, void *stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  meld::use_stream(stream);
  MELD_REQUIRE(g && coeffs_host && S && R, "cheby_sweep: NULL argument");
  MELD_REQUIRE(n_filters >= 1 && n_filters <= 65536, "cheby_sweep: n_filters=%d", n_filters);
  MELD_REQUIRE(n_coeffs >= 2 && n_coeffs <= 1024, "cheby_sweep: n_coeffs=%d outside 2..1024", n_coeffs);
  MELD_REQUIRE(p >= 1 && p <= 8, "cheby_sweep: p=%d outside 1..8 (split the signal into column chunks)", p);
  MELD_REQUIRE(lmax > 0.0 && isfinite(lmax), "cheby_sweep: lmax=%g", lmax);
  MELD_REQUIRE(g->row0 == 0 && g->n_rows == g->n_cols, "cheby_sweep: needs the full operator");
  const int64_t n = g->n_rows;
  // the whole basis T_0 .. T_m is kept (n_coeffs slots of the padded signal) + the coefficient table
  const int pw = padded_width(p);
  const size_t len = ((size_t)n * pw + 2 + 3) & ~(size_t)3;
  const size_t ctab = ((size_t)n_filters * n_coeffs + 3) & ~(size_t)3;
  MELD_CHECK(ensure_work(g, (size_t)n_coeffs * len + ctab));
  double *T = g->work.p, *Cd = T + (size_t)n_coeffs * len;
  {
    std::vector<double> ch((size_t)n_filters * n_coeffs);
    for (int f = 0; f < n_filters; ++f) {
      for (int k = 0; k < n_coeffs; ++k) ch[(size_t)f * n_coeffs + k] = coeffs_host[(size_t)f * n_coeffs + k];
      ch[(size_t)f * n_coeffs] *= 0.5;  // PyGSP uses c_0 / 2
    }
    MELD_CUDA(cudaMemcpyAsync(Cd, ch.data(), ch.size() * sizeof(double), cudaMemcpyHostToDevice, stream));
    MELD_SYNC(stream);  // ch is a local
  }
  const int pgrid = grid_for(n * pw, 256);
  permute_rows_kernel<<<pgrid, 256, 0, stream>>>(S, g->perm.p, n, p, pw, /*scatter=*/0, T);  // T_0 = S in graph order
  MELD_LAUNCH_CHECK();
  const double a1 = lmax / 2.0, a2 = lmax / 2.0;
  for (int k = 1; k < n_coeffs; ++k) {
    StepArgs a{};
    a.Tcur = T + (size_t)(k - 1) * len;
    a.Told = k >= 2 ? T + (size_t)(k - 2) * len : nullptr;
    a.Tnew = T + (size_t)k * len;
    a.R = nullptr;
    a.alpha = (k == 1 ? 1.0 : 2.0) / a1;
    a.shift = a2;
    a.gamma = k >= 2 ? 1.0 : 0.0;
    MELD_CHECK(launch_step(g, a, pw, stream));
  }
  constexpr int FC = 16;
  const size_t smem = (size_t)FC * n_coeffs * sizeof(double);
  MELD_CUDA(cudaFuncSetAttribute(combine_basis_kernel<FC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  for (int f0 = 0; f0 < n_filters; f0 += FC) {
    const int nf = n_filters - f0 < FC ? n_filters - f0 : FC;
    combine_basis_kernel<FC><<<pgrid, 256, smem, stream>>>(T, len, n_coeffs, Cd, f0, nf, g->perm.p, n, p, pw, R);
    MELD_LAUNCH_CHECK();
  }
  return 0;
}

static int lanczos_check(std::vector<double> &alpha, std::vector<double> &beta, int &k, double rel_tol, double &theta);

int meld_b200_estimate_lmax(meld_b200_graph_t *g, int max_iters, double rel_tol, void *stream_, double *lmax_host,
                            int *iters_host) {
  cudaStream_t stream = (cudaStream_t)stream_;
  meld::use_stream(stream);
  MELD_REQUIRE(g && lmax_host, "estimate_lmax: NULL argument");
  MELD_REQUIRE(g->row0 == 0 && g->n_rows == g->n_cols, "estimate_lmax: needs the full operator");
  const int64_t n = g->n_rows;
  if (max_iters <= 0) max_iters = 160;
  if (max_iters > n) max_iters = (int)n;
  // Stopping rule: a bound of the relative error of the largest Ritz value theta.  With the Ritz residual
  // r = |beta_k s_last| there is an eigenvalue within r of theta; once r is below half the gap to the second Ritz
  // value the Kato-Temple bound r^2 / gap applies (measured on config 4: r / theta = 1.1e-3 at 28 steps where the
  // actual error is 8e-6; waiting for r itself to reach 1e-5 costs 44 steps instead of 32).  Stop when
  // min(r, r^2 / gap) <= rel_tol * theta.  The reference asks ARPACK for a RESIDUAL tolerance of 5e-3 (observed
  // eigenvalue error up to 2.9e-3, tests/test_gpu_scale.py); 1e-5 keeps the densities within ~2e-6 of those of
  // the converged lmax (SURVEY finding 6).
  if (rel_tol <= 0) rel_tol = 1e-5;
  MELD_REQUIRE(n >= 1, "estimate_lmax: empty graph");
  const size_t len = ((size_t)n + 2 + 3) & ~(size_t)3;  // padded like the filter's work arrays
  const size_t need = 3 * len + 4 * kRedBlocks + 2 * ((size_t)max_iters + 2);
  MELD_CHECK(ensure_work(g, need));
  double *wa = g->work.p, *wb = wa + len, *y = wb + len;
  double *pa = y + len, *pb = pa + kRedBlocks;  // pb: three rotating sets of |w_j|^2 partials
  double *d_alpha = pb + 3 * kRedBlocks, *d_beta = d_alpha + max_iters + 2;
  MELD_CUDA(cudaMemsetAsync(wb, 0, (size_t)n * sizeof(double), stream));
  MELD_CUDA(cudaMemsetAsync(pa, 0, (4 * kRedBlocks + 2 * ((size_t)max_iters + 2)) * sizeof(double), stream));
  lanczos_init_kernel<<<kRedBlocks, kRedThreads, 0, stream>>>(wa, n, pb);  // w_0, |w_0|^2 -> pb[0]
  MELD_LAUNCH_CHECK();
  std::vector<double> alpha((size_t)max_iters + 2), beta((size_t)max_iters + 2);
  double theta = 0.0;
  int k = 0, next_check = max_iters < 8 ? max_iters : 8;
  bool done = false;
  double *w_cur = wa, *w_prev = wb;
  while (!done && k < max_iters) {
    for (int j = k; j < next_check; ++j) {
      StepArgs a{};  // y = L w_j, partial sums of w_j . y
      a.Tcur = w_cur;
      a.Tnew = y;
      a.alpha = 1.0;
      a.dot_partials = pa;
      MELD_CHECK(launch_step(g, a, 1, stream));
      lanczos_axpy_kernel<<<kRedBlocks, kRedThreads, 0, stream>>>(y, w_cur, w_prev, n, pa, pb + (j % 3) * kRedBlocks,
                                                                  pb + ((j + 2) % 3) * kRedBlocks, j, d_alpha, d_beta,
                                                                  pb + ((j + 1) % 3) * kRedBlocks);
      MELD_LAUNCH_CHECK();
      double *t = w_cur;  // w_{j+1} was written over w_{j-1}
      w_cur = w_prev;
      w_prev = t;
    }
    k = next_check;
    lanczos_tail_kernel<<<1, kRedThreads, 0, stream>>>(pb + (k % 3) * kRedBlocks, k, d_beta);
    MELD_LAUNCH_CHECK();
    MELD_CUDA(cudaMemcpyAsync(alpha.data(), d_alpha, (size_t)k * sizeof(double), cudaMemcpyDeviceToHost, stream));
    MELD_CUDA(cudaMemcpyAsync(beta.data(), d_beta, (size_t)(k + 1) * sizeof(double), cudaMemcpyDeviceToHost, stream));
    MELD_SYNC(stream);
    const int rc = lanczos_check(alpha, beta, k, rel_tol, theta);
    if (rc < 0) return rc;
    done = rc == 1;
    next_check = k + 4 < max_iters ? k + 4 : max_iters;
  }
  *lmax_host = 1.01 * theta;
  if (iters_host) *iters_host = k;
  return 0;
}

// Shared by both Lanczos drivers: Ritz analysis of T_k (alpha_0..alpha_{k-1}, beta_1..beta_{k-1}; beta_k couples to
// the next vector).  Returns < 0 on error, 1 when converged (k may shrink to an invariant subspace), else 0.
static int lanczos_check(std::vector<double> &alpha, std::vector<double> &beta, int &k, double rel_tol, double &theta) {
  int kk = k;
  bool done = false;
  double scale = 0.0;
  for (int j = 0; j < k; ++j) scale = fmax(scale, fabs(alpha[(size_t)j]));
  for (int j = 1; j <= k; ++j) {
    if (!(beta[(size_t)j] > 1e-13 * fmax(scale, 1e-300))) {
      kk = j;
      done = true;
      break;
    }
  }
  for (int j = 0; j < kk; ++j)
    if (!isfinite(alpha[(size_t)j])) {
      set_error("estimate_lmax: non-finite Lanczos coefficient at step %d", j);
      return MELD_B200_ERR_INTERNAL;
    }
  std::vector<double> dd(alpha.begin(), alpha.begin() + kk), ee((size_t)kk, 0.0), zl((size_t)kk, 0.0);
  for (int j = 0; j + 1 < kk; ++j) ee[(size_t)j] = beta[(size_t)j + 1];
  if (!tridiag_ql_last_row(dd, ee, zl, kk)) {
    set_error("estimate_lmax: tridiagonal QL iteration did not converge");
    return MELD_B200_ERR_INTERNAL;
  }
  int top = 0;
  for (int j = 1; j < kk; ++j)
    if (dd[(size_t)j] > dd[(size_t)top]) top = j;
  theta = dd[(size_t)top];
  const double resid = kk < k ? 0.0 : fabs(beta[(size_t)k] * zl[(size_t)top]);
  double second = -INFINITY;
  for (int j = 0; j < kk; ++j)
    if (j != top && dd[(size_t)j] > second) second = dd[(size_t)j];
  const double gap = theta - second;  // Ritz estimate of the spectral gap below lambda_max
  double err = resid;
  if (kk >= 2 && gap > 0.0 && resid <= 0.5 * gap) err = fmin(resid, resid * resid / gap);
  if (getenv("MELD_B200_LANCZOS_DEBUG"))
    fprintf(stderr, "[meld_b200 lanczos] k=%d theta=%.15g resid/theta=%.3e ritz gap/theta=%.3e bound/theta=%.3e\n", k,
            theta, resid / fabs(theta), gap / fabs(theta), err / fabs(theta));
  if (err <= rel_tol * fabs(theta)) done = true;
  if (done) k = kk;
  return done ? 1 : 0;
}

int meld_b200_estimate_lmax_dist(meld_b200_graph_t *gs, meld_b200_dist_t *d, int max_iters, double rel_tol,
                                 void *stream_, double *lmax_host, int *iters_host) {
  cudaStream_t stream = (cudaStream_t)stream_;
  meld::use_stream(stream);
  MELD_REQUIRE(gs && d && lmax_host, "estimate_lmax_dist: NULL argument");
  MELD_REQUIRE(d->connected, "estimate_lmax_dist: the peer buffers are not connected");
  MELD_REQUIRE(gs->n_cols == d->n, "estimate_lmax_dist: graph has %lld columns, context %lld", (long long)gs->n_cols,
               (long long)d->n);
  const int64_t n = gs->n_cols, nloc = gs->n_rows, row0 = gs->row0;
  if (max_iters <= 0) max_iters = 160;
  if (max_iters > n) max_iters = (int)n;
  if (rel_tol <= 0) rel_tol = 1e-5;
  const size_t lenl = ((size_t)nloc + 2 + 3) & ~(size_t)3;
  const size_t need = lenl + 3 * kRedBlocks + 2 * ((size_t)max_iters + 2);
  MELD_CHECK(ensure_work(gs, need));
  double *y = gs->work.p, *pa = y + lenl, *pb0 = pa + kRedBlocks, *pbn = pb0 + kRedBlocks;
  double *d_alpha = pbn + kRedBlocks, *d_beta = d_alpha + max_iters + 2;
  MELD_CUDA(cudaMemsetAsync(pa, 0, (3 * kRedBlocks + 2 * ((size_t)max_iters + 2)) * sizeof(double), stream));
  int ci = 0, oi = 1;
  MELD_CUDA(cudaMemsetAsync(d->buf(d->rank, oi), 0, (size_t)n * sizeof(double), stream));  // w_{-1} = 0
  lanczos_init_kernel<<<kRedBlocks, kRedThreads, 0, stream>>>(d->buf(d->rank, ci), n, pb0);  // every rank: whole w_0
  MELD_LAUNCH_CHECK();
  std::vector<double> alpha((size_t)max_iters + 2), beta((size_t)max_iters + 2);
  double theta = 0.0;
  int k = 0, next_check = max_iters < 8 ? max_iters : 8;
  bool done = false;
  unsigned long long norm_epoch = 0;
  while (!done && k < max_iters) {
    for (int j = k; j < next_check; ++j) {
      StepArgs a{};  // phase A: y = L[rows] w_j, this rank's part of w_j . y published
      a.Tcur = d->buf(d->rank, ci);
      a.Tnew = y;
      a.alpha = 1.0;
      a.dot_partials = pa;
      fill_peer_args(a, d, oi, row0, 1);
      a.peer_store = 0;
      a.scal_partials = pa;
      a.wait_epoch = d->epoch;  // the peers' rows of w_j have arrived
      a.post_epoch = ++d->epoch;
      const unsigned long long dot_epoch = a.post_epoch;
      if (nloc > 0) {
        MELD_CHECK(launch_step(gs, a, 1, stream));
      } else {
        a.n_scal_partials = 1;  // pa[0] = 0: an empty rank publishes a zero
        dist_unpermute_kernel<<<1, 256, 0, stream>>>(a, nullptr, nullptr, 0, 1, 1, nullptr);
        MELD_LAUNCH_CHECK();
      }
      StepArgs b{};  // phase B: w_{j+1} rows into every rank's vector, |w_{j+1}|^2 of these rows published
      fill_peer_args(b, d, oi, row0, 1);
      b.halo = gs->halo.p;
      b.scal_partials = pbn;
      b.n_scal_partials = kRedBlocks;
      b.wait_epoch = d->epoch;
      b.post_epoch = ++d->epoch;
      // few blocks for few rows (unwritten partial slots stay zero): a waiting grid should not cover the whole GPU
      int agrid = (int)ceil_div(nloc > 0 ? nloc : 1, 4 * kRedThreads);
      if (agrid > kRedBlocks) agrid = kRedBlocks;
      lanczos_axpy_peer_kernel<<<agrid, kRedThreads, 0, stream>>>(b, y, d->buf(d->rank, ci), d->buf(d->rank, oi), row0,
                                                                       nloc, dot_epoch, norm_epoch, pb0, j, d_alpha,
                                                                       d_beta, pbn);
      MELD_LAUNCH_CHECK();
      norm_epoch = b.post_epoch;
      const int t = ci;
      ci = oi;
      oi = t;
    }
    k = next_check;
    StepArgs t{};
    fill_peer_args(t, d, oi, row0, 1);
    t.wait_epoch = d->epoch;
    lanczos_tail_peer_kernel<<<1, 32, 0, stream>>>(t, norm_epoch, k, d_beta);
    MELD_LAUNCH_CHECK();
    MELD_CUDA(cudaMemcpyAsync(alpha.data(), d_alpha, (size_t)k * sizeof(double), cudaMemcpyDeviceToHost, stream));
    MELD_CUDA(cudaMemcpyAsync(beta.data(), d_beta, (size_t)(k + 1) * sizeof(double), cudaMemcpyDeviceToHost, stream));
    MELD_SYNC(stream);
    const int rc = lanczos_check(alpha, beta, k, rel_tol, theta);  // identical on every rank: same sums, same order
    if (rc < 0) return rc;
    done = rc == 1;
    next_check = k + 4 < max_iters ? k + 4 : max_iters;
  }
  *lmax_host = 1.01 * theta;
  if (iters_host) *iters_host = k;
  return 0;
}

int meld_b200_indicator_matrix(const int32_t *codes, int64_t n, int p, int sample_normalize, double *S,
                               void *stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  meld::use_stream(stream);
  MELD_REQUIRE(codes && S && n > 0 && p > 0, "indicator_matrix: bad argument");
  DevBuf<unsigned long long> cnt;
  MELD_CHECK(cnt.alloc((size_t)p));
  MELD_CUDA(cudaMemsetAsync(cnt.p, 0, (size_t)p * sizeof(unsigned long long), stream));
  count_codes_kernel<<<grid_for(n, 256), 256, 0, stream>>>(codes, n, p, cnt.p);
  MELD_LAUNCH_CHECK();
  fill_indicator_kernel<<<grid_for(n * p, 256), 256, 0, stream>>>(codes, n, p, sample_normalize, cnt.p, S);
  MELD_LAUNCH_CHECK();
  // cnt goes back to the pool in stream order (cudaFreeAsync on this stream); only plain cudaFree needs the sync
  if (!use_pool()) MELD_SYNC(stream);
  return 0;
}

int meld_b200_l1_normalize_rows(const double *in, int64_t n, int p, double *out, void *stream_) {
  meld::use_stream((cudaStream_t)stream_);
  MELD_REQUIRE(in && out && n >= 0 && p > 0, "l1_normalize_rows: bad argument");
  if (n == 0) return 0;
  l1_normalize_rows_kernel<<<grid_for(n, 256), 256, 0, (cudaStream_t)stream_>>>(in, n, p, out);
  MELD_LAUNCH_CHECK();
  return 0;
}

}  // extern "C"
