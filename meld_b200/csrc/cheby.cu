// Chebyshev three-term recurrence on the CSR Laplacian (the HBM-roofline kernel),
// the Lanczos lmax estimate that reuses it, and the tiny signal helpers.
//
// Reference behaviour restated (upstream PyGSP 0.5.1, reached from meld/filter.py:39-59):
//   cheby_op: T0 = S; T1 = (L S - a2 S)/a1; R = c0/2 T0 + c1 T1;
//             T_k = (2/a1)(L - a2 I) T_{k-1} - T_{k-2}; R += c_k T_k      (a1 = a2 = lmax/2)
//   estimate_lmax: 1.01 * largest eigenvalue of L.
//
// Kernel design (one launch per recurrence term):
//   * persistent CTAs; each walks row blocks b = blockIdx.x, += gridDim.x;
//   * a row block is a contiguous run of ~blk_chunk CSR entries; its columns and values
//     are staged into shared memory with two 1-D TMA bulk copies (cp.async.bulk +
//     mbarrier complete_tx), n_stage blocks deep, so HBM streaming of the matrix is
//     decoupled from the L2 gathers of T_{k-1};
//   * G lanes cooperate on a row: coalesced shared-memory reads of (col, val),
//     128-bit gathers of the P-wide signal row, warp-shuffle reduction over the G lanes;
//   * the three-term update and the R accumulation are fused into the epilogue, T_k is
//     written over T_{k-2} (row i of T_{k-2} is only ever read by row i).
#include "common.cuh"

#include <math.h>
#include <vector>

namespace meld {

// ---- PTX helpers: mbarrier + 1-D TMA bulk copy ---------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_barrier_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
  uint32_t addr = smem_u32(bar), ok;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(addr), "r"(parity)
        : "memory");
  } while (!ok);
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

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
  int cap;      // CSR entries per stage
  int rcap;     // row pointers per stage
  int n_stage;  // pipeline depth
};

// Gather one P-wide signal row with the widest aligned loads available: every distinct
// 128-byte line touched by a warp-level load costs ~2 L1 wavefront cycles, so a 32-byte row
// must be one 256-bit request, not two 128-bit ones.
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

constexpr int kUnroll = 4;  // gathers in flight per lane

// Rows [r0, r1) of one block.  cs/vs are indexed by absolute CSR entry and rp by absolute
// row (both already offset), whether they point into shared memory or global memory.
template <int P, int G>
__device__ __forceinline__ void process_rows(const StepArgs &a, int r0, int r1, const int32_t *cs, const double *vs,
                                             const int32_t *rp, int gid, int gl, int ngroups) {
  for (int rb = r0; rb < r1; rb += ngroups) {  // CTA-uniform trip count (full-mask shuffles below)
    const int r = rb + gid;
    const bool act = r < r1;
    int eb = 0, ee = 0;
    if (act) {
      eb = rp[r];
      ee = rp[r + 1];
    }
    // Epilogue operands are requested first so their DRAM latency overlaps the gathers.
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
    for (int e = eb + gl; e < ee; e += kUnroll * G) {
      int32_t c[kUnroll];
      double v[kUnroll];
      double x[kUnroll][P];
#pragma unroll
      for (int u = 0; u < kUnroll; ++u) {
        const bool in = e + u * G < ee;
        c[u] = in ? cs[e + u * G] : 0;
        v[u] = in ? vs[e + u * G] : 0.0;
      }
#pragma unroll
      for (int u = 0; u < kUnroll; ++u) {
        if (e + u * G < ee) {
          gather_row<P>(a.Tcur, c[u], x[u]);
        } else {
#pragma unroll
          for (int k = 0; k < P; ++k) x[u][k] = 0.0;
        }
      }
#pragma unroll
      for (int u = 0; u < kUnroll; ++u) {
#pragma unroll
        for (int k = 0; k < P; ++k) acc[k] = fma(v[u], x[u][k], acc[k]);
      }
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
      if (a.Tnew) a.Tnew[li] = tn;  // re-read by the next step's gathers: keep cacheable
      if (a.R) {
        double rv = a.c * tn + a.c_cur * tc;
        if (a.r_acc) rv += rold;
        st_stream(a.R + li, rv);
      }
    }
  }
}

template <int P, int G>
__global__ void cheby_step_kernel(const StepArgs a) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int cap = a.cap, rcap = a.rcap, ns = a.n_stage;
  double *sval = reinterpret_cast<double *>(smem_raw);
  int32_t *scol = reinterpret_cast<int32_t *>(smem_raw + (size_t)ns * cap * 8);
  int32_t *srp = reinterpret_cast<int32_t *>(smem_raw + (size_t)ns * cap * 12);
  uint64_t *bars = reinterpret_cast<uint64_t *>(smem_raw + (size_t)ns * cap * 12 + (size_t)ns * rcap * 4);

  const int tid = threadIdx.x;
  const int ngroups = blockDim.x / G;
  const int gid = tid / G, gl = tid % G;
  const int stride = gridDim.x;

  if (tid == 0) {
    for (int s = 0; s < ns; ++s) mbar_init(&bars[s], 1);
    fence_barrier_init();
  }
  __syncthreads();

  // Thread 0 is the TMA producer: arm the stage barrier, then three bulk copies
  // (values, columns, row pointers), each 16-byte aligned at both ends.
  auto issue = [&](int b, int s) {
    const int r0 = __ldg(a.blk + b), r1 = __ldg(a.blk + b + 1);
    const int e0 = __ldg(a.row_ptr + r0), e1 = __ldg(a.row_ptr + r1);
    const int a0 = e0 & ~3;                  // aligned start for the int32 columns
    const int n = ((e1 - a0) + 3) & ~3;      // multiple of 4 entries (tail lands in kCsrPad)
    const int ra = r0 & ~3;
    const int nr = ((r1 + 1 - ra) + 3) & ~3;
    if (n > 0 && n <= cap && nr <= rcap) {
      mbar_arrive_expect_tx(&bars[s], (uint32_t)n * 12u + (uint32_t)nr * 4u);
      bulk_g2s(sval + (size_t)s * cap, a.val + a0, (uint32_t)n * 8u, &bars[s]);
      bulk_g2s(scol + (size_t)s * cap, a.col + a0, (uint32_t)n * 4u, &bars[s]);
      bulk_g2s(srp + (size_t)s * rcap, a.row_ptr + ra, (uint32_t)nr * 4u, &bars[s]);
    } else {
      mbar_arrive(&bars[s]);  // oversize / empty block: nothing staged, phase still advances
    }
  };

  if (tid == 0) {
    for (int s = 0; s < ns; ++s) {
      const int b = blockIdx.x + s * stride;
      if (b < a.n_blk) issue(b, s);
    }
  }

  int it = 0;
  for (int b = blockIdx.x; b < a.n_blk; b += stride, ++it) {
    const int s = it % ns;
    const uint32_t parity = (uint32_t)(it / ns) & 1u;
    const int r0 = __ldg(a.blk + b), r1 = __ldg(a.blk + b + 1);
    const int e0 = __ldg(a.row_ptr + r0), e1 = __ldg(a.row_ptr + r1);
    const int a0 = e0 & ~3;
    const int n = ((e1 - a0) + 3) & ~3;
    const int ra = r0 & ~3;
    const int nr = ((r1 + 1 - ra) + 3) & ~3;
    mbar_wait(&bars[s], parity);
    if (n <= cap && nr <= rcap) {
      process_rows<P, G>(a, r0, r1, scol + (size_t)s * cap - a0, sval + (size_t)s * cap - a0,
                         srp + (size_t)s * rcap - ra, gid, gl, ngroups);
    } else {
      process_rows<P, G>(a, r0, r1, a.col, a.val, a.row_ptr, gid, gl, ngroups);  // block too long for a stage
    }
    __syncthreads();  // every lane is done reading stage s before it is refilled
    if (tid == 0) {
      const int nb = b + ns * stride;
      if (nb < a.n_blk) issue(nb, s);
    }
  }
}

typedef void (*StepKernel)(const StepArgs);

template <int P>
static StepKernel pick_group(int G) {
  if constexpr (P <= 4) {
    if (G == 4) return cheby_step_kernel<P, 4>;
  }
  switch (G) {
    case 4:
    case 8: return cheby_step_kernel<P, 8>;
    case 16: return cheby_step_kernel<P, 16>;
    default: return cheby_step_kernel<P, 32>;
  }
}

static StepKernel pick_kernel(int P, int G) {
  switch (P) {
    case 1: return pick_group<1>(G);
    case 2: return pick_group<2>(G);
    case 3: return pick_group<3>(G);
    case 4: return pick_group<4>(G);
    case 5: return pick_group<5>(G);
    case 6: return pick_group<6>(G);
    case 7: return pick_group<7>(G);
    case 8: return pick_group<8>(G);
    default: return nullptr;
  }
}

static int choose_group(const meld_b200_graph *g, int P) {
  int G = tuning().group;
  if (G != 4 && G != 8 && G != 16 && G != 32) {
    const double avg = g->n_rows > 0 ? (double)g->nnz / (double)g->n_rows : 0.0;
    G = avg <= 12.0 ? 4 : (avg <= 80.0 ? 8 : (avg <= 256.0 ? 16 : 32));
  }
  if (G < P) G = 8;
  return G;
}

static int launch_step(const meld_b200_graph *g, StepArgs a, int P, cudaStream_t stream) {
  const Tuning &t = tuning();
  const int G = choose_group(g, P);
  StepKernel k = pick_kernel(P, G);
  MELD_REQUIRE(k != nullptr, "cheby_step: p=%d outside 1..8", P);
  a.row_ptr = g->row_ptr.p;
  a.col = g->col.p;
  a.val = g->val.p;
  a.blk = g->blk.p;
  a.n_blk = g->n_blk;
  a.row0 = g->row0;
  a.cap = t.stage_cap;
  a.rcap = t.row_cap;
  a.n_stage = t.n_stage;
  const size_t smem = (size_t)t.n_stage * ((size_t)t.stage_cap * 12 + (size_t)t.row_cap * 4 + 8);
  MELD_REQUIRE(smem <= 227 * 1024, "cheby_step: %zu bytes of shared memory requested", smem);
  MELD_REQUIRE(t.threads % 32 == 0 && t.threads >= 32 && t.threads <= 1024 && t.stage_cap % 16 == 0 && t.row_cap % 16 == 0,
               "cheby_step: bad tuning");
  MELD_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int grid = sm_count() * t.ctas_per_sm;
  if (grid > g->n_blk) grid = g->n_blk;
  if (grid < 1) grid = 1;
  k<<<grid, t.threads, smem, stream>>>(a);
  MELD_LAUNCH_CHECK();
  return 0;
}

static int ensure_work(meld_b200_graph *g, size_t count) {
  if (g->work.n >= count) return 0;
  return g->work.alloc(count);
}

// ---- Lanczos helpers ----------------------------------------------------------------
constexpr int kRedBlocks = 256;  // partial sums per reduction (fixed => deterministic)
constexpr int kRedThreads = 256;

__device__ __forceinline__ double block_sum(double v, double *sh) {
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

__global__ void lanczos_scale_kernel(double *v, int64_t n, const double *partials) {
  __shared__ double sh[33];
  const double inv = 1.0 / sqrt(sum_partials(partials, sh));
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    v[i] *= inv;
}

__global__ void dot_partials_kernel(const double *__restrict__ x, const double *__restrict__ y, int64_t n,
                                    double *partials) {
  __shared__ double sh[33];
  double s = 0.0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    s = fma(x[i], y[i], s);
  s = block_sum(s, sh);
  if (threadIdx.x == 0) partials[blockIdx.x] = s;
}

// w -= alpha v + beta_j v_prev ; alpha = sum(partials_a) ; partials_b = ||w||^2 pieces
__global__ void lanczos_update_kernel(double *w, const double *__restrict__ v, const double *__restrict__ vprev,
                                      int64_t n, const double *partials_a, const double *beta_arr, int j,
                                      double *alpha_arr, double *partials_b) {
  __shared__ double sh[33];
  const double alpha = sum_partials(partials_a, sh);
  const double beta = beta_arr[j];
  if (blockIdx.x == 0 && threadIdx.x == 0) alpha_arr[j] = alpha;
  double s = 0.0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const double x = w[i] - alpha * v[i] - beta * vprev[i];
    w[i] = x;
    s = fma(x, x, s);
  }
  s = block_sum(s, sh);
  if (threadIdx.x == 0) partials_b[blockIdx.x] = s;
}

// beta_{j+1} = ||w|| ; v_prev = v ; v = w / beta_{j+1}
__global__ void lanczos_normalize_kernel(const double *__restrict__ w, double *v, double *vprev, int64_t n,
                                         const double *partials_b, double *beta_arr, int j) {
  __shared__ double sh[33];
  const double beta = sqrt(sum_partials(partials_b, sh));
  if (blockIdx.x == 0 && threadIdx.x == 0) beta_arr[j + 1] = beta;
  const double inv = beta > 0.0 ? 1.0 / beta : 0.0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    vprev[i] = v[i];
    v[i] = w[i] * inv;
  }
}

// Largest eigenvalue of the symmetric tridiagonal (alpha[0..k), beta[1..k)) by Sturm bisection.
static double tridiag_lmax(const std::vector<double> &alpha, const std::vector<double> &beta, int k) {
  double lo = alpha[0], hi = alpha[0];
  for (int i = 0; i < k; ++i) {
    const double bl = i > 0 ? fabs(beta[i]) : 0.0, br = i + 1 < k ? fabs(beta[i + 1]) : 0.0;
    lo = fmin(lo, alpha[i] - bl - br);
    hi = fmax(hi, alpha[i] + bl + br);
  }
  auto count_below = [&](double x) {  // number of eigenvalues < x
    int cnt = 0;
    double q = alpha[0] - x;
    if (q < 0) ++cnt;
    for (int i = 1; i < k; ++i) {
      if (q == 0.0) q = 1e-300;
      q = alpha[i] - x - beta[i] * beta[i] / q;
      if (q < 0) ++cnt;
    }
    return cnt;
  };
  for (int itn = 0; itn < 200 && hi - lo > 1e-15 * fmax(fabs(hi), fabs(lo)); ++itn) {
    const double mid = 0.5 * (lo + hi);
    if (count_below(mid) >= k)
      hi = mid;  // all eigenvalues below mid
    else
      lo = mid;
  }
  return 0.5 * (lo + hi);
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
  MELD_REQUIRE(g && T_cur, "cheby_step: NULL argument");
  MELD_REQUIRE(p >= 1 && p <= 8, "cheby_step: p=%d outside 1..8", p);
  MELD_REQUIRE(gamma == 0.0 || T_old != nullptr, "cheby_step: gamma != 0 needs T_old");
  MELD_REQUIRE((const double *)T_new != T_cur, "cheby_step: T_new may not alias T_cur");
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

int meld_b200_cheby_filter(meld_b200_graph_t *g, double lmax, const double *coeffs_host, int n_coeffs, const double *S,
                           int p, double *R, void *stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  MELD_REQUIRE(g && coeffs_host && S && R, "cheby_filter: NULL argument");
  MELD_REQUIRE(n_coeffs >= 2, "cheby_filter: need at least 2 coefficients (got %d)", n_coeffs);
  MELD_REQUIRE(p >= 1 && p <= 8, "cheby_filter: p=%d outside 1..8 (split the signal into column chunks)", p);
  MELD_REQUIRE(lmax > 0.0 && isfinite(lmax), "cheby_filter: lmax=%g", lmax);
  MELD_REQUIRE(g->row0 == 0 && g->n_rows == g->n_cols, "cheby_filter: needs the full operator (use cheby_step)");
  MELD_REQUIRE(S != R, "cheby_filter: R may not alias S");
  const size_t np = (size_t)g->n_rows * p;
  MELD_CHECK(ensure_work(g, 2 * np));
  double *Ta = g->work.p, *Tb = g->work.p + np;
  const double a1 = lmax / 2.0, a2 = lmax / 2.0;
  // k = 1: T1 = (L S - a2 S)/a1 ; R = c0/2 S + c1 T1
  StepArgs a{};
  a.Tcur = S;
  a.Told = nullptr;
  a.Tnew = (n_coeffs > 2) ? Ta : nullptr;
  a.R = R;
  a.alpha = 1.0 / a1;
  a.shift = a2;
  a.gamma = 0.0;
  a.c = coeffs_host[1];
  a.c_cur = 0.5 * coeffs_host[0];
  a.r_acc = 0;
  MELD_CHECK(launch_step(g, a, p, stream));
  // k = 2 reads T0 = S (caller-owned) so T2 goes to the second buffer; from k = 3 on
  // T_k overwrites T_{k-2} and the two workspace buffers ping-pong.
  const double *cur = Ta, *old = S;
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
    MELD_CHECK(launch_step(g, a, p, stream));
    old = cur;
    cur = nxt;
  }
  return 0;
}

int meld_b200_estimate_lmax(meld_b200_graph_t *g, int max_iters, double rel_tol, void *stream_, double *lmax_host,
                            int *iters_host) {
  cudaStream_t stream = (cudaStream_t)stream_;
  MELD_REQUIRE(g && lmax_host, "estimate_lmax: NULL argument");
  MELD_REQUIRE(g->row0 == 0 && g->n_rows == g->n_cols, "estimate_lmax: needs the full operator");
  const int64_t n = g->n_rows;
  if (max_iters <= 0) max_iters = 160;
  if (max_iters > n) max_iters = (int)n;
  if (rel_tol <= 0) rel_tol = 1e-9;
  MELD_REQUIRE(n >= 1, "estimate_lmax: empty graph");
  const size_t need = 3 * (size_t)n + 2 * kRedBlocks + 2 * ((size_t)max_iters + 2);
  MELD_CHECK(ensure_work(g, need));
  double *v = g->work.p, *vprev = v + n, *w = vprev + n;
  double *pa = w + n, *pb = pa + kRedBlocks;
  double *d_alpha = pb + kRedBlocks, *d_beta = d_alpha + max_iters + 2;
  MELD_CUDA(cudaMemsetAsync(vprev, 0, (size_t)n * sizeof(double), stream));
  MELD_CUDA(cudaMemsetAsync(pa, 0, (2 * kRedBlocks + 2 * ((size_t)max_iters + 2)) * sizeof(double), stream));
  lanczos_init_kernel<<<kRedBlocks, kRedThreads, 0, stream>>>(v, n, pa);
  MELD_LAUNCH_CHECK();
  lanczos_scale_kernel<<<kRedBlocks, kRedThreads, 0, stream>>>(v, n, pa);
  MELD_LAUNCH_CHECK();
  std::vector<double> alpha((size_t)max_iters + 2), beta((size_t)max_iters + 2);
  const int chunk = 8;
  double theta = 0.0, theta_prev = -1.0;
  int k = 0;
  bool done = false;
  while (!done && k < max_iters) {
    const int kend = (k + chunk < max_iters) ? k + chunk : max_iters;
    for (int j = k; j < kend; ++j) {
      StepArgs a{};  // w = L v
      a.Tcur = v;
      a.Tnew = w;
      a.alpha = 1.0;
      MELD_CHECK(launch_step(g, a, 1, stream));
      dot_partials_kernel<<<kRedBlocks, kRedThreads, 0, stream>>>(v, w, n, pa);
      MELD_LAUNCH_CHECK();
      lanczos_update_kernel<<<kRedBlocks, kRedThreads, 0, stream>>>(w, v, vprev, n, pa, d_beta, j, d_alpha, pb);
      MELD_LAUNCH_CHECK();
      lanczos_normalize_kernel<<<kRedBlocks, kRedThreads, 0, stream>>>(w, v, vprev, n, pb, d_beta, j);
      MELD_LAUNCH_CHECK();
    }
    k = kend;
    MELD_CUDA(cudaMemcpyAsync(alpha.data(), d_alpha, (size_t)(k + 1) * sizeof(double), cudaMemcpyDeviceToHost, stream));
    MELD_CUDA(cudaMemcpyAsync(beta.data(), d_beta, (size_t)(k + 1) * sizeof(double), cudaMemcpyDeviceToHost, stream));
    MELD_CUDA(cudaStreamSynchronize(stream));
    // An exactly invariant Krylov space (beta_j ~ 0) ends the recurrence early.
    int kk = k;
    double scale = 0.0;
    for (int j = 0; j < k; ++j) scale = fmax(scale, fabs(alpha[j]));
    for (int j = 1; j <= k; ++j) {
      if (!(beta[j] > 1e-13 * fmax(scale, 1e-300))) {
        kk = j;
        done = true;
        break;
      }
    }
    if (kk > k) kk = k;
    for (int j = 0; j < kk; ++j)
      if (!isfinite(alpha[j])) {
        set_error("estimate_lmax: non-finite Lanczos coefficient at step %d", j);
        return MELD_B200_ERR_INTERNAL;
      }
    theta = tridiag_lmax(alpha, beta, kk);
    if (theta_prev >= 0.0 && fabs(theta - theta_prev) <= rel_tol * fabs(theta)) done = true;
    theta_prev = theta;
    if (done) k = kk;
  }
  *lmax_host = 1.01 * theta;
  if (iters_host) *iters_host = k;
  return 0;
}

int meld_b200_indicator_matrix(const int32_t *codes, int64_t n, int p, int sample_normalize, double *S,
                               void *stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  MELD_REQUIRE(codes && S && n > 0 && p > 0, "indicator_matrix: bad argument");
  DevBuf<unsigned long long> cnt;
  MELD_CHECK(cnt.alloc((size_t)p));
  MELD_CUDA(cudaMemsetAsync(cnt.p, 0, (size_t)p * sizeof(unsigned long long), stream));
  count_codes_kernel<<<grid_for(n, 256), 256, 0, stream>>>(codes, n, p, cnt.p);
  MELD_LAUNCH_CHECK();
  fill_indicator_kernel<<<grid_for(n * p, 256), 256, 0, stream>>>(codes, n, p, sample_normalize, cnt.p, S);
  MELD_LAUNCH_CHECK();
  MELD_CUDA(cudaStreamSynchronize(stream));  // cnt is freed on return
  return 0;
}

int meld_b200_l1_normalize_rows(const double *in, int64_t n, int p, double *out, void *stream_) {
  MELD_REQUIRE(in && out && n >= 0 && p > 0, "l1_normalize_rows: bad argument");
  if (n == 0) return 0;
  l1_normalize_rows_kernel<<<grid_for(n, 256), 256, 0, (cudaStream_t)stream_>>>(in, n, p, out);
  MELD_LAUNCH_CHECK();
  return 0;
}

}  // extern "C"
