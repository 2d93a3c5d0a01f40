// kNN alpha-decay graph build on the device (SURVEY 8a rows C..F).
//
// What the reference computes (graphtools 1.5.x kNNGraph.build_kernel_to_data +
// BaseGraph.symmetrize_kernel / apply_anisotropy + PyGSPGraph + pygsp compute_laplacian,
// reached from MELD.fit, reference meld/meld.py:117-118,273):
//   eps_i = distance to the knn-th non-self neighbour (x bandwidth_scale, floored at eps)
//   K_ij  = exp(-(d_ij/eps_i)^decay) for every j with K_ij >= thresh   (K_ii = 1)
//   K     = (K + K^T)/2 ;  q = K 1 ;  K_ij /= (q_i q_j)^anisotropy
//   W     = K - diag(K) ;  L = diag(W 1) - W
//
// How it is computed here:
//   0. cells are re-ordered: k-means cluster first, Morton curve inside a cluster (cell_order), so that
//      256-cell tiles are compact and the search can skip tile pairs that are provably too far apart;
//   1. candidate search in reduced precision (tcgen05 bf16-split GEMM over the surviving tile pairs, or
//      the SIMT fp32 cross-check): pass 1 finds an upper bound of eps_i^2, pass 2 emits every j whose
//      approximate distance is inside the kernel radius plus a rigorous error margin, so
//      the candidate set is a superset of the exact neighbourhood;
//   2. exact float64 distances for the candidates only, summed in the same order as
//      scikit-learn's ball tree (sequential over features, no FMA) -> exact eps_i;
//   3. K_ij and K_ji both follow from d_ij (= d_ji bit for bit), eps_i and eps_j, so the
//      symmetrised value needs no transpose lookup; entries whose reverse edge is missing
//      are appended to the other row;
//   4. rows brought into column order by rank (merge_rows_kernel: own entries are already ascending, the
//      few mirrored ones are placed by counting), then anisotropy + Laplacian in place.
// Entry points: meld_b200_knn_graph_build (one GPU), meld_b200_dense_graph_build (thresh = 0: every pair is a
// candidate, test scale), meld_b200_knn_candidates + meld_b200_graph_from_candidates (stage 1 sharded by query rows,
// stage 2 replicated), meld_b200_stage2_begin / _records / _assemble / _finish (stage 2 row-partitioned as well: every
// rank assembles only its rows; the caller runs two 8 N-byte all-gathers and one all-to-all-v of mirror records).
// decay = 0 stands for the reference's decay=None (binary kNN kernel).
#include "common.cuh"
#include "knn_search.cuh"

#include <cub/device/device_scan.cuh>
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_segmented_sort.cuh>
#include <algorithm>
#include <vector>
#include <float.h>
#include <stdlib.h>
#include <math.h>

namespace meld {

// ---- 0. preparation: column means, centred norms -------------------------------------------
constexpr int kMeanBlocks = 592;  // 4 per SM; fixed, so the column sums do not depend on the device

__global__ void col_partial_sums_kernel(const double *__restrict__ X, int64_t n, int64_t d, double *partial) {
  // block b sums rows b, b + gridDim.x, ... ; thread t handles columns t, t + blockDim.x, ...
  for (int64_t k = threadIdx.x; k < d; k += blockDim.x) {
    double s = 0.0;
    for (int64_t i = blockIdx.x; i < n; i += gridDim.x) s += X[i * d + k];
    partial[(int64_t)blockIdx.x * d + k] = s;
  }
}

__global__ void col_mean_kernel(const double *__restrict__ partial, int64_t n, int64_t d, int nblocks, double *mu) {
  int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= d) return;
  double s = 0.0;
  for (int b = 0; b < nblocks; ++b) s += partial[(int64_t)b * d + k];
  mu[k] = s / (double)n;
}

// one warp per row: n_i = |x_i - mu|^2 (float64) and the running maximum over rows
__global__ void row_norms_kernel(const double *__restrict__ X, const double *__restrict__ mu, int64_t n, int64_t d,
                                 double *__restrict__ norm, unsigned long long *ymax2_bits) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t i = warp; i < n; i += nwarps) {
    double s = 0.0;
    for (int64_t k = lane; k < d; k += 32) {
      const double v = X[i * d + k] - mu[k];
      s = fma(v, v, s);
    }
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) {
      norm[i] = s;
      atomicMax(ymax2_bits, (unsigned long long)__double_as_longlong(s));  // s >= 0: bit order = value order
    }
  }
}

// ---- 0b. cell ordering along a Morton curve of the four highest-variance features ---------------
// Neighbouring rows of the kNN graph then share most of their columns, which is what the Chebyshev
// kernel's gathers (rows of T_k shared through L1 / L2 by neighbouring rows) feed on.
__global__ void col_partial_sq_kernel(const double *__restrict__ X, const double *__restrict__ mu, int64_t n, int64_t d,
                                      double *partial) {
  for (int64_t k = threadIdx.x; k < d; k += blockDim.x) {
    double s = 0.0;
    const double m = mu[k];
    for (int64_t i = blockIdx.x; i < n; i += gridDim.x) {
      const double v = X[i * d + k] - m;
      s = fma(v, v, s);
    }
    partial[(int64_t)blockIdx.x * d + k] = s;
  }
}

struct MortonDims {
  int dim[4];
  double mu[4];
  double inv_range[4];  // 1 / (12 sigma)
  int ndim;
};

__device__ __forceinline__ unsigned long long spread16(unsigned int v, int ndim) {
  unsigned long long r = 0;
  for (int b = 0; b < 16; ++b) r |= (unsigned long long)((v >> b) & 1u) << (b * ndim);
  return r;
}

__global__ void morton_keys_kernel(const double *__restrict__ X, int64_t n, int64_t d, MortonDims md,
                                   unsigned long long *__restrict__ keys, int32_t *__restrict__ idx) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  unsigned long long key = 0;
  for (int k = 0; k < md.ndim; ++k) {
    double q = (X[i * d + md.dim[k]] - md.mu[k]) * md.inv_range[k] + 0.5;
    q = fmin(fmax(q, 0.0), 1.0);
    const unsigned int v = (unsigned int)(q * 65535.0);
    key |= spread16(v, md.ndim) << k;
  }
  keys[i] = key;
  idx[i] = (int32_t)i;
}

__global__ void gather_rows_kernel(const double *__restrict__ X, const int32_t *__restrict__ perm, int64_t n, int64_t d,
                                   double *__restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t a = warp; a < n; a += nwarps) {
    const double *src = X + (int64_t)perm[a] * d;
    double *dst = out + a * d;
    for (int64_t k = lane; k < d; k += 32) dst[k] = src[k];
  }
}

// ---- 0c. k-means clusters of the cells (the top level of the internal order) ---------------------------
// The pruned candidate search (knn_tc.cu) bounds 256-cell tiles by balls; tiles are only compact when the
// cells of one tile come from one region of the data, which a Morton curve over four features cannot give
// in 100 dimensions.  So cells are first assigned to k-means clusters (a few Lloyd iterations over the
// highest-variance features, float32 -- a heuristic: ANY assignment gives correct graphs), and the Morton
// curve only orders the cells inside a cluster.  Everything is deterministic (no floating-point atomics):
// every rank of a sharded build derives the same order from the same data.
constexpr int kKmParts = 16;   // partial sums per cluster in the centroid update

// centroids are stored feature-major: cen[k * C + c]
template <int CPL>  // clusters per lane
__global__ void __launch_bounds__(256) kmeans_assign_kernel(const double *__restrict__ X, int64_t n, int64_t d,
                                                            const int32_t *__restrict__ sel, int nd,
                                                            const double *__restrict__ mu, const float *__restrict__ cen,
                                                            const float *__restrict__ cn, int C,
                                                            int32_t *__restrict__ cid) {
  extern __shared__ float km_sm[];
  float *scen = km_sm;                 // nd * C
  float *scn = scen + (size_t)nd * C;  // C
  float *smu = scn + C;                // kKmDims
  int *ssel = reinterpret_cast<int *>(smu + kKmDims);
  for (int t = threadIdx.x; t < nd * C; t += blockDim.x) scen[t] = cen[t];
  for (int t = threadIdx.x; t < C; t += blockDim.x) scn[t] = cn[t];
  for (int t = threadIdx.x; t < kKmDims; t += blockDim.x) {
    ssel[t] = t < nd ? sel[t] : 0;
    smu[t] = t < nd ? (float)mu[sel[t]] : 0.f;
  }
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t i = warp; i < n; i += nwarps) {
    float xr[kKmDims / 32];
#pragma unroll
    for (int j = 0; j < kKmDims / 32; ++j) {
      const int k = j * 32 + lane;
      xr[j] = k < nd ? (float)(X[i * d + ssel[k]] - (double)smu[k]) : 0.f;
    }
    float acc[CPL];
#pragma unroll
    for (int c = 0; c < CPL; ++c) acc[c] = 0.f;
#pragma unroll
    for (int j = 0; j < kKmDims / 32; ++j) {
      if (j * 32 < nd) {
        const int lim = min(32, nd - j * 32);
        for (int l = 0; l < lim; ++l) {
          const float xk = __shfl_sync(0xffffffffu, xr[j], l);
          const float *row = scen + (size_t)(j * 32 + l) * C + lane;
#pragma unroll
          for (int c = 0; c < CPL; ++c) acc[c] = fmaf(xk, row[32 * c], acc[c]);
        }
      }
    }
    float best = INFINITY;
    int bi = 0x7fffffff;
#pragma unroll
    for (int c = 0; c < CPL; ++c) {
      const int ci = lane + 32 * c;
      if (ci < C) {
        const float dist = scn[ci] - 2.f * acc[c];
        if (dist < best || (dist == best && ci < bi)) {
          best = dist;
          bi = ci;
        }
      }
    }
    for (int o = 16; o > 0; o >>= 1) {
      const float ob = __shfl_xor_sync(0xffffffffu, best, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (ob < best || (ob == best && oi < bi)) {
        best = ob;
        bi = oi;
      }
    }
    if (lane == 0) cid[i] = bi == 0x7fffffff ? 0 : bi;
  }
}

// seeds: cluster c starts at the cell in position (2c + 1) n / (2C) of the Morton order
__global__ void kmeans_seed_kernel(const double *__restrict__ X, int64_t n, int64_t d, const int32_t *__restrict__ sel,
                                   int nd, const double *__restrict__ mu, const int32_t *__restrict__ morton_perm,
                                   int C, float *__restrict__ cen) {
  const int c = blockIdx.x, k = threadIdx.x;
  if (k >= nd) return;
  const int64_t i = morton_perm[((2 * (int64_t)c + 1) * n) / (2 * (int64_t)C)];
  cen[(size_t)k * C + c] = (float)(X[i * d + sel[k]] - mu[sel[k]]);
}

// cn[c] = |centroid c|^2, summed in a fixed order
__global__ void kmeans_norms_kernel(const float *__restrict__ cen, int nd, int C, float *__restrict__ cn) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  float s = 0.f;
  for (int k = 0; k < nd; ++k) s = fmaf(cen[(size_t)k * C + c], cen[(size_t)k * C + c], s);
  cn[c] = s;
}

// members of cluster c are idx_s[seg[c] .. seg[c + 1]) (cells sorted by cluster id, stable)
__global__ void kmeans_segments_kernel(const int32_t *__restrict__ cid_sorted, int64_t n, int C, int64_t *__restrict__ seg) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c > C) return;
  int64_t lo = 0, hi = n;
  while (lo < hi) {
    const int64_t mid = (lo + hi) >> 1;
    if (cid_sorted[mid] < c)
      lo = mid + 1;
    else
      hi = mid;
  }
  seg[c] = lo;
}

__global__ void __launch_bounds__(kKmDims) kmeans_partial_kernel(const double *__restrict__ X, int64_t d,
                                                                 const int32_t *__restrict__ sel, int nd,
                                                                 const double *__restrict__ mu,
                                                                 const int32_t *__restrict__ idx_s,
                                                                 const int64_t *__restrict__ seg,
                                                                 float *__restrict__ partial) {
  const int part = blockIdx.x, c = blockIdx.y, k = threadIdx.x;
  const int64_t s0 = seg[c], cnt = seg[c + 1] - s0;
  const int64_t b = s0 + cnt * part / kKmParts, e = s0 + cnt * (part + 1) / kKmParts;
  float s = 0.f;
  if (k < nd) {
    const int64_t col = sel[k];
    const double m = mu[col];
    for (int64_t r = b; r < e; ++r) s += (float)(X[(int64_t)idx_s[r] * d + col] - m);
  }
  partial[((size_t)c * kKmParts + part) * kKmDims + k] = s;
}

__global__ void __launch_bounds__(kKmDims) kmeans_update_kernel(const float *__restrict__ partial,
                                                                const int64_t *__restrict__ seg, int nd, int C,
                                                                float *__restrict__ cen) {
  const int c = blockIdx.x, k = threadIdx.x;
  const int64_t cnt = seg[c + 1] - seg[c];
  if (k >= nd || cnt <= 0) return;  // an empty cluster keeps its centroid
  float s = 0.f;
  for (int part = 0; part < kKmParts; ++part) s += partial[((size_t)c * kKmParts + part) * kKmDims + k];
  cen[(size_t)k * C + c] = s / (float)cnt;
}

__global__ void iota_kernel(int32_t *p, int64_t n) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = (int32_t)i;
}

// final sort key: cluster id above the Morton key (whose 16 low bits are dropped)
__global__ void cluster_keys_kernel(unsigned long long *__restrict__ keys, const int32_t *__restrict__ cid, int64_t n) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) keys[i] = ((unsigned long long)(unsigned int)cid[i] << 48) | (keys[i] >> 16);
}

__global__ void cluster_of_key_kernel(const unsigned long long *__restrict__ keys_sorted, int64_t n,
                                      int32_t *__restrict__ cid_sorted) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) cid_sorted[i] = (int32_t)(keys_sorted[i] >> 48);
}

// ---- 1b. merge the per-(row, list) top-k1 lists of pass 1 into the emit threshold of pass 2 ----
// lists: [n][nlists][k1] floats sorted descending in s-space (s = x.y - n_j/2, larger = closer).
__global__ void merge_lists_kernel(const float *__restrict__ lists, int64_t row_begin, int64_t n, int nlists, int stride,
                                   int k1, const double *__restrict__ norm, const unsigned long long *ymax2_bits,
                                   double margin_c, double radius_factor, float *__restrict__ key2) {
  const int64_t i = row_begin + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;  // n = end of this call's row range
  const float *base = lists + (size_t)i * stride * k1;  // the first nlists of the row's `stride` lists are merged
  int idx[kMaxLists];
  for (int l = 0; l < nlists; ++l) idx[l] = 0;
  float sk = -INFINITY;
  for (int t = 0; t < k1; ++t) {  // t-th largest of the union
    float best = -INFINITY;
    int bl = 0;
    for (int l = 0; l < nlists; ++l) {
      const float v = idx[l] < k1 ? base[l * k1 + idx[l]] : -INFINITY;
      if (v > best) {
        best = v;
        bl = l;
      }
    }
    idx[bl]++;
    sk = best;
  }
  const double ni = norm[i];
  const double ymax2 = __longlong_as_double((long long)*ymax2_bits);
  const double m = margin_c * (ni + ymax2);
  double tau = ni - 2.0 * (double)sk;  // approximate k1-th smallest squared distance
  if (!(tau > 0.0)) tau = 0.0;
  if (!isfinite((double)sk)) tau = INFINITY;  // fewer than k1 columns seen: keep everything
  const double t2 = radius_factor * (tau + m) + m;  // emit radius^2, error margins on both sides
  // emit iff s >= (n_i - t2)/2 ; round the key down so float rounding cannot drop a candidate
  double key = 0.5 * (ni - t2);
  float kf = (float)key;
  if ((double)kf > key) kf = nextafterf(kf, -INFINITY);
  key2[i] = isfinite(tau) ? kf : -INFINITY;
}

// ---- 2. exact float64 distances of the candidates, eps_i ----------------------------------------
// One warp per row.  Candidates of row i are cand[cptr[i] .. cptr[i+1]) (columns ascending); d2buf (same indexing)
// receives the exact squared distances: sqrt(sum_k (x_ik - x_jk)^2) summed sequentially and unfused, the order of
// scikit-learn's ball tree, so d_ij equals the reference's bit for bit and d_ij = d_ji exactly.
// (Round 2 tried to compute a pair that is in both rows' lists only once -- pass 1 for j >= i, pass 2 looks the
// value up in row j's sorted list: 5.3 -> 5.8 ms at config 4.  The stage is bound by the latency chain of a lane's
// 25 dependent 256-bit gathers, not by bytes; halving the active lanes does not shorten that chain.  Reverted.)
__device__ __forceinline__ double exact_d2(const double *__restrict__ xi, const double *__restrict__ xj, int64_t d,
                                           bool vec4) {
  double acc = 0.0;
  if (vec4) {
    // rows are 32-byte aligned (d % 4 == 0): one 256-bit load per four features and lane instead of four
    // 8-byte ones; two feature quads are requested before either is consumed (more loads in flight per lane)
    int64_t k = 0;
    for (; k + 8 <= d; k += 8) {  // same sequential, unfused order as the scalar loop
      double a0, a1, a2, a3, b0, b1, b2, b3, c0, c1, c2, c3, e0, e1, e2, e3;
      asm volatile("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(a0), "=d"(a1), "=d"(a2), "=d"(a3) : "l"(xi + k));
      asm volatile("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(b0), "=d"(b1), "=d"(b2), "=d"(b3) : "l"(xj + k));
      asm volatile("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(c0), "=d"(c1), "=d"(c2), "=d"(c3) : "l"(xi + k + 4));
      asm volatile("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(e0), "=d"(e1), "=d"(e2), "=d"(e3) : "l"(xj + k + 4));
      double diff = __dsub_rn(a0, b0);
      acc = __dadd_rn(acc, __dmul_rn(diff, diff));
      diff = __dsub_rn(a1, b1);
      acc = __dadd_rn(acc, __dmul_rn(diff, diff));
      diff = __dsub_rn(a2, b2);
      acc = __dadd_rn(acc, __dmul_rn(diff, diff));
      diff = __dsub_rn(a3, b3);
      acc = __dadd_rn(acc, __dmul_rn(diff, diff));
      diff = __dsub_rn(c0, e0);
      acc = __dadd_rn(acc, __dmul_rn(diff, diff));
      diff = __dsub_rn(c1, e1);
      acc = __dadd_rn(acc, __dmul_rn(diff, diff));
      diff = __dsub_rn(c2, e2);
      acc = __dadd_rn(acc, __dmul_rn(diff, diff));
      diff = __dsub_rn(c3, e3);
      acc = __dadd_rn(acc, __dmul_rn(diff, diff));
    }
    for (; k < d; k += 4) {
      double a0, a1, a2, a3, b0, b1, b2, b3;
      asm volatile("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(a0), "=d"(a1), "=d"(a2), "=d"(a3) : "l"(xi + k));
      asm volatile("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(b0), "=d"(b1), "=d"(b2), "=d"(b3) : "l"(xj + k));
      double diff = __dsub_rn(a0, b0);
      acc = __dadd_rn(acc, __dmul_rn(diff, diff));
      diff = __dsub_rn(a1, b1);
      acc = __dadd_rn(acc, __dmul_rn(diff, diff));
      diff = __dsub_rn(a2, b2);
      acc = __dadd_rn(acc, __dmul_rn(diff, diff));
      diff = __dsub_rn(a3, b3);
      acc = __dadd_rn(acc, __dmul_rn(diff, diff));
    }
  } else {
    for (int64_t k = 0; k < d; ++k) {  // sequential, unfused: the ball tree's rdist order
      const double diff = __dsub_rn(xi[k], xj[k]);
      acc = __dadd_rn(acc, __dmul_rn(diff, diff));
    }
  }
  return acc;
}

__global__ void refine_dist_kernel(const double *__restrict__ X, int64_t row_begin, int64_t n, int64_t d,
                                   const int32_t *__restrict__ cand, const int64_t *__restrict__ cptr, int k1,
                                   double bandwidth_scale,
                                   double *__restrict__ d2buf, double *__restrict__ eps, int *__restrict__ err_flag) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const bool vec4 = (d % 4 == 0) && ((reinterpret_cast<uintptr_t>(X) & 31) == 0);
  for (int64_t i = warp; i < n; i += nwarps) {
    // i is a LOCAL row (n = number of local rows); the query point is global row row_begin + i
    const int c = (int)(cptr[i + 1] - cptr[i]);
    const int32_t *ci = cand + cptr[i];
    double *di = d2buf + cptr[i];
    const double *xi = X + (row_begin + i) * d;
    for (int t = lane; t < c; t += 32) di[t] = exact_d2(xi, X + (int64_t)ci[t] * d, d, vec4);
    __syncwarp();
    if (c < k1) {
      if (lane == 0) atomicExch(err_flag, 1);
      continue;
    }
    // k1-th smallest (value, slot) by repeated extraction of the next minimum
    double last_v = -1.0;
    int last_t = -1;
    for (int s = 0; s < k1; ++s) {
      double best_v = INFINITY;
      int best_t = 0x7fffffff;
      for (int t = lane; t < c; t += 32) {
        const double v = di[t];
        const bool after = (v > last_v) || (v == last_v && t > last_t);
        if (after && (v < best_v || (v == best_v && t < best_t))) {
          best_v = v;
          best_t = t;
        }
      }
      for (int o = 16; o > 0; o >>= 1) {
        const double ov = __shfl_xor_sync(0xffffffffu, best_v, o);
        const int ot = __shfl_xor_sync(0xffffffffu, best_t, o);
        if (ov < best_v || (ov == best_v && ot < best_t)) {
          best_v = ov;
          best_t = ot;
        }
      }
      last_v = best_v;
      last_t = best_t;
    }
    if (lane == 0) {
      double e = sqrt(last_v) * bandwidth_scale;
      eps[i] = e > DBL_EPSILON ? e : DBL_EPSILON;
    }
  }
}

// launch helper
static int refine_distances(const double *X, int64_t row_begin, int64_t nloc, int64_t d, const int32_t *cand,
                            const int64_t *cptr, int k1, double bandwidth_scale, double *d2buf, double *eps, int *err,
                            cudaStream_t stream);

__device__ __forceinline__ double alpha_decay(double dist, double eps, double decay) {
  // decay == 0 stands for the reference's decay=None: the unweighted kNN kernel (graphtools kNNGraph with
  // decay=None -> kneighbors_graph(mode="connectivity")): 1 for the knn nearest (self included), else 0
  if (decay == 0.0) return dist <= eps ? 1.0 : 0.0;
  double v = exp(-pow(dist / eps, decay));
  if (isnan(v)) v = 1.0;
  return v;
}

// ---- 3. kernel values, symmetrisation bookkeeping --------------------------------------------------
// Per candidate slot: if K_ij >= thresh the slot is kept with value (K_ij + K_ji)/2 (K_ji zeroed
// below thresh).  When K_ji is zero the mirrored entry (j, i) does not exist in row j's own list and
// must be appended there: the slot is flagged (column stored as ~j) and extra[j] is bumped.
// kraw (optional) keeps the un-symmetrised K_ij for export.
__global__ void kernel_values_kernel(int64_t n, int32_t *__restrict__ cand, const int64_t *__restrict__ cptr,
                                     double *__restrict__ d2buf, const double *__restrict__ eps, double decay,
                                     double thresh, int32_t *__restrict__ kept, int32_t *__restrict__ extra,
                                     double *__restrict__ kraw) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t i = warp; i < n; i += nwarps) {
    const int c = (int)(cptr[i + 1] - cptr[i]);
    int32_t *ci = cand + cptr[i];
    double *di = d2buf + cptr[i];
    const double ei = eps[i];
    int nk = 0;
    for (int t = lane; t < c; t += 32) {
      const int32_t j = ci[t];
      const double dist = sqrt(di[t]);
      const double kij = alpha_decay(dist, ei, decay);
      if (kij >= thresh) {
        double kji = (j == (int32_t)i) ? kij : alpha_decay(dist, eps[j], decay);
        if (kji < thresh) kji = 0.0;
        di[t] = (kij + kji) / 2;
        if (kraw) kraw[cptr[i] + t] = kij;
        if (kji == 0.0) {
          ci[t] = ~j;
          atomicAdd(extra + j, 1);
        }
        ++nk;
      } else {
        ci[t] = INT32_MIN;  // dead slot
      }
    }
    for (int o = 16; o > 0; o >>= 1) nk += __shfl_xor_sync(0xffffffffu, nk, o);
    if (lane == 0) kept[i] = nk;
  }
}

__global__ void add_counts_kernel(const int32_t *__restrict__ a, const int32_t *__restrict__ b, int64_t n,
                                  int32_t *__restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = a[i] + (b ? b[i] : 0);
  if (i == n) out[i] = 0;
}

// Scatter the kept slots into the CSR rows: own entries first (compacted in slot order), mirrored
// entries of flagged slots appended behind row j's own entries through an atomic cursor.
__global__ void fill_sym_kernel(int64_t n, const int32_t *__restrict__ cand, const int64_t *__restrict__ cptr,
                                const double *__restrict__ vbuf, const int32_t *__restrict__ row_ptr,
                                const int32_t *__restrict__ kept, int32_t *__restrict__ cursor,
                                int32_t *__restrict__ col, double *__restrict__ val) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t i = warp; i < n; i += nwarps) {
    const int c = (int)(cptr[i + 1] - cptr[i]);
    const int32_t *ci = cand + cptr[i];
    const double *vi = vbuf + cptr[i];
    int base = row_ptr[i];
    for (int t0 = 0; t0 < c; t0 += 32) {
      const int t = t0 + lane;
      int32_t j = INT32_MIN;
      double v = 0.0;
      if (t < c) {
        j = ci[t];
        v = vi[t];
      }
      const bool live = j != INT32_MIN;
      const unsigned m = __ballot_sync(0xffffffffu, live);
      if (live) {
        const bool flagged = j < 0;
        const int32_t jj = flagged ? ~j : j;
        const int pos = base + __popc(m & ((1u << lane) - 1u));
        col[pos] = jj;
        val[pos] = v;
        if (flagged) {
          const int p2 = row_ptr[jj] + kept[jj] + atomicAdd(cursor + jj, 1);
          col[p2] = (int32_t)i;
          val[p2] = v;
        }
      }
      base += __popc(m);
    }
  }
}

// Rows come out of fill_sym_kernel as [own entries, ascending by column | mirrored entries, any order], all
// columns of a row distinct.  The final position of an entry is the number of entries of its row with a smaller
// column: its index among the own entries (a binary search for a mirrored one) plus a count over the few
// mirrored ones -- one pass over the matrix instead of a segmented sort of all of it.
__global__ void merge_rows_kernel(int64_t n, const int32_t *__restrict__ row_ptr, const int32_t *__restrict__ kept,
                                  const int32_t *__restrict__ ucol, const double *__restrict__ uval,
                                  int32_t *__restrict__ col, double *__restrict__ val) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t i = warp; i < n; i += nwarps) {
    const int base = row_ptr[i], k = kept[i], x = row_ptr[i + 1] - base - k;
    const int32_t *own = ucol + base, *ext = ucol + base + k;
    for (int t = lane; t < k; t += 32) {
      const int32_t c = own[t];
      int cnt = 0;
      for (int e = 0; e < x; ++e) cnt += ext[e] < c ? 1 : 0;
      col[base + t + cnt] = c;
      val[base + t + cnt] = uval[base + t];
    }
    for (int e = lane; e < x; e += 32) {
      const int32_t c = ext[e];
      int cnt = 0;
      for (int f = 0; f < x; ++f) cnt += ext[f] < c ? 1 : 0;
      int lo = 0, hi = k;  // own entries below c
      while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (own[mid] < c)
          lo = mid + 1;
        else
          hi = mid;
      }
      col[base + lo + cnt] = c;
      val[base + lo + cnt] = uval[base + k + e];
    }
  }
}

__global__ void max_int_kernel(const int32_t *__restrict__ a, int64_t n, int32_t *out) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int32_t m = 0;
  for (; i < n; i += (int64_t)gridDim.x * blockDim.x) m = max(m, a[i]);
  for (int o = 16; o > 0; o >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0 && m > 0) atomicMax(out, m);
}

// Compact the un-symmetrised kernel (slot order -> sorted later on the host side of the test).
__global__ void fill_raw_kernel(int64_t n, const int32_t *__restrict__ cand, const int64_t *__restrict__ cptr,
                                const double *__restrict__ kraw, const int64_t *__restrict__ out_ptr,
                                int32_t *__restrict__ col, double *__restrict__ val) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t i = warp; i < n; i += nwarps) {
    const int c = (int)(cptr[i + 1] - cptr[i]);
    int64_t base = out_ptr[i];
    for (int t0 = 0; t0 < c; t0 += 32) {
      const int t = t0 + lane;
      int32_t j = INT32_MIN;
      if (t < c) j = cand[cptr[i] + t];
      const bool live = j != INT32_MIN;
      const unsigned m = __ballot_sync(0xffffffffu, live);
      if (live) {
        const int64_t pos = base + __popc(m & ((1u << lane) - 1u));
        col[pos] = j < 0 ? ~j : j;
        val[pos] = kraw[cptr[i] + t];
      }
      base += __popc(m);
    }
  }
}

// ---- 4. anisotropy and Laplacian (rows sorted by column) ------------------------------------------
__global__ void row_sum_kernel(int64_t n, const int32_t *__restrict__ row_ptr, const double *__restrict__ val,
                               double *__restrict__ q) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t i = warp; i < n; i += nwarps) {
    double s = 0.0;
    for (int e = row_ptr[i] + lane; e < row_ptr[i + 1]; e += 32) s += val[e];
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) q[i] = s;
  }
}

// K_ij <- K_ij (q_i q_j)^-a off the diagonal; L_ij = -K_ij; L_ii = sum_j K_ij (j != i)
__global__ void laplacian_kernel(int64_t n, const int32_t *__restrict__ row_ptr, const int32_t *__restrict__ col,
                                 double *__restrict__ val, const double *__restrict__ q, double anisotropy) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t i = warp; i < n; i += nwarps) {
    const double qi = q[i];
    double dw = 0.0;
    int diag = -1;
    for (int e = row_ptr[i] + lane; e < row_ptr[i + 1]; e += 32) {
      const int32_t j = col[e];
      if (j == (int32_t)i) {
        diag = e;
        continue;
      }
      double w = val[e];
      if (anisotropy != 0.0) {
        const double qq = qi * q[j];
        w = (anisotropy == 1.0 ? 1.0 / qq : pow(qq, -anisotropy)) * w;
      }
      val[e] = -w;
      dw += w;
    }
    for (int o = 16; o > 0; o >>= 1) {
      dw += __shfl_xor_sync(0xffffffffu, dw, o);
      diag = max(diag, __shfl_xor_sync(0xffffffffu, diag, o));
    }
    if (lane == 0 && diag >= 0) val[diag] = dw;
  }
}

// candidate pairs (row << 32 | col), sorted: cptr[r] = first pair of row r (lower bound), cptr[n] = total
__global__ void pair_row_ptr_kernel(const unsigned long long *__restrict__ keys, int64_t total, int64_t row_begin,
                                    int64_t n, int64_t *__restrict__ cptr) {
  const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;  // local row; n = number of local rows
  if (r > n) return;
  const unsigned long long target = (unsigned long long)(row_begin + r) << 32;
  int64_t lo = 0, hi = total;
  while (lo < hi) {
    const int64_t mid = (lo + hi) >> 1;
    if (keys[mid] < target)
      lo = mid + 1;
    else
      hi = mid;
  }
  cptr[r] = lo;
}

__global__ void pair_cols_kernel(const unsigned long long *__restrict__ keys, int64_t total, int32_t *__restrict__ col) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x)
    col[i] = (int32_t)(keys[i] & 0xffffffffull);
}

__global__ void row_counts_kernel(const int64_t *__restrict__ cptr, int64_t n, int32_t *__restrict__ cnt) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) cnt[i] = (int32_t)(cptr[i + 1] - cptr[i]);
}

__global__ void row_counts64_kernel(const int64_t *__restrict__ cptr, int64_t n, int64_t *__restrict__ cnt) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) cnt[i] = cptr[i + 1] - cptr[i];
}

__global__ void max_count_kernel(const int64_t *__restrict__ cptr, int64_t n, int32_t *out) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int32_t m = 0;
  for (; i < n; i += (int64_t)gridDim.x * blockDim.x) m = max(m, (int32_t)(cptr[i + 1] - cptr[i]));
  for (int o = 16; o > 0; o >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0) atomicMax(out, m);
}

// MELD_B200_TIMING=1: print per-stage wall times of a build to stderr (synchronises after each stage).
struct StageTimer {
  bool on;
  cudaStream_t stream;
  cudaEvent_t e0, e1;
  explicit StageTimer(cudaStream_t s) : on(getenv("MELD_B200_TIMING") != nullptr), stream(s) {
    if (on) {
      cudaEventCreate(&e0);
      cudaEventCreate(&e1);
      cudaEventRecord(e0, stream);
    }
  }
  void lap(const char *what) {
    if (!on) return;
    cudaEventRecord(e1, stream);
    cudaEventSynchronize(e1);
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    fprintf(stderr, "[meld_b200 timing] %-28s %9.3f ms\n", what, ms);
    cudaEventRecord(e0, stream);
  }
  ~StageTimer() {
    if (on) {
      cudaEventDestroy(e0);
      cudaEventDestroy(e1);
    }
  }
};

static int grid_for_rows(int64_t n) {
  int64_t b = ceil_div(n > 0 ? n : 1, 256);
  const int64_t cap = (int64_t)sm_count() * 8;
  return (int)(b < cap ? b : (cap > 0 ? cap : 1184));
}

static int warp_grid(int64_t n_rows, int threads) {
  const int64_t warps_per_block = threads / 32;
  int64_t b = ceil_div(n_rows, warps_per_block);
  const int64_t cap = (int64_t)sm_count() * 16;
  return (int)(b < cap ? (b > 0 ? b : 1) : cap);
}

static int refine_distances(const double *X, int64_t row_begin, int64_t nloc, int64_t d, const int32_t *cand,
                            const int64_t *cptr, int k1, double bandwidth_scale, double *d2buf, double *eps, int *err,
                            cudaStream_t stream) {
  refine_dist_kernel<<<warp_grid(nloc, 128), 128, 0, stream>>>(X, row_begin, nloc, d, cand, cptr, k1, bandwidth_scale,
                                                              d2buf, eps, err);
  MELD_LAUNCH_CHECK();
  return 0;
}

}  // namespace meld

using namespace meld;

// Internal cell order: k-means cluster first, Morton curve of the four highest-variance features inside a
// cluster.  perm[a] = original index of the cell placed at position a; Xp = X rows in that order;
// cl = the clusters (cluster id of every position, centroids), left empty when clustering is off.
static int cell_order(const double *X, int64_t n, int64_t d, cudaStream_t stream, DevBuf<int32_t> &perm,
                      DevBuf<double> &Xp, CellClusters &cl) {
  StageTimer tm(stream);
  DevBuf<double> partial, var;
  DevBuf<double> &mu = cl.mu;
  DevBuf<int32_t> &cid_sorted = cl.cid;
  MELD_CHECK(partial.alloc((size_t)kMeanBlocks * d));
  MELD_CHECK(mu.alloc((size_t)d));
  MELD_CHECK(var.alloc((size_t)d));
  col_partial_sums_kernel<<<kMeanBlocks, 256, 0, stream>>>(X, n, d, partial.p);
  MELD_LAUNCH_CHECK();
  col_mean_kernel<<<(unsigned)ceil_div(d, 256), 256, 0, stream>>>(partial.p, n, d, kMeanBlocks, mu.p);
  MELD_LAUNCH_CHECK();
  col_partial_sq_kernel<<<kMeanBlocks, 256, 0, stream>>>(X, mu.p, n, d, partial.p);
  MELD_LAUNCH_CHECK();
  col_mean_kernel<<<(unsigned)ceil_div(d, 256), 256, 0, stream>>>(partial.p, n, d, kMeanBlocks, var.p);
  MELD_LAUNCH_CHECK();
  std::vector<double> h_mu((size_t)d), h_var((size_t)d);
  MELD_CUDA(cudaMemcpyAsync(h_mu.data(), mu.p, (size_t)d * sizeof(double), cudaMemcpyDeviceToHost, stream));
  MELD_CUDA(cudaMemcpyAsync(h_var.data(), var.p, (size_t)d * sizeof(double), cudaMemcpyDeviceToHost, stream));
  MELD_SYNC(stream);
  tm.lap("  order: means + variances");
  std::vector<int> order((size_t)d);
  for (int64_t k = 0; k < d; ++k) order[(size_t)k] = (int)k;
  const int ndim = d < 4 ? (int)d : 4;
  int nd = d < kKmDims ? (int)d : kKmDims;  // features of the k-means step
  std::partial_sort(order.begin(), order.begin() + nd, order.end(),
                    [&](int a, int b) { return h_var[a] > h_var[b] || (h_var[a] == h_var[b] && a < b); });
  // Clustering (and the projection bound that rides on it) only needs the features that carry the variance: keep the
  // leading ones up to km_var_pct % of the total, in steps of 32.  PCA-like data (config 4: 99.8 % in 32 of 100
  // features) runs the assignment and the tile projections over a quarter of the columns; a flat spectrum keeps all
  // 128.  Any subset is valid: a bound on the distance within a feature subset bounds the full distance.
  if (tuning().km_var_pct > 0 && tuning().km_var_pct < 100 && nd > 32) {
    double tot = 0.0, cum = 0.0;
    for (int64_t k = 0; k < d; ++k) tot += h_var[(size_t)k];
    int keep = nd;
    for (int k = 0; k < nd; ++k) {
      cum += h_var[(size_t)order[(size_t)k]];
      if (cum >= 0.01 * tuning().km_var_pct * tot) {
        keep = k + 1;
        break;
      }
    }
    keep = (keep + 31) / 32 * 32;
    if (keep < nd) nd = keep;
  }
  MortonDims md;
  md.ndim = ndim;
  for (int k = 0; k < 4; ++k) {
    const int dim = k < ndim ? order[(size_t)k] : 0;
    md.dim[k] = dim;
    md.mu[k] = h_mu[(size_t)dim];
    const double sd = sqrt(h_var[(size_t)dim]);
    md.inv_range[k] = sd > 0 ? 1.0 / (12.0 * sd) : 0.0;
  }
  DevBuf<unsigned long long> keys, keys_sorted;
  DevBuf<int32_t> idx;
  MELD_CHECK(keys.alloc((size_t)n));
  MELD_CHECK(keys_sorted.alloc((size_t)n));
  MELD_CHECK(idx.alloc((size_t)n));
  MELD_CHECK(perm.alloc((size_t)n));
  morton_keys_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, stream>>>(X, n, d, md, keys.p, idx.p);
  MELD_LAUNCH_CHECK();
  size_t tmp_bytes = 0, tmp_bytes32 = 0;
  MELD_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, keys.p, keys_sorted.p, idx.p, perm.p, (int)n, 0, 64,
                                            stream));
  int C = tuning().clusters;
  if (C > kKmMaxC) C = kKmMaxC;
  const int64_t min_cells = tuning().cluster_cells > 0 ? tuning().cluster_cells : 1024;
  while (C >= 32 && (int64_t)C * min_cells > n) C -= 32;  // at least ~4 tiles per cluster
  C = C / 32 * 32;
  DevBuf<int32_t> cid, cid_s, idx_s;
  if (C >= 32) {
    MELD_CHECK(cid.alloc((size_t)n));
    MELD_CHECK(cid_s.alloc((size_t)n));
    MELD_CHECK(idx_s.alloc((size_t)n));
    MELD_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes32, cid.p, cid_s.p, idx.p, idx_s.p, (int)n, 0, 8, stream));
    if (tmp_bytes32 > tmp_bytes) tmp_bytes = tmp_bytes32;
  }
  DevBuf<unsigned char> tmp;
  MELD_CHECK(tmp.alloc(tmp_bytes));
  if (C >= 32) {
    // Morton order first: the seeds are evenly spaced along the curve
    MELD_CUDA(cub::DeviceRadixSort::SortPairs(tmp.p, tmp_bytes, keys.p, keys_sorted.p, idx.p, perm.p, (int)n, 0, 64,
                                              stream));
    DevBuf<int32_t> &sel = cl.sel;
    DevBuf<float> &cen = cl.cen;
    DevBuf<float> cn, kpart;
    DevBuf<int64_t> seg;
    cl.nd = nd;
    cl.C = C;
    MELD_CHECK(sel.alloc((size_t)kKmDims));
    MELD_CHECK(cen.alloc((size_t)nd * C));
    MELD_CHECK(cn.alloc((size_t)C));
    MELD_CHECK(kpart.alloc((size_t)C * kKmParts * kKmDims));
    MELD_CHECK(seg.alloc((size_t)C + 1));
    int32_t h_sel[kKmDims];
    for (int k = 0; k < kKmDims; ++k) h_sel[k] = k < nd ? order[(size_t)k] : 0;
    MELD_CUDA(cudaMemcpyAsync(sel.p, h_sel, sizeof(h_sel), cudaMemcpyHostToDevice, stream));
    MELD_SYNC(stream);  // h_sel is a stack array
    tm.lap("  order: morton sort");
    kmeans_seed_kernel<<<C, kKmDims, 0, stream>>>(X, n, d, sel.p, nd, mu.p, perm.p, C, cen.p);
    MELD_LAUNCH_CHECK();
    const size_t smem = ((size_t)nd * C + C + kKmDims) * sizeof(float) + kKmDims * sizeof(int);
    auto assign = C == 32 ? kmeans_assign_kernel<1> : C == 64 ? kmeans_assign_kernel<2>
                : C == 96 ? kmeans_assign_kernel<3> : kmeans_assign_kernel<4>;
    MELD_CUDA(cudaFuncSetAttribute(assign, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int iters = tuning().kmeans_iters;
    if (iters < 0) iters = 0;
    if (iters > 16) iters = 16;
    for (int it = 0; it <= iters; ++it) {
      kmeans_norms_kernel<<<(unsigned)ceil_div(C, 128), 128, 0, stream>>>(cen.p, nd, C, cn.p);
      MELD_LAUNCH_CHECK();
      assign<<<sm_count() * 8, 256, smem, stream>>>(X, n, d, sel.p, nd, mu.p, cen.p, cn.p, C, cid.p);
      MELD_LAUNCH_CHECK();
      if (it == iters) break;
      // centroid update in a fixed summation order: members sorted by cluster (stable), kKmParts partial sums each
      MELD_CUDA(cub::DeviceRadixSort::SortPairs(tmp.p, tmp_bytes, cid.p, cid_s.p, idx.p, idx_s.p, (int)n, 0, 8, stream));
      kmeans_segments_kernel<<<(unsigned)ceil_div(C + 1, 128), 128, 0, stream>>>(cid_s.p, n, C, seg.p);
      MELD_LAUNCH_CHECK();
      kmeans_partial_kernel<<<dim3(kKmParts, (unsigned)C), kKmDims, 0, stream>>>(X, d, sel.p, nd, mu.p, idx_s.p, seg.p,
                                                                                kpart.p);
      MELD_LAUNCH_CHECK();
      kmeans_update_kernel<<<C, kKmDims, 0, stream>>>(kpart.p, seg.p, nd, C, cen.p);
      MELD_LAUNCH_CHECK();
    }
    tm.lap("  order: k-means");
    cluster_keys_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, stream>>>(keys.p, cid.p, n);
    MELD_LAUNCH_CHECK();
    MELD_CUDA(cub::DeviceRadixSort::SortPairs(tmp.p, tmp_bytes, keys.p, keys_sorted.p, idx.p, perm.p, (int)n, 0, 64,
                                              stream));
    MELD_CHECK(cid_sorted.alloc((size_t)n));
    cluster_of_key_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, stream>>>(keys_sorted.p, n, cid_sorted.p);
    MELD_LAUNCH_CHECK();
  } else {
    MELD_CUDA(cub::DeviceRadixSort::SortPairs(tmp.p, tmp_bytes, keys.p, keys_sorted.p, idx.p, perm.p, (int)n, 0, 64,
                                              stream));
  }
  tm.lap("  order: final sort");
  MELD_CHECK(Xp.alloc((size_t)n * d));
  gather_rows_kernel<<<warp_grid(n, 256), 256, 0, stream>>>(X, perm.p, n, d, Xp.p);
  MELD_LAUNCH_CHECK();
  // no host wait: the temporaries are arena bumps or stream-ordered pool frees
  tm.lap("  order: gather rows");
  return 0;
}

// Steps 0 and 1: centred norms, pass 1 (eps upper bounds), pass 2 (candidate pairs), pairs sorted into a
// CSR of candidates: cand[cptr[i] .. cptr[i+1]) = candidate columns of row i, ascending.
struct Candidates {
  int64_t row_begin = 0, row_end = 0;  // query rows of this call; cptr is local (row_end - row_begin + 1 entries)
  DevBuf<float> key2;
  DevBuf<int64_t> cptr;   // n + 1
  DevBuf<int32_t> cand;   // total
  int64_t total = 0;
  int64_t pair_cap = 0;
  int passes = 0;
  int32_t max_per_row = 0;
  int64_t retries = 0;
  double pass1_ms = 0, pass2_ms = 0, gemm_flops_per_pass = 0;  // CUDA-event times of the two GEMM passes
  double gemm_flops_pass1 = 0;  // pruned search: pass 1 (window + list pass) and pass 2 multiply different tile sets
  double tile_frac = 1.0;       // (row tile, column tile) products of pass 2 / all of them
};

static int candidate_search(const double *X, int64_t n, int64_t d, int k1, double radius_factor, bool simt,
                            int64_t row_begin, int64_t row_end, const CellClusters *cl, cudaStream_t stream,
                            Candidates &out) {
  const int64_t nloc = row_end - row_begin;
  out.row_begin = row_begin;
  out.row_end = row_end;
  // -- 0. means / norms
  DevBuf<double> partial, mu, norm;
  DevBuf<unsigned long long> ymax2;
  MELD_CHECK(partial.alloc((size_t)kMeanBlocks * d));
  MELD_CHECK(mu.alloc((size_t)d));
  MELD_CHECK(norm.alloc((size_t)n));
  MELD_CHECK(ymax2.alloc(1));
  MELD_CUDA(cudaMemsetAsync(ymax2.p, 0, sizeof(unsigned long long), stream));
  col_partial_sums_kernel<<<kMeanBlocks, 256, 0, stream>>>(X, n, d, partial.p);
  MELD_LAUNCH_CHECK();
  col_mean_kernel<<<(unsigned)ceil_div(d, 256), 256, 0, stream>>>(partial.p, n, d, kMeanBlocks, mu.p);
  MELD_LAUNCH_CHECK();
  row_norms_kernel<<<warp_grid(n, 256), 256, 0, stream>>>(X, mu.p, n, d, norm.p, ymax2.p);
  MELD_LAUNCH_CHECK();

  // -- 1. candidate search
  SearchPlan plan;
  MELD_CHECK(search_plan(simt, n, d, k1, row_begin, row_end, &plan));
  SearchState st;
  MELD_CHECK(search_prepare(plan, X, mu.p, norm.p, stream, &st));
  DevBuf<float> lists;
  MELD_CHECK(lists.alloc((size_t)n * plan.nlists * k1));
  MELD_CHECK(out.key2.alloc((size_t)n));
  StageTimer tm(stream);
  // the two GEMM passes are always timed with CUDA events (bench.py reports their tensor throughput);
  // reading the timers costs nothing extra: the stream is synchronised after pass 2 anyway
  cudaEvent_t ev[4];
  for (int i = 0; i < 4; ++i) MELD_CUDA(cudaEventCreate(&ev[i]));
  struct EvGuard {
    cudaEvent_t *e;
    ~EvGuard() {
      for (int i = 0; i < 4; ++i) cudaEventDestroy(e[i]);
    }
  } ev_guard{ev};
  const bool prune = plan.prune && !plan.simt;
  TileLists tl;
  MELD_CUDA(cudaEventRecord(ev[0], stream));
  if (prune) {
    // window pass (own tile +- window -> a first bound of eps_i), tile lists from that bound, list pass over
    // the surviving tiles minus the window: the union of both passes' lists holds the exact k1 nearest
    MELD_CHECK(tc_tile_balls(plan, X, cl, stream, &st));
    tm.lap("  tile balls");
    MELD_CHECK(tc_tile_lists(plan, st, 0, nullptr, nullptr, 0, stream, &tl));
    tl.chunks = 1;
    tl.slot0 = 0;
    MELD_CHECK(tc_pass(plan, st, 1, lists.p, nullptr, nullptr, nullptr, 0, stream, &tl));
    tm.lap("  window pass");
    merge_lists_kernel<<<(unsigned)ceil_div(nloc, 128), 128, 0, stream>>>(lists.p, row_begin, row_end, 2, plan.nlists, k1,
                                                                           norm.p, ymax2.p, plan.margin_c, radius_factor,
                                                                           out.key2.p);
    MELD_LAUNCH_CHECK();
    tm.lap("  window merge");
    MELD_CHECK(tc_tile_lists(plan, st, 1, out.key2.p, norm.p, 1, stream, &tl));
    tm.lap("  tile lists");
    tl.chunks = plan.nchunk;
    tl.slot0 = 1;
    MELD_CHECK(tc_pass(plan, st, 1, lists.p, nullptr, nullptr, nullptr, 0, stream, &tl));
    tm.lap("  list pass");
  } else {
    MELD_CHECK(search_pass1(plan, st, lists.p, stream));
  }
  MELD_CUDA(cudaEventRecord(ev[1], stream));
  tm.lap("search pass 1 (top-k)");
  merge_lists_kernel<<<(unsigned)ceil_div(nloc, 128), 128, 0, stream>>>(lists.p, row_begin, row_end, plan.nlists,
                                                                         plan.nlists, k1, norm.p, ymax2.p, plan.margin_c,
                                                                         radius_factor, out.key2.p);
  MELD_LAUNCH_CHECK();
  lists.release();
  out.passes = 1;
  if (prune) {
    MELD_CHECK(tc_tile_lists(plan, st, 2, out.key2.p, norm.p, 2, stream, &tl));
    tl.chunks = plan.nchunk;
    tl.slot0 = 0;
  }

  // pass 2 appends (row, col) pairs to one global buffer; sized generously, retried with the exact
  // total if it overflows
  int64_t pair_cap = nloc * (k1 <= 8 ? 96 : 192);
  if (pair_cap > nloc * n) pair_cap = nloc * n;
  DevBuf<unsigned long long> pairs, pairs_sorted, gcount;
  MELD_CHECK(gcount.alloc(1));
  unsigned long long h_total = 0;
  for (int attempt = 0; attempt < 2; ++attempt) {
    MELD_CHECK(pairs.alloc((size_t)pair_cap));
    MELD_CUDA(cudaMemsetAsync(gcount.p, 0, sizeof(unsigned long long), stream));
    tm.lap("merge + alloc");
    MELD_CUDA(cudaEventRecord(ev[2], stream));
    if (prune)
      MELD_CHECK(tc_pass(plan, st, 2, nullptr, out.key2.p, pairs.p, gcount.p, pair_cap, stream, &tl));
    else
      MELD_CHECK(search_pass2(plan, st, out.key2.p, pairs.p, gcount.p, pair_cap, stream));
    MELD_CUDA(cudaEventRecord(ev[3], stream));
    tm.lap("search pass 2 (emit)");
    ++out.passes;
    unsigned long long h_steps[4] = {0, 0, 0, 0};
    MELD_CUDA(cudaMemcpyAsync(&h_total, gcount.p, sizeof(h_total), cudaMemcpyDeviceToHost, stream));
    if (prune) MELD_CUDA(cudaMemcpyAsync(h_steps, st.tl_steps.p, sizeof(h_steps), cudaMemcpyDeviceToHost, stream));
    MELD_SYNC(stream);
    {
      float ms1 = 0.f, ms2 = 0.f;
      cudaEventElapsedTime(&ms1, ev[0], ev[1]);
      cudaEventElapsedTime(&ms2, ev[2], ev[3]);
      out.pass1_ms = ms1;
      out.pass2_ms = ms2;
      out.gemm_flops_per_pass = 2.0 * (double)nloc * (double)plan.n_pad_cols * (double)plan.kp_used;
      out.gemm_flops_pass1 = out.gemm_flops_per_pass;
      if (prune) {  // count the (128-row tile x 256-column tile) products actually issued
        const double per_step = 2.0 * 128.0 * 256.0 * (double)plan.kp_used;
        const double full = out.gemm_flops_per_pass;
        out.gemm_flops_pass1 = per_step * (double)(h_steps[0] + h_steps[1]);
        out.gemm_flops_per_pass = per_step * (double)h_steps[2];
        out.tile_frac = full > 0 ? out.gemm_flops_per_pass / full : 1.0;
      }
    }
    if ((int64_t)h_total <= pair_cap) break;
    if (attempt == 1) {
      set_error("knn_graph_build: candidate overflow persisted (%llu pairs > capacity %lld)", h_total,
                (long long)pair_cap);
      return MELD_B200_ERR_INTERNAL;
    }
    out.retries = 1;
    pair_cap = (int64_t)h_total + 1024;
  }
  search_release(&st);
  out.total = (int64_t)h_total;
  out.pair_cap = pair_cap;
  MELD_REQUIRE(out.total < (int64_t)2147483647, "knn_graph_build: %lld candidate pairs overflow int32 sort size",
               (long long)out.total);

  // sort the pairs by (row, col) and cut them into rows
  MELD_CHECK(pairs_sorted.alloc((size_t)out.total));
  {
    int end_bit = 33;
    while (end_bit < 64 && ((unsigned long long)n >> (end_bit - 32)) != 0) ++end_bit;
    size_t tmp_bytes = 0;
    MELD_CUDA(cub::DeviceRadixSort::SortKeys(nullptr, tmp_bytes, pairs.p, pairs_sorted.p, (int)out.total, 0, end_bit,
                                             stream));
    DevBuf<unsigned char> tmp;
    MELD_CHECK(tmp.alloc(tmp_bytes));
    MELD_CUDA(cub::DeviceRadixSort::SortKeys(tmp.p, tmp_bytes, pairs.p, pairs_sorted.p, (int)out.total, 0, end_bit,
                                             stream));
    MELD_CHECK(out.cptr.alloc((size_t)nloc + 1));
    MELD_CHECK(out.cand.alloc((size_t)out.total));
    pair_row_ptr_kernel<<<(unsigned)ceil_div(nloc + 1, 256), 256, 0, stream>>>(pairs_sorted.p, out.total, row_begin, nloc,
                                                                              out.cptr.p);
    MELD_LAUNCH_CHECK();
    pair_cols_kernel<<<sm_count() * 8, 256, 0, stream>>>(pairs_sorted.p, out.total, out.cand.p);
    MELD_LAUNCH_CHECK();
    DevBuf<int32_t> mx;
    MELD_CHECK(mx.alloc(1));
    MELD_CUDA(cudaMemsetAsync(mx.p, 0, sizeof(int32_t), stream));
    max_count_kernel<<<296, 256, 0, stream>>>(out.cptr.p, nloc, mx.p);
    MELD_LAUNCH_CHECK();
    MELD_CUDA(cudaMemcpyAsync(&out.max_per_row, mx.p, sizeof(int32_t), cudaMemcpyDeviceToHost, stream));
    MELD_SYNC(stream);  // temporaries die here
  }
  tm.lap("sort pairs -> candidate CSR");
  return 0;
}

// ---- the two stages of a build ------------------------------------------------------------------------
struct BuildParams {
  int knn = 0, k1 = 0;
  double decay = 0, thresh_eff = 0, anisotropy = 0, bandwidth_scale = 1, radius_factor = 1;
  bool simt = false, keep_raw = false;
};

static int parse_build_params(const char *who, int64_t n, int64_t d, int knn, double decay, double thresh,
                              double anisotropy, double bandwidth_scale, int flags, BuildParams *bp) {
  MELD_REQUIRE(n >= 2 && d >= 1, "%s: bad data shape (%lld, %lld)", who, (long long)n, (long long)d);
  MELD_REQUIRE(n < (int64_t)2000000000, "%s: n too large for int32 columns", who);
  MELD_REQUIRE(knn >= 1 && (int64_t)knn + 1 <= n, "%s: knn=%d with n=%lld", who, knn, (long long)n);
  MELD_REQUIRE(knn + 1 <= kMaxK1, "%s: knn + 1 = %d exceeds the engine limit %d", who, knn + 1, kMaxK1);
  MELD_REQUIRE(decay >= 0 && isfinite(decay), "%s: decay=%g (pass 0 for the reference's decay=None)", who, decay);
  if (decay == 0.0) {  // binary kNN kernel: the bandwidth is the knn-th neighbour itself, nothing to threshold
    thresh = 0.5;
    bandwidth_scale = 1.0;
  }
  MELD_REQUIRE(thresh > 0 && thresh < 1, "%s: thresh=%g outside (0, 1)", who, thresh);
  MELD_REQUIRE(anisotropy >= 0 && anisotropy <= 1, "%s: anisotropy=%g outside [0, 1]", who, anisotropy);
  MELD_REQUIRE(bandwidth_scale > 0, "%s: bandwidth_scale=%g", who, bandwidth_scale);
  bp->knn = knn;
  bp->k1 = knn + 1;
  bp->decay = decay;
  bp->anisotropy = anisotropy;
  bp->bandwidth_scale = bandwidth_scale;
  bp->thresh_eff = thresh > DBL_EPSILON ? thresh : DBL_EPSILON;
  const double rho = decay == 0.0 ? 1.0 : pow(-log(bp->thresh_eff), 1.0 / decay);  // kernel radius in units of eps_i
  bp->radius_factor = rho * rho * bandwidth_scale * bandwidth_scale;
  if (bp->radius_factor < 1.0) bp->radius_factor = 1.0;  // the k1 nearest must be candidates to get eps_i
  bp->simt = (flags & MELD_B200_FLAG_SIMT_SEARCH) != 0;
  bp->keep_raw = (flags & MELD_B200_FLAG_KEEP_KNN_KERNEL) != 0;
  return 0;
}

// Stage 1 (row-local, shards over query rows with no exchange): candidate search + exact distances + eps
// for rows [row_begin, row_end) of X (already in internal cell order).
static int build_stage1(const double *X, int64_t n, int64_t d, const BuildParams &bp, int64_t row_begin,
                        int64_t row_end, const CellClusters *cl, cudaStream_t stream, Candidates &cs, DevBuf<double> &d2,
                        DevBuf<double> &eps) {
  StageTimer tm(stream);
  MELD_CHECK(candidate_search(X, n, d, bp.k1, bp.radius_factor, bp.simt, row_begin, row_end, cl, stream, cs));
  cs.key2.release();
  tm.lap("candidate search total");
  const int64_t nloc = row_end - row_begin;
  DevBuf<int> err;
  MELD_CHECK(d2.alloc((size_t)cs.total));
  MELD_CHECK(eps.alloc((size_t)nloc));
  MELD_CHECK(err.alloc(1));
  MELD_CUDA(cudaMemsetAsync(err.p, 0, sizeof(int), stream));
  MELD_CHECK(refine_distances(X, row_begin, nloc, d, cs.cand.p, cs.cptr.p, bp.k1, bp.bandwidth_scale, d2.p, eps.p, err.p,
                              stream));
  int h_err = 0;
  MELD_CUDA(cudaMemcpyAsync(&h_err, err.p, sizeof(int), cudaMemcpyDeviceToHost, stream));
  MELD_SYNC(stream);
  if (h_err) {
    set_error("knn graph build: a row has fewer than knn+1 candidates (candidate search failed)");
    return MELD_B200_ERR_INTERNAL;
  }
  tm.lap("refine (exact distances)");
  return 0;
}

// Stage 2 (needs every row's candidates and eps): kernel values, symmetrisation, sort, anisotropy, Laplacian.
// cand / vbuf are consumed (flags and values are written in place); perm is handed to the graph.
static int build_stage2(int64_t n, const int64_t *cptr, int32_t *cand, double *vbuf, int64_t total, const double *eps,
                        const BuildParams &bp, DevBuf<int32_t> &perm, cudaStream_t stream, meld_b200_graph **graph_out) {
  StageTimer tm(stream);
  DevBuf<double> kraw;
  DevBuf<int32_t> kept, extra, tot, max_extra;
  int32_t h_max_extra = 0;
  MELD_CHECK(kept.alloc((size_t)n));
  MELD_CHECK(extra.alloc((size_t)n));
  MELD_CHECK(tot.alloc((size_t)n + 1));
  MELD_CUDA(cudaMemsetAsync(extra.p, 0, (size_t)n * sizeof(int32_t), stream));
  if (bp.keep_raw) MELD_CHECK(kraw.alloc((size_t)total));
  kernel_values_kernel<<<warp_grid(n, 128), 128, 0, stream>>>(n, cand, cptr, vbuf, eps, bp.decay, bp.thresh_eff, kept.p,
                                                              extra.p, bp.keep_raw ? kraw.p : nullptr);
  MELD_LAUNCH_CHECK();
  add_counts_kernel<<<(unsigned)ceil_div(n + 1, 256), 256, 0, stream>>>(kept.p, extra.p, n, tot.p);
  MELD_LAUNCH_CHECK();

  meld_b200_graph *g = new (std::nothrow) meld_b200_graph();
  if (!g) {
    set_error("knn graph build: host allocation failed");
    return MELD_B200_ERR_NOMEM;
  }
  struct Guard {
    meld_b200_graph *g;
    ~Guard() { delete g; }
  } guard{g};
  g->n_rows = g->n_cols = n;
  g->row0 = 0;
  MELD_CHECK(g->row_ptr.alloc((size_t)n + 1 + kCsrPad));
  MELD_CUDA(cudaMemsetAsync(g->row_ptr.p + n + 1, 0, kCsrPad * sizeof(int32_t), stream));
  {
    size_t tmp_bytes = 0;
    MELD_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, tot.p, g->row_ptr.p, (int)(n + 1), stream));
    DevBuf<unsigned char> tmp;
    MELD_CHECK(tmp.alloc(tmp_bytes));
    MELD_CUDA(cub::DeviceScan::ExclusiveSum(tmp.p, tmp_bytes, tot.p, g->row_ptr.p, (int)(n + 1), stream));
    int32_t h_nnz = 0;
    MELD_CHECK(max_extra.alloc(1));
    MELD_CUDA(cudaMemsetAsync(max_extra.p, 0, sizeof(int32_t), stream));
    max_int_kernel<<<296, 256, 0, stream>>>(extra.p, n, max_extra.p);
    MELD_LAUNCH_CHECK();
    MELD_CUDA(cudaMemcpyAsync(&h_max_extra, max_extra.p, sizeof(int32_t), cudaMemcpyDeviceToHost, stream));
    MELD_CUDA(cudaMemcpyAsync(&h_nnz, g->row_ptr.p + n, sizeof(int32_t), cudaMemcpyDeviceToHost, stream));
    MELD_SYNC(stream);
    if (h_nnz < 0) {
      set_error("knn graph build: nnz overflows int32");
      return MELD_B200_ERR_UNSUPPORTED;
    }
    g->nnz = h_nnz;
  }
  {
    DevBuf<int32_t> ucol, cursor;
    DevBuf<double> uval;
    MELD_CHECK(ucol.alloc((size_t)g->nnz));
    MELD_CHECK(uval.alloc((size_t)g->nnz));
    MELD_CHECK(cursor.alloc((size_t)n));
    MELD_CUDA(cudaMemsetAsync(cursor.p, 0, (size_t)n * sizeof(int32_t), stream));
    fill_sym_kernel<<<warp_grid(n, 128), 128, 0, stream>>>(n, cand, cptr, vbuf, g->row_ptr.p, kept.p, cursor.p, ucol.p,
                                                           uval.p);
    MELD_LAUNCH_CHECK();
    MELD_CHECK(g->col.alloc((size_t)g->nnz + kCsrPad));
    MELD_CHECK(g->val.alloc((size_t)g->nnz + kCsrPad));
    MELD_CUDA(cudaMemsetAsync(g->col.p + g->nnz, 0, kCsrPad * sizeof(int32_t), stream));
    MELD_CUDA(cudaMemsetAsync(g->val.p + g->nnz, 0, kCsrPad * sizeof(double), stream));
    if (h_max_extra <= 2048 && tuning().merge_rows) {
      // few mirrored entries per row (the usual case): place every entry by rank
      merge_rows_kernel<<<warp_grid(n, 256), 256, 0, stream>>>(n, g->row_ptr.p, kept.p, ucol.p, uval.p, g->col.p,
                                                                g->val.p);
      MELD_LAUNCH_CHECK();
    } else {  // a hub row collected thousands of mirrored entries: sort the rows (library plumbing)
      size_t tmp_bytes = 0;
      MELD_CUDA(cub::DeviceSegmentedSort::SortPairs(nullptr, tmp_bytes, ucol.p, g->col.p, uval.p, g->val.p, (int)g->nnz,
                                                    (int)n, g->row_ptr.p, g->row_ptr.p + 1, stream));
      DevBuf<unsigned char> tmp;
      MELD_CHECK(tmp.alloc(tmp_bytes));
      MELD_CUDA(cub::DeviceSegmentedSort::SortPairs(tmp.p, tmp_bytes, ucol.p, g->col.p, uval.p, g->val.p,
                                                    (int)g->nnz, (int)n, g->row_ptr.p, g->row_ptr.p + 1, stream));
    }
    // no host wait: ucol / uval / cursor are arena bumps or stream-ordered pool frees
  }
  tm.lap("kernel values, fill, sort");

  if (bp.keep_raw) {  // export copy of the un-symmetrised kernel
    DevBuf<int64_t> rp64;
    MELD_CHECK(rp64.alloc((size_t)n + 1));
    MELD_CHECK(g->knn_cnt.alloc((size_t)n + 1));
    add_counts_kernel<<<(unsigned)ceil_div(n + 1, 256), 256, 0, stream>>>(kept.p, nullptr, n, g->knn_cnt.p);
    MELD_LAUNCH_CHECK();
    size_t tmp_bytes = 0;
    MELD_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, g->knn_cnt.p, rp64.p, (int)(n + 1), stream));
    DevBuf<unsigned char> tmp;
    MELD_CHECK(tmp.alloc(tmp_bytes));
    MELD_CUDA(cub::DeviceScan::ExclusiveSum(tmp.p, tmp_bytes, g->knn_cnt.p, rp64.p, (int)(n + 1), stream));
    int64_t h_raw = 0;
    MELD_CUDA(cudaMemcpyAsync(&h_raw, rp64.p + n, sizeof(int64_t), cudaMemcpyDeviceToHost, stream));
    MELD_SYNC(stream);
    g->knn_nnz = h_raw;
    MELD_CHECK(g->knn_col.alloc((size_t)h_raw));
    MELD_CHECK(g->knn_val.alloc((size_t)h_raw));
    MELD_CHECK(g->knn_ptr.alloc((size_t)n + 1));
    MELD_CUDA(cudaMemcpyAsync(g->knn_ptr.p, rp64.p, ((size_t)n + 1) * sizeof(int64_t), cudaMemcpyDeviceToDevice, stream));
    fill_raw_kernel<<<warp_grid(n, 128), 128, 0, stream>>>(n, cand, cptr, kraw.p, g->knn_ptr.p, g->knn_col.p,
                                                           g->knn_val.p);
    MELD_LAUNCH_CHECK();
    MELD_SYNC(stream);
  }

  {  // anisotropy + Laplacian
    DevBuf<double> q;
    MELD_CHECK(q.alloc((size_t)n));
    row_sum_kernel<<<warp_grid(n, 256), 256, 0, stream>>>(n, g->row_ptr.p, g->val.p, q.p);
    MELD_LAUNCH_CHECK();
    laplacian_kernel<<<warp_grid(n, 256), 256, 0, stream>>>(n, g->row_ptr.p, g->col.p, g->val.p, q.p, bp.anisotropy);
    MELD_LAUNCH_CHECK();
    MELD_SYNC(stream);
  }
  tm.lap("anisotropy + laplacian");
  MELD_CHECK(graph_finalize(g, stream));
  tm.lap("finalize (row-block partition)");
  if (perm.p) {  // the graph keeps its own copy (perm may live in the build arena)
    MELD_CHECK(g->perm.alloc(perm.n));
    MELD_CUDA(cudaMemcpyAsync(g->perm.p, perm.p, perm.n * sizeof(int32_t), cudaMemcpyDeviceToDevice, stream));
    MELD_SYNC(stream);
    perm.release();
  }
  g->stats[2] = total;
  g->stats[4] = bp.simt ? 1 : 0;
  guard.g = nullptr;
  *graph_out = g;
  return 0;
}

// Opaque result of stage 1 for a row range (include/meld_b200.h: meld_b200_cands_t).
struct meld_b200_cands {
  int64_t n = 0, d = 0, row_begin = 0, row_end = 0, total = 0;
  // persistent: these outlive the call that built them (temporaries of a build live in the arena)
  meld::DevBuf<int64_t> cptr{true};  // local rows + 1
  meld::DevBuf<int32_t> cand{true};  // total
  meld::DevBuf<double> d2{true};     // total
  meld::DevBuf<double> eps{true};    // local rows
  meld::DevBuf<int32_t> perm{true};  // n, or empty (identity)
  int32_t max_per_row = 0;
  int passes = 0;
  double times[8] = {0, 0, 0, 0, 0, 0, 0, 0};  // as meld_b200_graph::times, for the local rows
};

// ---- sharded stage 2: a rank assembles ONLY its rows of L ---------------------------------------------------------
// Inputs per rank: its stage-1 result (candidates + exact d^2 of rows [row0, row0 + nloc)), eps of ALL rows
// (all-gathered, 8 N bytes).  K_ij and K_ji follow from d_ij, eps_i, eps_j, so kernel values are row-local.  An entry
// (i, j) whose reverse edge is below threshold has to appear in row j as well: j local -> appended here, j remote -> a
// 16-byte record (j, i, v) for j's owner (all-to-all-v by the caller).  Anisotropy needs the row sums q of all rows
// (all-gathered, 8 N bytes).  Nothing O(nnz) is replicated or exchanged.
struct MirrorRec {
  int32_t j, i;
  double v;
};

namespace meld {

__global__ void kernel_values_local_kernel(int64_t nloc, int64_t row0, int32_t *__restrict__ cand,
                                           const int64_t *__restrict__ cptr, double *__restrict__ d2buf,
                                           const double *__restrict__ eps_full, double decay, double thresh,
                                           int32_t *__restrict__ kept, int32_t *__restrict__ extra) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t i = warp; i < nloc; i += nwarps) {
    const int c = (int)(cptr[i + 1] - cptr[i]);
    int32_t *ci = cand + cptr[i];
    double *di = d2buf + cptr[i];
    const int32_t ig = (int32_t)(row0 + i);
    const double ei = eps_full[ig];
    int nk = 0;
    for (int t = lane; t < c; t += 32) {
      const int32_t j = ci[t];
      const double dist = sqrt(di[t]);
      const double kij = alpha_decay(dist, ei, decay);
      if (kij >= thresh) {
        double kji = (j == ig) ? kij : alpha_decay(dist, eps_full[j], decay);
        if (kji < thresh) kji = 0.0;
        di[t] = (kij + kji) / 2;
        if (kji == 0.0) {
          ci[t] = ~j;  // mirrored entry needed in row j
          if (j >= row0 && j < row0 + nloc) atomicAdd(extra + (j - row0), 1);
        }
        ++nk;
      } else {
        ci[t] = INT32_MIN;  // dead slot
      }
    }
    for (int o = 16; o > 0; o >>= 1) nk += __shfl_xor_sync(0xffffffffu, nk, o);
    if (lane == 0) kept[i] = nk;
  }
}

__device__ __forceinline__ int owner_of(const int64_t *bounds, int world, int64_t j) {
  int w = 0;
  while (w + 1 < world && j >= bounds[w + 1]) ++w;
  return w;
}

// flagged slots whose row j lives on another rank: counted per owner (fill == 0) or written as records into the
// owner's region of the send buffer (fill == 1; cursor[w] starts at the region's offset)
__global__ void mirror_records_kernel(int64_t nloc, int64_t row0, const int32_t *__restrict__ cand,
                                      const int64_t *__restrict__ cptr, const double *__restrict__ vbuf,
                                      const int64_t *__restrict__ bounds, int world, int fill,
                                      unsigned long long *__restrict__ cursor, MirrorRec *__restrict__ out) {
  __shared__ int64_t sb[9];
  if (threadIdx.x <= world) sb[threadIdx.x] = bounds[threadIdx.x];
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t i = warp; i < nloc; i += nwarps) {
    const int c = (int)(cptr[i + 1] - cptr[i]);
    const int32_t *ci = cand + cptr[i];
    const double *vi = vbuf + cptr[i];
    for (int t = lane; t < c; t += 32) {
      const int32_t j = ci[t];
      if (j < 0 && j != INT32_MIN) {
        const int32_t jj = ~j;
        if (jj < row0 || jj >= row0 + nloc) {
          const int w = owner_of(sb, world, jj);
          const unsigned long long pos = atomicAdd(cursor + w, 1ull);
          if (fill) {
            MirrorRec r;
            r.j = jj;
            r.i = (int32_t)(row0 + i);
            r.v = vi[t];
            out[pos] = r;
          }
        }
      }
    }
  }
}

__global__ void count_received_kernel(const MirrorRec *__restrict__ rec, int64_t n_rec, int64_t row0, int64_t nloc,
                                      int32_t *__restrict__ extra, int *__restrict__ err) {
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < n_rec; t += (int64_t)gridDim.x * blockDim.x) {
    const int64_t j = rec[t].j - row0;
    if (j < 0 || j >= nloc) {
      atomicExch(err, 1);
      continue;
    }
    atomicAdd(extra + j, 1);
  }
}

// fill_sym_kernel for a row slice: columns are global, mirrored entries are only appended for local rows j
__global__ void fill_sym_local_kernel(int64_t nloc, int64_t row0, const int32_t *__restrict__ cand,
                                      const int64_t *__restrict__ cptr, const double *__restrict__ vbuf,
                                      const int32_t *__restrict__ row_ptr, const int32_t *__restrict__ kept,
                                      int32_t *__restrict__ cursor, int32_t *__restrict__ col, double *__restrict__ val) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t i = warp; i < nloc; i += nwarps) {
    const int c = (int)(cptr[i + 1] - cptr[i]);
    const int32_t *ci = cand + cptr[i];
    const double *vi = vbuf + cptr[i];
    int base = row_ptr[i];
    for (int t0 = 0; t0 < c; t0 += 32) {
      const int t = t0 + lane;
      int32_t j = INT32_MIN;
      double v = 0.0;
      if (t < c) {
        j = ci[t];
        v = vi[t];
      }
      const bool live = j != INT32_MIN;
      const unsigned m = __ballot_sync(0xffffffffu, live);
      if (live) {
        const bool flagged = j < 0;
        const int32_t jj = flagged ? ~j : j;
        const int pos = base + __popc(m & ((1u << lane) - 1u));
        col[pos] = jj;
        val[pos] = v;
        if (flagged && jj >= row0 && jj < row0 + nloc) {
          const int64_t jl = jj - row0;
          const int p2 = row_ptr[jl] + kept[jl] + atomicAdd(cursor + jl, 1);
          col[p2] = (int32_t)(row0 + i);
          val[p2] = v;
        }
      }
      base += __popc(m);
    }
  }
}

__global__ void append_received_kernel(const MirrorRec *__restrict__ rec, int64_t n_rec, int64_t row0,
                                       const int32_t *__restrict__ row_ptr, const int32_t *__restrict__ kept,
                                       int32_t *__restrict__ cursor, int32_t *__restrict__ col, double *__restrict__ val) {
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < n_rec; t += (int64_t)gridDim.x * blockDim.x) {
    const MirrorRec r = rec[t];
    const int64_t jl = r.j - row0;
    const int p2 = row_ptr[jl] + kept[jl] + atomicAdd(cursor + jl, 1);
    col[p2] = r.i;
    val[p2] = r.v;
  }
}

// laplacian_kernel for a row slice (q holds the row sums of ALL rows)
__global__ void laplacian_local_kernel(int64_t nloc, int64_t row0, const int32_t *__restrict__ row_ptr,
                                       const int32_t *__restrict__ col, double *__restrict__ val,
                                       const double *__restrict__ q, double anisotropy) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t i = warp; i < nloc; i += nwarps) {
    const int32_t ig = (int32_t)(row0 + i);
    const double qi = q[ig];
    double dw = 0.0;
    int diag = -1;
    for (int e = row_ptr[i] + lane; e < row_ptr[i + 1]; e += 32) {
      const int32_t j = col[e];
      if (j == ig) {
        diag = e;
        continue;
      }
      double w = val[e];
      if (anisotropy != 0.0) {
        const double qq = qi * q[j];
        w = (anisotropy == 1.0 ? 1.0 / qq : pow(qq, -anisotropy)) * w;
      }
      val[e] = -w;
      dw += w;
    }
    for (int o = 16; o > 0; o >>= 1) {
      dw += __shfl_xor_sync(0xffffffffu, dw, o);
      diag = max(diag, __shfl_xor_sync(0xffffffffu, diag, o));
    }
    if (lane == 0 && diag >= 0) val[diag] = dw;
  }
}

}  // namespace meld

// State of a sharded stage 2 between its three calls (the caller runs the collectives in between).
struct meld_b200_stage2 {
  meld_b200_cands *cands = nullptr;  // borrowed; its candidate / distance arrays are rewritten in place
  BuildParams bp;
  int64_t n = 0, row0 = 0, nloc = 0;
  int world = 1;
  int64_t bounds[9] = {0};
  meld::DevBuf<int64_t> d_bounds{true};
  meld::DevBuf<int32_t> kept{true}, extra{true};
  meld::DevBuf<unsigned long long> cursor{true};
  int64_t send_counts[8] = {0}, n_send = 0;
  meld_b200_graph *g = nullptr;  // under construction
  ~meld_b200_stage2() { delete g; }
};

extern "C" {

int meld_b200_stage2_begin(meld_b200_cands_t *c, const double *eps_full, const int64_t *bounds_host, int world, int knn,
                           double decay, double thresh, double anisotropy, double bandwidth_scale, void *stream_,
                           meld_b200_stage2_t **out, int64_t *send_counts_host) {
  cudaStream_t stream = (cudaStream_t)stream_;
  meld::use_stream(stream);
  MELD_REQUIRE(c && eps_full && bounds_host && out && send_counts_host, "stage2_begin: NULL argument");
  MELD_REQUIRE(world >= 1 && world <= 8, "stage2_begin: world=%d", world);
  *out = nullptr;
  meld_b200_stage2 *st = new (std::nothrow) meld_b200_stage2();
  if (!st) {
    set_error("stage2_begin: host allocation failed");
    return MELD_B200_ERR_NOMEM;
  }
  struct Guard {
    meld_b200_stage2 *s;
    ~Guard() { delete s; }
  } guard{st};
  st->cands = c;
  st->n = c->n;
  st->row0 = c->row_begin;
  st->nloc = c->row_end - c->row_begin;
  st->world = world;
  for (int w = 0; w <= world; ++w) st->bounds[w] = bounds_host[w];
  MELD_REQUIRE(st->bounds[0] == 0 && st->bounds[world] == c->n, "stage2_begin: bounds do not cover [0, n)");
  MELD_CHECK(parse_build_params("stage2_begin", c->n, 1, knn, decay, thresh, anisotropy, bandwidth_scale, 0, &st->bp));
  const int64_t nloc = st->nloc;
  MELD_CHECK(st->d_bounds.alloc(9));
  MELD_CUDA(cudaMemcpyAsync(st->d_bounds.p, st->bounds, 9 * sizeof(int64_t), cudaMemcpyHostToDevice, stream));
  MELD_CHECK(st->kept.alloc((size_t)nloc + 1));
  MELD_CHECK(st->extra.alloc((size_t)nloc + 1));
  MELD_CHECK(st->cursor.alloc(8));
  MELD_CUDA(cudaMemsetAsync(st->extra.p, 0, ((size_t)nloc + 1) * sizeof(int32_t), stream));
  MELD_CUDA(cudaMemsetAsync(st->cursor.p, 0, 8 * sizeof(unsigned long long), stream));
  kernel_values_local_kernel<<<warp_grid(nloc, 128), 128, 0, stream>>>(nloc, st->row0, c->cand.p, c->cptr.p, c->d2.p,
                                                                      eps_full, st->bp.decay, st->bp.thresh_eff,
                                                                      st->kept.p, st->extra.p);
  MELD_LAUNCH_CHECK();
  mirror_records_kernel<<<warp_grid(nloc, 256), 256, 0, stream>>>(nloc, st->row0, c->cand.p, c->cptr.p, c->d2.p,
                                                                 st->d_bounds.p, world, 0, st->cursor.p, nullptr);
  MELD_LAUNCH_CHECK();
  unsigned long long h_cnt[8] = {0};
  MELD_CUDA(cudaMemcpyAsync(h_cnt, st->cursor.p, 8 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, stream));
  MELD_SYNC(stream);
  st->n_send = 0;
  for (int w = 0; w < world; ++w) {
    st->send_counts[w] = (int64_t)h_cnt[w];
    st->n_send += (int64_t)h_cnt[w];
    send_counts_host[w] = (int64_t)h_cnt[w];
  }
  guard.s = nullptr;
  *out = st;
  return 0;
}

int meld_b200_stage2_records(meld_b200_stage2_t *st, void *records_out, void *stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  meld::use_stream(stream);
  MELD_REQUIRE(st && st->cands && (st->n_send == 0 || records_out), "stage2_records: NULL argument");
  if (st->n_send == 0) return 0;
  unsigned long long off[8] = {0};
  int64_t run = 0;
  for (int w = 0; w < st->world; ++w) {
    off[w] = (unsigned long long)run;
    run += st->send_counts[w];
  }
  MELD_CUDA(cudaMemcpyAsync(st->cursor.p, off, 8 * sizeof(unsigned long long), cudaMemcpyHostToDevice, stream));
  MELD_SYNC(stream);  // off is a local
  meld_b200_cands *c = st->cands;
  mirror_records_kernel<<<warp_grid(st->nloc, 256), 256, 0, stream>>>(st->nloc, st->row0, c->cand.p, c->cptr.p, c->d2.p,
                                                                     st->d_bounds.p, st->world, 1, st->cursor.p,
                                                                     static_cast<MirrorRec *>(records_out));
  MELD_LAUNCH_CHECK();
  return 0;
}

int meld_b200_stage2_assemble(meld_b200_stage2_t *st, const void *recv_records, int64_t n_recv, void *stream_,
                              double *q_local_out) {
  cudaStream_t stream = (cudaStream_t)stream_;
  meld::use_stream(stream);
  MELD_REQUIRE(st && st->cands && q_local_out && (n_recv == 0 || recv_records), "stage2_assemble: NULL argument");
  const MirrorRec *rec = static_cast<const MirrorRec *>(recv_records);
  meld_b200_cands *c = st->cands;
  const int64_t nloc = st->nloc, row0 = st->row0;
  DevBuf<int> err{true};
  MELD_CHECK(err.alloc(1));
  MELD_CUDA(cudaMemsetAsync(err.p, 0, sizeof(int), stream));
  if (n_recv > 0) {
    count_received_kernel<<<grid_for_rows(n_recv), 256, 0, stream>>>(rec, n_recv, row0, nloc, st->extra.p, err.p);
    MELD_LAUNCH_CHECK();
  }
  DevBuf<int32_t> tot{true}, cur32{true}, max_extra{true};
  MELD_CHECK(tot.alloc((size_t)nloc + 1));
  add_counts_kernel<<<(unsigned)ceil_div(nloc + 1, 256), 256, 0, stream>>>(st->kept.p, st->extra.p, nloc, tot.p);
  MELD_LAUNCH_CHECK();
  meld_b200_graph *g = new (std::nothrow) meld_b200_graph();
  if (!g) {
    set_error("stage2_assemble: host allocation failed");
    return MELD_B200_ERR_NOMEM;
  }
  delete st->g;
  st->g = g;
  g->n_rows = nloc;
  g->n_cols = st->n;
  g->row0 = row0;
  MELD_CHECK(g->row_ptr.alloc((size_t)nloc + 1 + kCsrPad));
  MELD_CUDA(cudaMemsetAsync(g->row_ptr.p + nloc + 1, 0, kCsrPad * sizeof(int32_t), stream));
  int32_t h_nnz = 0, h_max_extra = 0;
  int h_err = 0;
  {
    size_t tmp_bytes = 0;
    MELD_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, tot.p, g->row_ptr.p, (int)(nloc + 1), stream));
    DevBuf<unsigned char> tmp{true};
    MELD_CHECK(tmp.alloc(tmp_bytes));
    MELD_CUDA(cub::DeviceScan::ExclusiveSum(tmp.p, tmp_bytes, tot.p, g->row_ptr.p, (int)(nloc + 1), stream));
    MELD_CHECK(max_extra.alloc(1));
    MELD_CUDA(cudaMemsetAsync(max_extra.p, 0, sizeof(int32_t), stream));
    max_int_kernel<<<296, 256, 0, stream>>>(st->extra.p, nloc, max_extra.p);
    MELD_LAUNCH_CHECK();
    MELD_CUDA(cudaMemcpyAsync(&h_max_extra, max_extra.p, sizeof(int32_t), cudaMemcpyDeviceToHost, stream));
    MELD_CUDA(cudaMemcpyAsync(&h_nnz, g->row_ptr.p + nloc, sizeof(int32_t), cudaMemcpyDeviceToHost, stream));
    MELD_CUDA(cudaMemcpyAsync(&h_err, err.p, sizeof(int), cudaMemcpyDeviceToHost, stream));
    MELD_SYNC(stream);
  }
  MELD_REQUIRE(h_err == 0, "stage2_assemble: a received mirror record does not belong to this rank's rows");
  MELD_REQUIRE(h_nnz >= 0, "stage2_assemble: nnz overflows int32");
  g->nnz = h_nnz;
  {
    DevBuf<int32_t> ucol{true};
    DevBuf<double> uval{true};
    MELD_CHECK(ucol.alloc((size_t)g->nnz + 1));
    MELD_CHECK(uval.alloc((size_t)g->nnz + 1));
    MELD_CHECK(cur32.alloc((size_t)nloc + 1));
    MELD_CUDA(cudaMemsetAsync(cur32.p, 0, ((size_t)nloc + 1) * sizeof(int32_t), stream));
    fill_sym_local_kernel<<<warp_grid(nloc, 128), 128, 0, stream>>>(nloc, row0, c->cand.p, c->cptr.p, c->d2.p,
                                                                   g->row_ptr.p, st->kept.p, cur32.p, ucol.p, uval.p);
    MELD_LAUNCH_CHECK();
    if (n_recv > 0) {
      append_received_kernel<<<grid_for_rows(n_recv), 256, 0, stream>>>(rec, n_recv, row0, g->row_ptr.p, st->kept.p,
                                                                       cur32.p, ucol.p, uval.p);
      MELD_LAUNCH_CHECK();
    }
    MELD_CHECK(g->col.alloc((size_t)g->nnz + kCsrPad));
    MELD_CHECK(g->val.alloc((size_t)g->nnz + kCsrPad));
    MELD_CUDA(cudaMemsetAsync(g->col.p + g->nnz, 0, kCsrPad * sizeof(int32_t), stream));
    MELD_CUDA(cudaMemsetAsync(g->val.p + g->nnz, 0, kCsrPad * sizeof(double), stream));
    if (h_max_extra <= 2048 && tuning().merge_rows) {
      merge_rows_kernel<<<warp_grid(nloc, 256), 256, 0, stream>>>(nloc, g->row_ptr.p, st->kept.p, ucol.p, uval.p,
                                                                   g->col.p, g->val.p);
      MELD_LAUNCH_CHECK();
    } else {
      size_t tmp_bytes = 0;
      MELD_CUDA(cub::DeviceSegmentedSort::SortPairs(nullptr, tmp_bytes, ucol.p, g->col.p, uval.p, g->val.p, (int)g->nnz,
                                                    (int)nloc, g->row_ptr.p, g->row_ptr.p + 1, stream));
      DevBuf<unsigned char> tmp{true};
      MELD_CHECK(tmp.alloc(tmp_bytes));
      MELD_CUDA(cub::DeviceSegmentedSort::SortPairs(tmp.p, tmp_bytes, ucol.p, g->col.p, uval.p, g->val.p, (int)g->nnz,
                                                    (int)nloc, g->row_ptr.p, g->row_ptr.p + 1, stream));
    }
    row_sum_kernel<<<warp_grid(nloc, 256), 256, 0, stream>>>(nloc, g->row_ptr.p, g->val.p, q_local_out);
    MELD_LAUNCH_CHECK();
  }
  return 0;
}

int meld_b200_stage2_finish(meld_b200_stage2_t *st, const double *q_full, void *stream_, meld_b200_graph_t **graph_out) {
  cudaStream_t stream = (cudaStream_t)stream_;
  meld::use_stream(stream);
  MELD_REQUIRE(st && st->g && q_full && graph_out, "stage2_finish: NULL argument (stage2_assemble first)");
  meld_b200_graph *g = st->g;
  laplacian_local_kernel<<<warp_grid(st->nloc, 256), 256, 0, stream>>>(st->nloc, st->row0, g->row_ptr.p, g->col.p,
                                                                      g->val.p, q_full, st->bp.anisotropy);
  MELD_LAUNCH_CHECK();
  MELD_CHECK(graph_finalize(g, stream));
  if (st->cands->perm.p) {
    MELD_CHECK(g->perm.alloc((size_t)st->n));
    MELD_CUDA(cudaMemcpyAsync(g->perm.p, st->cands->perm.p, (size_t)st->n * sizeof(int32_t), cudaMemcpyDeviceToDevice,
                              stream));
  }
  g->stats[0] = st->cands->passes;
  g->stats[1] = st->cands->max_per_row;
  g->stats[2] = st->cands->total;
  for (int i = 0; i < 8; ++i) g->times[i] = st->cands->times[i];
  st->g = nullptr;
  *graph_out = g;
  return 0;
}

int meld_b200_stage2_destroy(meld_b200_stage2_t *st) {
  if (st) {
    cudaDeviceSynchronize();
    meld::use_stream(nullptr);
  }
  delete st;
  return 0;
}

}  // extern "C"

extern "C" {

int meld_b200_knn_graph_build(const double *X, int64_t n, int64_t d, int knn, double decay, double thresh,
                              double anisotropy, double bandwidth_scale, int flags, void *stream_,
                              meld_b200_graph_t **graph_out) {
  cudaStream_t stream = (cudaStream_t)stream_;
  meld::use_stream(stream);
  meld::ArenaScope arena_scope(stream);
  MELD_REQUIRE(graph_out != nullptr, "knn_graph_build: graph_out is NULL");
  *graph_out = nullptr;
  MELD_REQUIRE(X != nullptr, "knn_graph_build: X is NULL");
  BuildParams bp;
  MELD_CHECK(parse_build_params("knn_graph_build", n, d, knn, decay, thresh, anisotropy, bandwidth_scale, flags, &bp));
  StageTimer tm(stream);
  // internal cell order (Morton curve over the highest-variance features)
  DevBuf<double> Xp;
  DevBuf<int32_t> perm;
  CellClusters cl;
  if (tuning().reorder && n >= tuning().reorder_min_n) {
    MELD_CHECK(cell_order(X, n, d, stream, perm, Xp, cl));
    X = Xp.p;
  }
  tm.lap("cell order (k-means + morton)");
  Candidates cs;
  DevBuf<double> d2, eps;
  MELD_CHECK(build_stage1(X, n, d, bp, 0, n, &cl, stream, cs, d2, eps));
  Xp.release();
  meld_b200_graph *g = nullptr;
  MELD_CHECK(build_stage2(n, cs.cptr.p, cs.cand.p, d2.p, cs.total, eps.p, bp, perm, stream, &g));
  g->stats[0] = cs.passes;
  g->stats[1] = cs.max_per_row;
  g->stats[3] = cs.retries;
  g->times[0] = cs.pass1_ms;
  g->times[1] = cs.pass2_ms;
  g->times[2] = cs.gemm_flops_per_pass;
  g->times[3] = cs.gemm_flops_pass1;
  g->times[4] = cs.tile_frac > 0 ? cs.gemm_flops_per_pass / cs.tile_frac : cs.gemm_flops_per_pass;
  *graph_out = g;
  return 0;
}

// every cell is a candidate of every row (dense graphs): cand[i * n + j] = j
__global__ void dense_candidates_kernel(int64_t n, int64_t *__restrict__ cptr, int32_t *__restrict__ cand) {
  const int64_t total = n * n;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x)
    cand[t] = (int32_t)(t % n);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i <= n; i += (int64_t)gridDim.x * blockDim.x)
    cptr[i] = i * n;
}

int meld_b200_dense_graph_build(const double *X, int64_t n, int64_t d, int knn, double decay, double anisotropy,
                                double bandwidth_scale, int flags, void *stream_, meld_b200_graph_t **graph_out) {
  cudaStream_t stream = (cudaStream_t)stream_;
  meld::use_stream(stream);
  meld::ArenaScope arena_scope(stream);
  MELD_REQUIRE(graph_out != nullptr && X != nullptr, "dense_graph_build: NULL argument");
  *graph_out = nullptr;
  MELD_REQUIRE(n <= 16384, "dense_graph_build: n=%lld; the dense exact graph (thresh = 0) is for n <= 16384",
               (long long)n);
  MELD_REQUIRE(decay > 0, "dense_graph_build: decay=%g", decay);
  BuildParams bp;
  MELD_CHECK(parse_build_params("dense_graph_build", n, d, knn, decay, 0.5, anisotropy, bandwidth_scale, flags, &bp));
  bp.thresh_eff = 4.9406564584124654e-324;  // thresh = 0: every pair whose kernel value has not underflowed to 0
  DevBuf<int64_t> cptr;
  DevBuf<int32_t> cand, noperm;
  DevBuf<double> d2, eps;
  DevBuf<int> err;
  MELD_CHECK(cptr.alloc((size_t)n + 1));
  MELD_CHECK(cand.alloc((size_t)n * n));
  MELD_CHECK(d2.alloc((size_t)n * n));
  MELD_CHECK(eps.alloc((size_t)n));
  MELD_CHECK(err.alloc(1));
  MELD_CUDA(cudaMemsetAsync(err.p, 0, sizeof(int), stream));
  dense_candidates_kernel<<<sm_count() * 8, 256, 0, stream>>>(n, cptr.p, cand.p);
  MELD_LAUNCH_CHECK();
  MELD_CHECK(refine_distances(X, 0, n, d, cand.p, cptr.p, bp.k1, bp.bandwidth_scale, d2.p, eps.p, err.p, stream));
  meld_b200_graph *g = nullptr;
  MELD_CHECK(build_stage2(n, cptr.p, cand.p, d2.p, n * n, eps.p, bp, noperm, stream, &g));
  g->stats[0] = 0;  // no search passes
  g->stats[1] = n;
  *graph_out = g;
  return 0;
}

int meld_b200_knn_candidates(const double *X, int64_t n, int64_t d, int knn, double decay, double thresh,
                             double bandwidth_scale, int64_t row_begin, int64_t row_end, int flags, void *stream_,
                             meld_b200_cands_t **cands_out) {
  cudaStream_t stream = (cudaStream_t)stream_;
  meld::use_stream(stream);
  meld::ArenaScope arena_scope(stream);
  MELD_REQUIRE(cands_out != nullptr && X != nullptr, "knn_candidates: NULL argument");
  *cands_out = nullptr;
  BuildParams bp;
  MELD_CHECK(parse_build_params("knn_candidates", n, d, knn, decay, thresh, 1.0, bandwidth_scale, flags, &bp));
  MELD_REQUIRE(row_begin >= 0 && row_begin < row_end && row_end <= n, "knn_candidates: bad row range [%lld, %lld)",
               (long long)row_begin, (long long)row_end);
  meld_b200_cands *c = new (std::nothrow) meld_b200_cands();
  if (!c) {
    set_error("knn_candidates: host allocation failed");
    return MELD_B200_ERR_NOMEM;
  }
  struct Guard {
    meld_b200_cands *c;
    ~Guard() { delete c; }
  } guard{c};
  c->n = n;
  c->d = d;
  c->row_begin = row_begin;
  c->row_end = row_end;
  DevBuf<double> Xp;
  CellClusters cl;
  if (tuning().reorder && n >= tuning().reorder_min_n) {  // every rank derives the same order from the same data
    MELD_CHECK(cell_order(X, n, d, stream, c->perm, Xp, cl));
    X = Xp.p;
  }
  Candidates cs;
  MELD_CHECK(build_stage1(X, n, d, bp, row_begin, row_end, &cl, stream, cs, c->d2, c->eps));
  c->total = cs.total;
  c->max_per_row = cs.max_per_row;
  c->passes = cs.passes;
  c->times[0] = cs.pass1_ms;
  c->times[1] = cs.pass2_ms;
  c->times[2] = cs.gemm_flops_per_pass;
  c->times[3] = cs.gemm_flops_pass1;
  c->times[4] = cs.tile_frac > 0 ? cs.gemm_flops_per_pass / cs.tile_frac : cs.gemm_flops_per_pass;
  // the handle keeps its own copies (the search's buffers live in the build arena)
  MELD_CHECK(c->cptr.alloc(cs.cptr.n));
  MELD_CHECK(c->cand.alloc(cs.cand.n));
  MELD_CUDA(cudaMemcpyAsync(c->cptr.p, cs.cptr.p, cs.cptr.n * sizeof(int64_t), cudaMemcpyDeviceToDevice, stream));
  MELD_CUDA(cudaMemcpyAsync(c->cand.p, cs.cand.p, cs.cand.n * sizeof(int32_t), cudaMemcpyDeviceToDevice, stream));
  MELD_SYNC(stream);
  guard.c = nullptr;
  *cands_out = c;
  return 0;
}

int meld_b200_cands_info(const meld_b200_cands_t *c, int64_t *n_rows_host, int64_t *total_host, int *has_perm_host,
                         int64_t *max_per_row_host) {
  MELD_REQUIRE(c != nullptr, "cands_info: NULL handle");
  if (n_rows_host) *n_rows_host = c->row_end - c->row_begin;
  if (total_host) *total_host = c->total;
  if (has_perm_host) *has_perm_host = c->perm.p ? 1 : 0;
  if (max_per_row_host) *max_per_row_host = c->max_per_row;
  return 0;
}

int meld_b200_cands_export(const meld_b200_cands_t *c, int64_t *counts, int32_t *cand, double *d2, double *eps,
                           int32_t *perm, void *stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  meld::use_stream(stream);
  MELD_REQUIRE(c != nullptr, "cands_export: NULL handle");
  const int64_t nloc = c->row_end - c->row_begin;
  if (counts) {
    row_counts64_kernel<<<(unsigned)ceil_div(nloc, 256), 256, 0, stream>>>(c->cptr.p, nloc, counts);
    MELD_LAUNCH_CHECK();
  }
  if (cand && c->total)
    MELD_CUDA(cudaMemcpyAsync(cand, c->cand.p, (size_t)c->total * sizeof(int32_t), cudaMemcpyDeviceToDevice, stream));
  if (d2 && c->total)
    MELD_CUDA(cudaMemcpyAsync(d2, c->d2.p, (size_t)c->total * sizeof(double), cudaMemcpyDeviceToDevice, stream));
  if (eps) MELD_CUDA(cudaMemcpyAsync(eps, c->eps.p, (size_t)nloc * sizeof(double), cudaMemcpyDeviceToDevice, stream));
  if (perm && c->perm.p)
    MELD_CUDA(cudaMemcpyAsync(perm, c->perm.p, (size_t)c->n * sizeof(int32_t), cudaMemcpyDeviceToDevice, stream));
  return 0;
}

int meld_b200_cands_destroy(meld_b200_cands_t *c) {
  if (c) {
    cudaDeviceSynchronize();
    meld::use_stream(nullptr);  // free on the legacy default stream
  }
  delete c;
  return 0;
}

int meld_b200_graph_from_candidates(int64_t n, const int64_t *counts, const int32_t *cand, const double *d2,
                                    int64_t total, const double *eps, const int32_t *perm, int knn, double decay,
                                    double thresh, double anisotropy, double bandwidth_scale, int flags, void *stream_,
                                    meld_b200_graph_t **graph_out) {
  cudaStream_t stream = (cudaStream_t)stream_;
  meld::use_stream(stream);
  meld::ArenaScope arena_scope(stream);
  MELD_REQUIRE(graph_out && counts && cand && d2 && eps, "graph_from_candidates: NULL argument");
  *graph_out = nullptr;
  BuildParams bp;
  MELD_CHECK(parse_build_params("graph_from_candidates", n, 1, knn, decay, thresh, anisotropy, bandwidth_scale, flags,
                                &bp));
  MELD_REQUIRE(total >= n && total < (int64_t)2147483647, "graph_from_candidates: total=%lld", (long long)total);
  // private, mutable copies (stage 2 rewrites candidates and distances in place)
  DevBuf<int64_t> cptr;
  DevBuf<int32_t> ccand, cperm;
  DevBuf<double> vbuf;
  MELD_CHECK(cptr.alloc((size_t)n + 1));
  MELD_CHECK(ccand.alloc((size_t)total));
  MELD_CHECK(vbuf.alloc((size_t)total));
  {
    size_t tmp_bytes = 0;
    DevBuf<int64_t> cnt1;
    MELD_CHECK(cnt1.alloc((size_t)n + 1));
    MELD_CUDA(cudaMemcpyAsync(cnt1.p, counts, (size_t)n * sizeof(int64_t), cudaMemcpyDeviceToDevice, stream));
    MELD_CUDA(cudaMemsetAsync(cnt1.p + n, 0, sizeof(int64_t), stream));
    MELD_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, cnt1.p, cptr.p, (int)(n + 1), stream));
    DevBuf<unsigned char> tmp;
    MELD_CHECK(tmp.alloc(tmp_bytes));
    MELD_CUDA(cub::DeviceScan::ExclusiveSum(tmp.p, tmp_bytes, cnt1.p, cptr.p, (int)(n + 1), stream));
    int64_t h_total = 0;
    MELD_CUDA(cudaMemcpyAsync(&h_total, cptr.p + n, sizeof(int64_t), cudaMemcpyDeviceToHost, stream));
    MELD_SYNC(stream);
    MELD_REQUIRE(h_total == total, "graph_from_candidates: counts sum to %lld, total says %lld", (long long)h_total,
                 (long long)total);
  }
  MELD_CUDA(cudaMemcpyAsync(ccand.p, cand, (size_t)total * sizeof(int32_t), cudaMemcpyDeviceToDevice, stream));
  MELD_CUDA(cudaMemcpyAsync(vbuf.p, d2, (size_t)total * sizeof(double), cudaMemcpyDeviceToDevice, stream));
  if (perm) {
    MELD_CHECK(cperm.alloc((size_t)n));
    MELD_CUDA(cudaMemcpyAsync(cperm.p, perm, (size_t)n * sizeof(int32_t), cudaMemcpyDeviceToDevice, stream));
  }
  meld_b200_graph *g = nullptr;
  MELD_CHECK(build_stage2(n, cptr.p, ccand.p, vbuf.p, total, eps, bp, cperm, stream, &g));
  g->stats[0] = 2;
  *graph_out = g;
  return 0;
}

int meld_b200_debug_candidate_search(const double *X, int64_t n, int64_t d, int knn, double decay, double thresh,
                                     double bandwidth_scale, int flags, void *stream_, float *key2_out,
                                     int32_t *cnt_out, int64_t *cap_host) {
  cudaStream_t stream = (cudaStream_t)stream_;
  meld::use_stream(stream);
  meld::ArenaScope arena_scope(stream);
  MELD_REQUIRE(X && key2_out && cnt_out && n >= 2 && d >= 1 && knn >= 1 && knn + 1 <= n && knn + 1 <= kMaxK1,
               "debug_candidate_search: bad argument");
  const double thresh_eff = thresh > DBL_EPSILON ? thresh : DBL_EPSILON;
  const double rho = decay == 0.0 ? 1.0 : pow(-log(thresh_eff), 1.0 / decay);
  double radius_factor = rho * rho * bandwidth_scale * bandwidth_scale;
  if (radius_factor < 1.0) radius_factor = 1.0;
  Candidates cs;
  MELD_CHECK(candidate_search(X, n, d, knn + 1, radius_factor, (flags & MELD_B200_FLAG_SIMT_SEARCH) != 0, 0, n, nullptr,
                              stream, cs));
  MELD_CUDA(cudaMemcpyAsync(key2_out, cs.key2.p, (size_t)n * sizeof(float), cudaMemcpyDeviceToDevice, stream));
  row_counts_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, stream>>>(cs.cptr.p, n, cnt_out);
  MELD_LAUNCH_CHECK();
  MELD_SYNC(stream);
  if (cap_host) *cap_host = cs.total;
  return 0;
}

int meld_b200_graph_build_times(const meld_b200_graph_t *g, double *times8_host) {
  MELD_REQUIRE(g && times8_host, "graph_build_times: NULL argument");
  for (int i = 0; i < 8; ++i) times8_host[i] = g->times[i];
  return 0;
}

int meld_b200_graph_permutation(const meld_b200_graph_t *g, int32_t *perm_out, int *is_identity_host, void *stream_) {
  MELD_REQUIRE(g && is_identity_host, "graph_permutation: NULL argument");
  *is_identity_host = g->perm.p ? 0 : 1;
  if (g->perm.p && perm_out)  // n_cols entries: a row slice keeps the whole operator's cell order
    MELD_CUDA(cudaMemcpyAsync(perm_out, g->perm.p, (size_t)g->n_cols * sizeof(int32_t), cudaMemcpyDeviceToDevice,
                              (cudaStream_t)stream_));
  return 0;
}

int meld_b200_graph_knn_kernel_nnz(const meld_b200_graph_t *g, int64_t *nnz_host) {
  MELD_REQUIRE(g && nnz_host, "graph_knn_kernel_nnz: NULL argument");
  MELD_REQUIRE(g->knn_nnz >= 0, "graph_knn_kernel_nnz: graph was built without MELD_B200_FLAG_KEEP_KNN_KERNEL");
  *nnz_host = g->knn_nnz;
  return 0;
}

int meld_b200_graph_export_knn_kernel(const meld_b200_graph_t *g, int64_t *indptr, int32_t *indices, double *data,
                                      void *stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  meld::use_stream(stream);
  MELD_REQUIRE(g && indptr && indices && data, "graph_export_knn_kernel: NULL argument");
  MELD_REQUIRE(g->knn_nnz >= 0, "graph_export_knn_kernel: graph was built without MELD_B200_FLAG_KEEP_KNN_KERNEL");
  MELD_CUDA(cudaMemcpyAsync(indptr, g->knn_ptr.p, ((size_t)g->n_rows + 1) * sizeof(int64_t), cudaMemcpyDeviceToDevice,
                            stream));
  if (g->knn_nnz > 0) {
    MELD_CUDA(cudaMemcpyAsync(indices, g->knn_col.p, (size_t)g->knn_nnz * sizeof(int32_t), cudaMemcpyDeviceToDevice,
                              stream));
    MELD_CUDA(cudaMemcpyAsync(data, g->knn_val.p, (size_t)g->knn_nnz * sizeof(double), cudaMemcpyDeviceToDevice,
                              stream));
  }
  return 0;
}

}  // extern "C"
