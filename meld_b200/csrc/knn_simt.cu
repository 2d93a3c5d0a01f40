// SIMT float32 candidate search: the cross-check implementation of knn_search.cuh (test flag
// MELD_B200_FLAG_SIMT_SEARCH), plus the dispatch between it and the tcgen05 path.
#include "knn_search.cuh"

namespace meld {

constexpr int kRows = 128;  // rows per CTA = threads per CTA (one row per thread)
constexpr int kCols = 32;   // columns per tile
constexpr int kChunk = 32;  // features per shared-memory chunk
constexpr int kYPitch = 36; // padded feature pitch of the column tile (16-byte aligned rows)

__global__ void simt_prep_kernel(const double *__restrict__ X, const double *__restrict__ mu,
                                 const double *__restrict__ norm, int64_t n, int64_t d, float *__restrict__ xc,
                                 float *__restrict__ hn) {
  const int64_t total = n * d;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
    const int64_t k = t % d;
    xc[t] = (float)(X[t] - mu[k]);
  }
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    hn[i] = (float)(-0.5 * norm[i]);
}

template <int MODE>
__global__ void __launch_bounds__(kRows) simt_search_kernel(const float *__restrict__ xc, const float *__restrict__ hn,
                                                           int64_t n, int64_t d, int64_t row_begin, int64_t row_end,
                                                           int k1, int nseg,
                                                           float *__restrict__ lists, const float *__restrict__ key2,
                                                           unsigned long long *__restrict__ pairs,
                                                           unsigned long long *__restrict__ count, int64_t cap) {
  extern __shared__ __align__(16) float sm[];
  float *Xs = sm;                           // [kChunk][kRows + 1]
  float *Ys = Xs + kChunk * (kRows + 1);    // [kCols][kYPitch]
  float *lst = Ys + kCols * kYPitch;        // [k1][kRows] (MODE 1)
  const int t = threadIdx.x, lane = t & 31, w = t >> 5;
  const int64_t row0 = row_begin + (int64_t)blockIdx.x * kRows;
  const int64_t row = row0 + t;
  const int seg = blockIdx.y;
  const int64_t ntile = (n + kCols - 1) / kCols;
  const int64_t tile_lo = ntile * seg / nseg, tile_hi = ntile * (seg + 1) / nseg;
  const int nchunk = (int)((d + kChunk - 1) / kChunk);
  float thr = -INFINITY;
  if (MODE == 1) {
    for (int s = 0; s < k1; ++s) lst[s * kRows + t] = -INFINITY;
  } else {
    thr = row < row_end ? key2[row] : INFINITY;
  }
  for (int64_t tile = tile_lo; tile < tile_hi; ++tile) {
    const int64_t col0 = tile * kCols;
    float acc[kCols];
#pragma unroll
    for (int j = 0; j < kCols; ++j) acc[j] = 0.f;
    for (int kc = 0; kc < nchunk; ++kc) {
      const int64_t k0 = (int64_t)kc * kChunk;
      __syncthreads();
      for (int r = 0; r < 32; ++r) {  // warp w stages rows w*32 .. w*32+31, lane = feature
        const int64_t rr = row0 + w * 32 + r;
        const int64_t kk = k0 + lane;
        Xs[lane * (kRows + 1) + w * 32 + r] = (rr < row_end && kk < d) ? xc[rr * d + kk] : 0.f;
      }
      for (int q = t; q < kCols * kChunk; q += kRows) {
        const int c = q / kChunk, k = q % kChunk;
        const int64_t cc = col0 + c, kk = k0 + k;
        Ys[c * kYPitch + k] = (cc < n && kk < d) ? xc[cc * d + kk] : 0.f;
      }
      __syncthreads();
#pragma unroll
      for (int k4 = 0; k4 < kChunk / 4; ++k4) {
        const float x0 = Xs[(4 * k4 + 0) * (kRows + 1) + t], x1 = Xs[(4 * k4 + 1) * (kRows + 1) + t];
        const float x2 = Xs[(4 * k4 + 2) * (kRows + 1) + t], x3 = Xs[(4 * k4 + 3) * (kRows + 1) + t];
#pragma unroll
        for (int j = 0; j < kCols; ++j) {
          const float4 y = *reinterpret_cast<const float4 *>(&Ys[j * kYPitch + 4 * k4]);
          acc[j] = fmaf(x0, y.x, acc[j]);
          acc[j] = fmaf(x1, y.y, acc[j]);
          acc[j] = fmaf(x2, y.z, acc[j]);
          acc[j] = fmaf(x3, y.w, acc[j]);
        }
      }
    }
    if (row < row_end) {
#pragma unroll
      for (int j = 0; j < kCols; ++j) {
        const int64_t col = col0 + j;
        if (col >= n) continue;
        const float s = acc[j] + hn[col];
        if (MODE == 1) {
          if (s > thr) {
            int pos = k1 - 1;
            while (pos > 0 && lst[(pos - 1) * kRows + t] < s) {
              lst[pos * kRows + t] = lst[(pos - 1) * kRows + t];
              --pos;
            }
            lst[pos * kRows + t] = s;
            thr = lst[(k1 - 1) * kRows + t];
          }
        } else {
          if (s >= thr) {
            const unsigned long long pos = atomicAdd(count, 1ull);
            if ((int64_t)pos < cap) pairs[pos] = ((unsigned long long)row << 32) | (unsigned long long)col;
          }
        }
      }
    }
  }
  if (MODE == 1 && row < row_end) {
    float *out = lists + ((size_t)row * nseg + seg) * k1;
    for (int s = 0; s < k1; ++s) out[s] = lst[s * kRows + t];
  }
}

static size_t simt_smem(int k1) {
  return sizeof(float) * ((size_t)kChunk * (kRows + 1) + (size_t)kCols * kYPitch + (size_t)k1 * kRows);
}

int search_plan(bool simt, int64_t n, int64_t d, int k1, int64_t row_begin, int64_t row_end, SearchPlan *plan) {
  if (!simt) return tc_plan(n, d, k1, row_begin, row_end, plan);
  plan->simt = true;
  plan->n = n;
  plan->d = d;
  plan->k1 = k1;
  plan->row_begin = row_begin;
  plan->row_end = row_end;
  const int64_t row_tiles = ceil_div(row_end - row_begin, kRows), col_tiles = ceil_div(n, kCols);
  int64_t nseg = ceil_div(4 * (int64_t)sm_count(), row_tiles);
  if (nseg > kMaxLists) nseg = kMaxLists;
  if (nseg > col_tiles) nseg = col_tiles;
  if (nseg < 1) nseg = 1;
  plan->nseg = (int)nseg;
  plan->nlists = (int)nseg;
  plan->margin_c = (double)(d + 16) * ldexp(1.0, -22);
  plan->n_pad_cols = n;
  plan->kp_used = (int)d;
  return 0;
}

int search_prepare(const SearchPlan &plan, const double *X, const double *mu, const double *norm, cudaStream_t stream,
                   SearchState *st) {
  if (!plan.simt) return tc_prepare(plan, X, mu, norm, stream, st);
  MELD_CHECK(st->xc32.alloc((size_t)plan.n * plan.d));
  MELD_CHECK(st->hn32.alloc((size_t)plan.n));
  simt_prep_kernel<<<sm_count() * 8, 256, 0, stream>>>(X, mu, norm, plan.n, plan.d, st->xc32.p, st->hn32.p);
  MELD_LAUNCH_CHECK();
  return 0;
}

int search_pass1(const SearchPlan &plan, SearchState &st, float *lists, cudaStream_t stream) {
  if (!plan.simt) return tc_pass(plan, st, 1, lists, nullptr, nullptr, nullptr, 0, stream);
  const size_t smem = simt_smem(plan.k1);
  MELD_CUDA(cudaFuncSetAttribute(simt_search_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid((unsigned)ceil_div(plan.row_end - plan.row_begin, kRows), (unsigned)plan.nseg);
  simt_search_kernel<1><<<grid, kRows, smem, stream>>>(st.xc32.p, st.hn32.p, plan.n, plan.d, plan.row_begin, plan.row_end,
                                                       plan.k1, plan.nseg, lists,
                                                       nullptr, nullptr, nullptr, 0);
  MELD_LAUNCH_CHECK();
  return 0;
}

int search_pass2(const SearchPlan &plan, SearchState &st, const float *key2, unsigned long long *pairs,
                 unsigned long long *count, int64_t cap, cudaStream_t stream) {
  if (!plan.simt) return tc_pass(plan, st, 2, nullptr, key2, pairs, count, cap, stream);
  const size_t smem = simt_smem(0);
  MELD_CUDA(cudaFuncSetAttribute(simt_search_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid((unsigned)ceil_div(plan.row_end - plan.row_begin, kRows), (unsigned)plan.nseg);
  simt_search_kernel<2><<<grid, kRows, smem, stream>>>(st.xc32.p, st.hn32.p, plan.n, plan.d, plan.row_begin, plan.row_end,
                                                       plan.k1, plan.nseg,
                                                       nullptr, key2, pairs, count, cap);
  MELD_LAUNCH_CHECK();
  return 0;
}

void search_release(SearchState *st) {
  st->xc32.release();
  st->hn32.release();
  st->thr_g.release();
  st->a_op.release();
  st->b_op.release();
}

}  // namespace meld
