// Microbenchmark (development probe, run under gpurun): how many random 32-byte rows per cycle can one SM pull
// out of an L2-resident table through (a) LDG.256 into registers, (b) cp.async (LDGSTS) into shared memory,
// (c) cp.async.bulk (TMA, UBLKCP) 32-byte copies into shared memory?  The Chebyshev SpMM gathers one such row
// per nonzero; its L1 path saturates at ~1 row / 2 cycles / SM.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gather_bench gather_bench.cu
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint32_t s32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *b, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s32(b)), "r"(c) : "memory"); }
__device__ __forceinline__ void mbar_expect(uint64_t *b, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(b)), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t *b, uint32_t parity) {
  uint32_t ok;
  do {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(s32(b)), "r"(parity) : "memory");
  } while (!ok);
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(s32(dst)), "l"(src), "r"(bytes), "r"(s32(bar)) : "memory");
}

// (a) registers
__global__ void k_ldg(const double4 *__restrict__ T, const int *__restrict__ idx, int per_cta, double *out) {
  const int *my = idx + (size_t)blockIdx.x * per_cta;
  double acc = 0;
  for (int i = threadIdx.x; i < per_cta; i += blockDim.x * 4) {
    double4 v[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int j = i + u * blockDim.x;
      v[u] = make_double4(0, 0, 0, 0);
      if (j < per_cta)
        asm volatile("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(v[u].x), "=d"(v[u].y), "=d"(v[u].z), "=d"(v[u].w) : "l"(T + my[j]));
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) acc += v[u].x + v[u].w;
  }
  if (acc == 12345.678) out[0] = acc;
}

// (c) TMA bulk copies: `issuers` threads each issue copies; stages of `batch` rows
template <int ROW_BYTES>
__global__ void k_bulk(const char *__restrict__ T, const int *__restrict__ idx, int per_cta, int batch, int stages, double *out) {
  extern __shared__ __align__(128) unsigned char sm[];
  uint64_t *bars = reinterpret_cast<uint64_t *>(sm);
  unsigned char *buf = sm + 128;
  const int *my = idx + (size_t)blockIdx.x * per_cta;
  if (threadIdx.x == 0) {
    for (int s = 0; s < stages; ++s) mbar_init(&bars[s], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const int nb = per_cta / batch;
  double acc = 0;
  // prologue: fill the pipeline
  for (int b = 0; b < stages && b < nb; ++b) {
    if (threadIdx.x == 0) mbar_expect(&bars[b], (uint32_t)batch * ROW_BYTES);
    __syncthreads();
    for (int r = threadIdx.x; r < batch; r += blockDim.x)
      bulk_g2s(buf + ((size_t)b * batch + r) * ROW_BYTES, T + (size_t)my[b * batch + r] * ROW_BYTES, ROW_BYTES, &bars[b]);
  }
  for (int b = 0; b < nb; ++b) {
    const int s = b % stages;
    mbar_wait(&bars[s], (uint32_t)(b / stages) & 1u);
    // consume one double per row (keeps the data path honest, costs one LDS per row)
    for (int r = threadIdx.x; r < batch; r += blockDim.x) acc += *reinterpret_cast<double *>(buf + ((size_t)s * batch + r) * ROW_BYTES);
    __syncthreads();
    const int nbx = b + stages;
    if (nbx < nb) {
      if (threadIdx.x == 0) mbar_expect(&bars[s], (uint32_t)batch * ROW_BYTES);
      __syncthreads();
      for (int r = threadIdx.x; r < batch; r += blockDim.x)
        bulk_g2s(buf + ((size_t)s * batch + r) * ROW_BYTES, T + (size_t)my[nbx * batch + r] * ROW_BYTES, ROW_BYTES, &bars[s]);
    }
  }
  if (acc == 12345.678) out[0] = acc;
}

// (b) LDGSTS 16-byte copies (2 per row) + wait_group pipeline
__global__ void k_ldgsts(const char *__restrict__ T, const int *__restrict__ idx, int per_cta, int batch, double *out) {
  extern __shared__ __align__(128) unsigned char sm[];
  unsigned char *buf = sm;
  const int *my = idx + (size_t)blockIdx.x * per_cta;
  const int nb = per_cta / batch;
  double acc = 0;
  auto issue = [&](int b, int s) {
    for (int r = threadIdx.x; r < batch * 2; r += blockDim.x) {
      const int row = r >> 1, half = r & 1;
      const char *src = T + (size_t)my[b * batch + row] * 32 + half * 16;
      asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(s32(buf + ((size_t)s * batch + row) * 32 + half * 16)), "l"(src) : "memory");
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  issue(0, 0);
  for (int b = 0; b < nb; ++b) {
    if (b + 1 < nb) issue(b + 1, (b + 1) & 1);
    if (b + 1 < nb) asm volatile("cp.async.wait_group 1;" ::: "memory"); else asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();
    for (int r = threadIdx.x; r < batch; r += blockDim.x) acc += *reinterpret_cast<double *>(buf + ((size_t)(b & 1) * batch + r) * 32);
    __syncthreads();
  }
  if (acc == 12345.678) out[0] = acc;
}

int main() {
  const int N = 500000, SM = 148;
  const int per_cta = 178176;  // multiple of 1024; ~26.4M gathers in total
  std::vector<int> h((size_t)SM * per_cta);
  uint64_t z = 88172645463325252ull;
  for (auto &v : h) { z ^= z << 13; z ^= z >> 7; z ^= z << 17; v = (int)(z % N); }
  int *idx; char *T; double *out;
  CK(cudaMalloc(&idx, h.size() * 4));
  CK(cudaMalloc(&T, (size_t)N * 64));
  CK(cudaMalloc(&out, 8));
  CK(cudaMemcpy(idx, h.data(), h.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemset(T, 0, (size_t)N * 64));
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  int clk_khz = 0; cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
  auto report = [&](const char *name, float ms) {
    const double g = (double)SM * per_cta;
    printf("%-44s %8.1f us  %6.2f Ggather/s  %5.3f rows/cycle/SM (at %.0f MHz nominal)\n", name, ms * 1e3, g / ms / 1e6,
           g / SM / (ms * 1e-3 * clk_khz * 1e3), clk_khz / 1e3);
    fflush(stdout);
  };
  float ms;
  for (int threads : {256, 512, 1024}) {
    for (int rep = 0; rep < 2; ++rep) {
      cudaEventRecord(e0);
      k_ldg<<<SM, threads>>>((const double4 *)T, idx, per_cta, out);
      cudaEventRecord(e1); CK(cudaEventSynchronize(e1)); cudaEventElapsedTime(&ms, e0, e1);
    }
    char nm[96]; snprintf(nm, 96, "LDG.256 -> registers, %d threads", threads); report(nm, ms);
  }
  for (int threads : {256, 512}) {
    const int batch = 1024;
    CK(cudaFuncSetAttribute(k_ldgsts, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * batch * 32));
    for (int rep = 0; rep < 2; ++rep) {
      cudaEventRecord(e0);
      k_ldgsts<<<SM, threads, 2 * batch * 32>>>(T, idx, per_cta, batch, out);
      cudaEventRecord(e1); CK(cudaEventSynchronize(e1)); cudaEventElapsedTime(&ms, e0, e1);
    }
    char nm[96]; snprintf(nm, 96, "LDGSTS 2x16B -> smem, %d threads", threads); report(nm, ms);
  }
  for (int threads : {32, 128, 512}) {
    for (int batch : {256, 1024}) {
      for (int stages : {2, 4}) {
        const size_t smem = 128 + (size_t)stages * batch * 32;
        CK(cudaFuncSetAttribute(k_bulk<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        for (int rep = 0; rep < 2; ++rep) {
          cudaEventRecord(e0);
          k_bulk<32><<<SM, threads, smem>>>(T, idx, per_cta, batch, stages, out);
          cudaEventRecord(e1); CK(cudaEventSynchronize(e1)); cudaEventElapsedTime(&ms, e0, e1);
        }
        char nm[96]; snprintf(nm, 96, "TMA bulk 32B -> smem, %d thr, batch %d, %d st", threads, batch, stages); report(nm, ms);
      }
    }
  }
  {  // 64-byte rows (p = 8)
    const int threads = 128, batch = 1024, stages = 2;
    const size_t smem = 128 + (size_t)stages * batch * 64;
    CK(cudaFuncSetAttribute(k_bulk<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    for (int rep = 0; rep < 2; ++rep) {
      cudaEventRecord(e0);
      k_bulk<64><<<SM, threads, smem>>>(T, idx, per_cta, batch, stages, out);
      cudaEventRecord(e1); CK(cudaEventSynchronize(e1)); cudaEventElapsedTime(&ms, e0, e1);
    }
    report("TMA bulk 64B -> smem, 128 thr, batch 1024, 2 st", ms);
  }
  CK(cudaDeviceSynchronize());
  return 0;
}
