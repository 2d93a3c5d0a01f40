// tcgen05 candidate search: the pairwise-distance contraction on 5th-gen tensor cores, restricted to the
// (row tile, column tile) pairs a one-level ball tree cannot rule out.
//
// s_ij = x_i . x_j - n_j/2 for all pairs is one bf16 GEMM  S = A' B'^T  with fp32 accumulation:
//   A'_i = [ hi(x_i) | hi(x_i) | lo(x_i) | 1 1 1 | 0.. ]      (x = hi + lo, both bf16)
//   B'_j = [ hi(x_j) | lo(x_j) | hi(x_j) | h0 h1 h2 | 0.. ]   (h0+h1+h2 = -n_j/2 in bf16 pieces)
// so the epilogue needs ONE compare per element: pass 1 keeps the k1 largest s per row, pass 2
// appends every column with s >= key2_i.  K' = 3d+3 is padded to a multiple of 64.
//
// Kernel: persistent, warp-specialised, one CTA per SM (320 threads), 2-CTA clusters sharing B tiles by multicast:
//   warp 0     TMA producer  (cp.async.bulk.tensor 2-D, 128B swizzle, mbarrier complete_tx)
//   warp 1     MMA issuer    (tcgen05.mma cta_group::1 kind::f16, M=128 N=256 K=16; TMEM alloc)
//   warps 2-9  epilogue      (tcgen05.ld 32x32b: thread = row, 32 columns per load; two warps per TMEM lane
//                             quadrant, each scanning half of a tile's columns)
// A work unit is (128-row tile, chunk of the column-tile list of its 256-row group) -- or, with pruning off,
// (row tile, column segment).  When K' <= 384 the row tile's A operand stays resident in shared memory for the
// whole unit and only B tiles stream through a 4-stage ring; otherwise A and B stream together.  The 512 TMEM
// columns hold two 128x256 fp32 accumulators so the epilogue of tile t overlaps the MMAs of tile t+1.
// Which tiles are listed: tile_balls_kernel / tile_proj_kernel / tile_lists_kernel below (bounding balls and
// centroid-axis projections of 256-cell tiles of the k-means + Morton cell order, triangle inequality).
#include "knn_search.cuh"

#include <cuda.h>
#include <cuda_bf16.h>

namespace meld {

constexpr int BM = 128, BN = 256, BK = 64;
constexpr int kMaxStages = 8;
constexpr int kEpiWarps = 8;            // two per TMEM lane quadrant, each scanning half of a tile's columns
constexpr int kTcThreads = (2 + kEpiWarps) * 32;
constexpr int kABytes = BM * BK * 2;   // 16 KB
constexpr int kBBytes = BN * BK * 2;   // 32 KB
constexpr int kMaxResidentKb = 6;      // A resident up to K' = 384
constexpr float kPadSentinel = -1e30f;
constexpr int kQueueCap = 256;          // pass 2: per-warp queue of hit pairs (flushed when the next chunk would not fit)

// ---- operand preparation ------------------------------------------------------------------------
__global__ void tc_prep_kernel(const double *__restrict__ X, const double *__restrict__ mu,
                               const double *__restrict__ norm, int64_t n, int64_t n_pad, int64_t d, int kp,
                               __nv_bfloat16 *__restrict__ A, __nv_bfloat16 *__restrict__ B) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const __nv_bfloat16 zero = __float2bfloat16(0.f), one = __float2bfloat16(1.f);
  for (int64_t i = warp; i < n_pad; i += nwarps) {
    __nv_bfloat16 *a = A + i * kp, *b = B + i * kp;
    if (i >= n) {
      for (int k = lane; k < kp; k += 32) {
        a[k] = zero;
        b[k] = (k == 3 * d) ? __float2bfloat16(kPadSentinel) : zero;
      }
      continue;
    }
    for (int64_t k = lane; k < d; k += 32) {
      const float xf = (float)(X[i * d + k] - mu[k]);
      const __nv_bfloat16 hi = __float2bfloat16(xf);
      const __nv_bfloat16 lo = __float2bfloat16(xf - __bfloat162float(hi));
      a[k] = hi;
      a[d + k] = hi;
      a[2 * d + k] = lo;
      b[k] = hi;
      b[d + k] = lo;
      b[2 * d + k] = hi;
    }
    const float h = (float)(-0.5 * norm[i]);
    const __nv_bfloat16 h0 = __float2bfloat16(h);
    const float r1 = h - __bfloat162float(h0);
    const __nv_bfloat16 h1 = __float2bfloat16(r1);
    const __nv_bfloat16 h2 = __float2bfloat16(r1 - __bfloat162float(h1));
    for (int k = 3 * (int)d + lane; k < kp; k += 32) {
      const int t = k - 3 * (int)d;
      a[k] = t < 3 ? one : zero;
      b[k] = t == 0 ? h0 : (t == 1 ? h1 : (t == 2 ? h2 : zero));
    }
  }
}

__global__ void fill_float_kernel(float *p, int64_t n, float v) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) p[i] = v;
}

// ---- PTX wrappers -----------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t s32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void bar_init(uint64_t *bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void bar_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bar_arrive(uint64_t *bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(s32(bar)) : "memory");
}
__device__ __forceinline__ void bar_wait(uint64_t *bar, uint32_t parity) {
  const uint32_t addr = s32(bar);
  uint32_t ok;
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
__device__ __forceinline__ void tma_load_2d(const void *tmap, uint64_t *bar, void *dst, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
          s32(dst)),
      "l"(tmap), "r"(s32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d_mc(const void *tmap, uint64_t *bar, void *dst, int c0, int c1,
                                               uint16_t cta_mask) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster "
      "[%0], [%1, {%3, %4}], [%2], %5;" ::"r"(s32(dst)),
      "l"(tmap), "r"(s32(bar)), "r"(c0), "r"(c1), "h"(cta_mask)
      : "memory");
}
__device__ __forceinline__ void tc_commit_mc(uint64_t *bar, uint16_t cta_mask) {
  asm volatile(
      "tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
          s32(bar)),
      "h"(cta_mask)
      : "memory");
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t *bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(s32(bar))
               : "memory");
}
__device__ __forceinline__ void tc_mma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                            uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// K-major operand tile, 128-byte swizzle: rows of 128 B, 8-row groups 1024 B apart.
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);  // start address
  d |= (uint64_t)1 << 16;                     // leading byte offset (unused for swizzled K-major)
  d |= (uint64_t)(1024 >> 4) << 32;           // stride byte offset: 8 rows x 128 B
  d |= (uint64_t)1 << 46;                     // descriptor version (sm_100)
  d |= (uint64_t)2 << 61;                     // SWIZZLE_128B
  return d;
}
// kind::f16 instruction descriptor: D = f32, A = B = bf16, both K-major, M x N
__host__ __device__ constexpr uint32_t umma_idesc_bf16(int m, int n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
        "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
        "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

struct TcArgs {
  int mode;  // 1: top-k1 lists, 2: emit candidates
  int64_t n;
  int n_row_tiles, n_col_tiles, nseg, nkb, a_resident, k1;  // n_row_tiles: row tiles of THIS call
  int rt_begin;        // first (global) row tile of this call
  long long row_end;   // rows at or beyond it are not this call's
  int mc;         // cluster size (1, 2 or 4): each CTA loads 1/mc of a B tile and multicasts it to the cluster
  int n_passes;   // segments visited per row tile (pass 1 may stop early: any subset gives a valid bound)
  int n_stages;   // B (or A+B) ring depth
  int a_region;   // bytes reserved for the A operand (resident k-blocks, or one A tile per stage)
  float *thr_g;   // pass 1: per-row running lower bound of the k1-th largest s (carried across segments)
  float *lists;
  const float *key2;
  unsigned long long *pairs;   // pass 2: global append buffer of (row << 32 | col)
  unsigned long long *count;   // its fill count (keeps growing past cap)
  long long cap;
  // Pruned search (ball-tree style): when tl_list is set, the column tiles a 256-row group must visit come
  // from a per-group list (tl_list[(g - tl_g0) * tl_stride ..], tl_len[g - tl_g0] entries, ascending) instead
  // of "all of them"; a unit is (row tile, chunk of its group's list), n_passes chunks per row tile.
  const int32_t *tl_list;
  const int32_t *tl_len;
  int tl_stride, tl_g0;
  int slot0;      // pass 1: first list slot of this launch (the window pass writes slot 0, the list pass 1..)
  int n_slots;    // pass 1: list slots per (row, half) pair in `lists`
  int interleave; // 1: chunk c = list entries c, c + n_passes, ... ; 0: contiguous chunks
};

// Work unit u -> (row tile, segment, column-tile order).  Units are issued segment-major so concurrent CTAs
// stream the same L2-resident slice of B, but every row tile visits ITS OWN segment first and, inside a
// segment, starts at the column tile that holds its own rows and wraps around: cells are Morton-ordered, so
// a row's true neighbours sit near its own position -- scanning outwards from there makes the running
// top-k threshold tight after the first tile instead of after a slow monotone approach.
struct UnitPlan {
  int rt, seg, pass, ct0, len, start;
  const int32_t *lp;  // listed column tiles of this unit (pruned search), or nullptr
  int lstep;          // stride through the list (chunks interleave)
  __device__ __forceinline__ int tile(int j) const { return lp ? __ldg(lp + (size_t)j * lstep) : ct0 + (start + j) % len; }
};
__device__ __forceinline__ UnitPlan unit_plan(const TcArgs &a, int u) {
  UnitPlan p;
  const int pass = u / a.n_row_tiles;  // 0: own segment, then the following ones cyclically
  p.rt = a.rt_begin + u % a.n_row_tiles;
  p.pass = pass;
  p.lp = nullptr;
  p.lstep = 1;
  if (a.tl_list != nullptr) {
    // chunk `pass` of the list of the 256-row group this row tile belongs to; the CTAs of a 2-cluster hold the
    // two row tiles of one group, so they walk identical sequences in lockstep
    // Chunks interleave (chunk c = list entries c, c + n_passes, ...): every chunk is a uniform sample of the
    // group's tiles, so in pass 1 the bound published by chunk c (units are issued chunk-major) is already
    // within a few percent of the final one and later chunks rarely take the insertion path.
    const int gl = (p.rt * BM) / BN - a.tl_g0;
    const int n = __ldg(a.tl_len + gl);
    if (a.interleave) {
      p.lp = a.tl_list + (size_t)gl * a.tl_stride + pass;
      p.lstep = a.n_passes;
      p.len = n > pass ? (n - pass + a.n_passes - 1) / a.n_passes : 0;
    } else {
      const int b = (int)(((long long)n * pass) / a.n_passes), e = (int)(((long long)n * (pass + 1)) / a.n_passes);
      p.lp = a.tl_list + (size_t)gl * a.tl_stride + b;
      p.lstep = 1;
      p.len = e - b;
    }
    p.seg = pass;
    p.ct0 = 0;
    p.start = 0;
    return p;
  }
  const int own_ct = ((p.rt - p.rt % a.mc) * BM) / BN;  // identical for every CTA of a cluster (lockstep)
  int own_seg = (int)(((int64_t)own_ct * a.nseg) / a.n_col_tiles);
  while (own_seg + 1 < a.nseg && (int)((int64_t)a.n_col_tiles * (own_seg + 1) / a.nseg) <= own_ct) ++own_seg;
  while (own_seg > 0 && (int)((int64_t)a.n_col_tiles * own_seg / a.nseg) > own_ct) --own_seg;
  p.seg = (own_seg + pass) % a.nseg;
  p.ct0 = (int)((int64_t)a.n_col_tiles * p.seg / a.nseg);
  const int ct1 = (int)((int64_t)a.n_col_tiles * (p.seg + 1) / a.nseg);
  p.len = ct1 - p.ct0;
  int st = own_ct - p.ct0;
  if (st < 0) st = 0;
  if (st >= p.len) st = p.len - 1;
  p.start = st;
  return p;
}

struct Bars {
  uint64_t full[kMaxStages];
  uint64_t empty[kMaxStages];
  uint64_t a_full, a_empty;
  uint64_t tmem_full[2], tmem_empty[2];
  uint32_t tmem_base;
};

// KR > 0 (pass 1, k1 <= KR): every epilogue thread keeps its row's k1 largest s in KR registers, sorted, and
// inserts with a branch-free max/min network (2 KR instructions, no memory).  The KR - k1 leading slots are
// preset to +inf, so the k1-th largest value always sits in the LAST register (a runtime-indexed slot would
// push the whole list to local memory).  The shared-memory lists of the
// KR = 0 variant cost a dependent load per shifted slot -- ~30x slower on the tiles that hold a row's true
// neighbours, which is most of what the pruned search still visits.
template <int KR>
__device__ __forceinline__ void topk_insert(float (&l)[KR > 0 ? KR : 1], float s) {
#pragma unroll
  for (int i = 0; i < KR; ++i) {
    const float hi = fmaxf(l[i], s);
    s = fminf(l[i], s);
    l[i] = hi;
  }
}
template <int KR>
__global__ void __launch_bounds__(kTcThreads, 1)
    tc_search_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                     const TcArgs a) {
  extern __shared__ unsigned char smem_dyn[];
  unsigned char *smem = reinterpret_cast<unsigned char *>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~(uintptr_t)1023);
  const int kStages = a.n_stages;
  unsigned char *smA = smem;                   // resident A (nkb x 16 KB) or A stages
  unsigned char *smB = smem + a.a_region;      // B stages
  float *lst = reinterpret_cast<float *>(smB + (size_t)kStages * kBBytes);  // [k1][128] (pass 1)
  Bars *bars = reinterpret_cast<Bars *>(reinterpret_cast<unsigned char *>(lst) + (a.mode == 1 ? (size_t)a.k1 * 2 * BM * 4 : (size_t)kEpiWarps * kQueueCap * 8 + 64));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // Work units are (segment, row tile), segment-major.  With multicast the two CTAs of a cluster take
  // adjacent row tiles of the same segment and walk identical column-tile sequences in lockstep.
  const int n_units = a.n_row_tiles * a.n_passes;  // n_passes segments per row tile, its own first
  const uint32_t crank = a.mc > 1 ? cluster_ctarank() : 0u;
  const int u_first = a.mc * (int)(blockIdx.x / a.mc) + (int)crank;
  const int u_step = (int)gridDim.x;  // even when clustered

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_a) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_b) : "memory");
  }
  if (warp == 1) {
    if (lane == 0) {
      for (int s = 0; s < kStages; ++s) {
        bar_init(&bars->full[s], 1);
        bar_init(&bars->empty[s], a.mc);  // clustered: every CTA's MMAs must release a stage
      }
      bar_init(&bars->a_full, 1);
      bar_init(&bars->a_empty, 1);
      for (int b = 0; b < 2; ++b) {
        bar_init(&bars->tmem_full[b], 1);
        bar_init(&bars->tmem_empty[b], kEpiWarps);  // one arrival per epilogue warp
      }
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s32(&bars->tmem_base)),
                 "r"(512u)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  if (a.mc > 1) cluster_sync_all();  // peer barriers are initialised before any multicast can target them
  tc_fence_after();
  const uint32_t tmem_base = bars->tmem_base;

  if (warp == 0) {
    // ===== TMA producer =====
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0, uphase = 0;
      for (int u = u_first; u < n_units; u += u_step) {
        const UnitPlan up = unit_plan(a, u);
        const int rt = up.rt;
        if (up.len <= 0) continue;  // empty chunk of a pruned list: no role touches a barrier for it
        if (a.a_resident) {
          bar_wait(&bars->a_empty, uphase ^ 1u);
          bar_expect_tx(&bars->a_full, (uint32_t)a.nkb * kABytes);
          for (int kb = 0; kb < a.nkb; ++kb) tma_load_2d(&tmap_a, &bars->a_full, smA + kb * kABytes, kb * BK, rt * BM);
        }
        for (int j = 0; j < up.len; ++j) {
          const int ct = up.tile(j);
          for (int kb = 0; kb < a.nkb; ++kb) {
            bar_wait(&bars->empty[stage], phase ^ 1u);
            bar_expect_tx(&bars->full[stage], a.a_resident ? kBBytes : kBBytes + kABytes);
            if (a.mc > 1) {  // my slice of the B tile, delivered to every CTA of the cluster
              const int slice = BN / a.mc;
              tma_load_2d_mc(&tmap_b, &bars->full[stage], smB + stage * kBBytes + crank * (kBBytes / a.mc), kb * BK,
                             ct * BN + (int)crank * slice, (uint16_t)((1u << a.mc) - 1u));
            } else {
              tma_load_2d(&tmap_b, &bars->full[stage], smB + stage * kBBytes, kb * BK, ct * BN);
            }
            if (!a.a_resident) tma_load_2d(&tmap_a, &bars->full[stage], smA + stage * kABytes, kb * BK, rt * BM);
            if (++stage == kStages) {
              stage = 0;
              phase ^= 1u;
            }
          }
        }
        uphase ^= 1u;
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer (one thread) =====
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(BM, BN);
      int stage = 0, buf = 0;
      uint32_t phase = 0, bphase = 0, uphase = 0;
      for (int u = u_first; u < n_units; u += u_step) {
        const UnitPlan up = unit_plan(a, u);
        if (up.len <= 0) continue;
        if (a.a_resident) {
          bar_wait(&bars->a_full, uphase);
          tc_fence_after();
        }
        for (int j = 0; j < up.len; ++j) {
          bar_wait(&bars->tmem_empty[buf], bphase ^ 1u);
          tc_fence_after();
          const uint32_t tmem_d = tmem_base + (uint32_t)buf * BN;
          for (int kb = 0; kb < a.nkb; ++kb) {
            bar_wait(&bars->full[stage], phase);
            tc_fence_after();
            const uint32_t a_base = s32(a.a_resident ? smA + kb * kABytes : smA + stage * kABytes);
            const uint32_t b_base = s32(smB + stage * kBBytes);
#pragma unroll
            for (int k = 0; k < BK / 16; ++k) {
              tc_mma_bf16(tmem_d, umma_desc_sw128(a_base + k * 32), umma_desc_sw128(b_base + k * 32), idesc,
                          (uint32_t)((kb | k) != 0));
            }
            if (a.mc > 1)
              tc_commit_mc(&bars->empty[stage], (uint16_t)((1u << a.mc) - 1u));  // releases the stage cluster-wide
            else
              tc_commit(&bars->empty[stage]);  // stage reusable once these MMAs retire
            if (++stage == kStages) {
              stage = 0;
              phase ^= 1u;
            }
          }
          tc_commit(&bars->tmem_full[buf]);
          buf ^= 1;
          if (buf == 0) bphase ^= 1u;
        }
        if (a.a_resident) tc_commit(&bars->a_empty);
        uphase ^= 1u;
      }
    }
  } else {
    // ===== epilogue: thread = row of the tile =====
    const int q = warp & 3;            // TMEM lane quadrant this warp may read
    const int half = (warp - 2) >> 2;  // which half of a tile's 256 columns this warp scans
    const int rin = q * 32 + lane;     // row of the tile
    const int lcol = half * BM + rin;  // this thread's column in the shared top-k lists
    int buf = 0;
    uint32_t bphase = 0;
    for (int u = u_first; u < n_units; u += u_step) {
      const UnitPlan up = unit_plan(a, u);
      const int rt = up.rt;
      const int64_t row = (int64_t)rt * BM + rin;
      if (up.len <= 0) {  // empty chunk: its pass-1 list is all -inf
        if (a.mode == 1 && row < a.row_end) {
          float *out = a.lists + ((size_t)row * a.n_slots * 2 + (size_t)(a.slot0 + up.pass) * 2 + half) * a.k1;
          for (int s = 0; s < a.k1; ++s) out[s] = -INFINITY;
        }
        continue;
      }
      float thr;
      float rl[KR > 0 ? KR : 1];
#pragma unroll
      for (int i = 0; i < (KR > 0 ? KR : 1); ++i) rl[i] = (i < KR - a.k1) ? INFINITY : -INFINITY;
      if (a.mode == 1) {
        if (KR == 0)
          for (int s = 0; s < a.k1; ++s) lst[s * 2 * BM + lcol] = -INFINITY;
        // start from the bound earlier segments of this row have published: values below it cannot be
        // among the k1 largest of the union, so this list only needs what beats it
        thr = row < a.row_end ? __ldcg(a.thr_g + row) : INFINITY;
      } else {
        thr = row < a.row_end ? a.key2[row] : INFINITY;
      }
      // pass 2: hits are queued per warp in shared memory and appended to the global pair buffer in bulk (one
      // global atomic per flush).  The queue length lives in a warp-uniform register: a chunk's hits are placed by
      // a warp prefix sum over the per-lane hit counts -- ONE reservation per warp per chunk, no shared-memory atomic
      // (round 1 did atomicAdd(wqn, 1) per hit: 32 lanes on one bank, 71 % of the kernel's shared wavefronts were
      // bank conflicts, profiles/r01d_ncu_tc_search_pruned_c4.txt).
      unsigned long long *wq = reinterpret_cast<unsigned long long *>(lst) + (size_t)(warp - 2) * kQueueCap;
      int qn = 0;  // warp-uniform
      auto flush_queue = [&]() {  // warp-uniform
        __syncwarp();
        if (qn > 0) {
          unsigned long long base = 0;
          if (lane == 0) base = atomicAdd(a.count, (unsigned long long)qn);
          base = __shfl_sync(0xffffffffu, base, 0);
          for (int i = lane; i < qn; i += 32)
            if ((long long)(base + i) < a.cap) a.pairs[base + i] = wq[i];
        }
        __syncwarp();
        qn = 0;
      };
      for (int j = 0; j < up.len; ++j) {
        const int ct = up.tile(j);
        bar_wait(&bars->tmem_full[buf], bphase);
        tc_fence_after();
        const uint32_t tbase = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)buf * BN;
        // Branch-free fast path: one max over the 32 columns of a chunk and a single compare; only a
        // chunk that holds a hit is scanned element by element.  Two chunks are kept in flight so the
        // TMEM load of the next one overlaps the scan of the current one.
        auto scan = [&](const uint32_t (&v)[32], int chunk, float (&rl)[KR > 0 ? KR : 1]) {
          float t16[16];  // log-depth max (a serial chain of 31 dependent FMNMX would cost ~130 cycles per chunk)
#pragma unroll
          for (int c = 0; c < 16; ++c) t16[c] = fmaxf(__uint_as_float(v[c]), __uint_as_float(v[c + 16]));
#pragma unroll
          for (int w = 8; w > 0; w >>= 1) {
#pragma unroll
            for (int c = 0; c < w; ++c) t16[c] = fmaxf(t16[c], t16[c + w]);
          }
          const float mx = t16[0];
          const int col0 = ct * BN + chunk * 32;
          if (a.mode == 1) {
            if (mx > thr) {
              if (KR > 0) {
                // hit mask first (branch-free), then ONE copy of the insertion code in a loop over the set bits:
                // the element is picked out of the 32 registers by a 5-level select tree (static indices only).
                // Fully unrolling 32 insertion networks costs ~1100 instructions per call site and stalls the
                // epilogue warps on instruction fetch.
                unsigned m = 0;
#pragma unroll
                for (int c = 0; c < 32; ++c) m |= (__uint_as_float(v[c]) > thr) ? (1u << c) : 0u;
                while (m) {
                  const int c = __ffs(m) - 1;
                  m &= m - 1;
                  uint32_t t16[16], t8[8], t4[4], t2[2];
#pragma unroll
                  for (int i = 0; i < 16; ++i) t16[i] = (c & 1) ? v[2 * i + 1] : v[2 * i];
#pragma unroll
                  for (int i = 0; i < 8; ++i) t8[i] = (c & 2) ? t16[2 * i + 1] : t16[2 * i];
#pragma unroll
                  for (int i = 0; i < 4; ++i) t4[i] = (c & 4) ? t8[2 * i + 1] : t8[2 * i];
#pragma unroll
                  for (int i = 0; i < 2; ++i) t2[i] = (c & 8) ? t4[2 * i + 1] : t4[2 * i];
                  const float s = __uint_as_float((c & 16) ? t2[1] : t2[0]);
                  if (s > thr) {  // thr may have risen since the mask was taken
                    topk_insert<KR>(rl, s);
                    thr = fmaxf(thr, rl[KR > 0 ? KR - 1 : 0]);
                  }
                }
              } else {
#pragma unroll  // static register indices (a dynamic index would spill v to local memory)
                for (int c = 0; c < 32; ++c) {
                  const float s = __uint_as_float(v[c]);
                  if (s > thr) {
                    int pos = a.k1 - 1;
                    while (pos > 0 && lst[(pos - 1) * 2 * BM + lcol] < s) {
                      lst[pos * 2 * BM + lcol] = lst[(pos - 1) * 2 * BM + lcol];
                      --pos;
                    }
                    lst[pos * 2 * BM + lcol] = s;
                    thr = fmaxf(thr, lst[(a.k1 - 1) * 2 * BM + lcol]);
                  }
                }
              }
            }
          } else {
            // warp-uniform vote first: most chunks of a surviving tile hold no hit for any of the 32 rows
            if (__ballot_sync(0xffffffffu, mx >= thr) != 0u) {
              unsigned m = 0;
              if (mx >= thr) {
#pragma unroll  // branch-free hit mask (static register indices)
                for (int c = 0; c < 32; ++c) m |= (__uint_as_float(v[c]) >= thr) ? (1u << c) : 0u;
              }
              const int cnt = __popc(m);
              int pre = cnt;  // inclusive prefix sum of the lanes' hit counts
#pragma unroll
              for (int o = 1; o < 32; o <<= 1) {
                const int t = __shfl_up_sync(0xffffffffu, pre, o);
                if (lane >= o) pre += t;
              }
              const int total = __shfl_sync(0xffffffffu, pre, 31);
              pre -= cnt;
              if (total > kQueueCap) {  // dense chunk: straight to the pair buffer, still one reservation per warp
                unsigned long long base = 0;
                if (lane == 0) base = atomicAdd(a.count, (unsigned long long)total);
                base = __shfl_sync(0xffffffffu, base, 0) + (unsigned long long)pre;
                while (m) {
                  const int c = __ffs(m) - 1;
                  m &= m - 1;
                  if ((long long)base < a.cap) a.pairs[base] = ((unsigned long long)row << 32) | (unsigned long long)(col0 + c);
                  ++base;
                }
              } else {
                if (qn + total > kQueueCap) flush_queue();
                int pos = qn + pre;
                while (m) {
                  const int c = __ffs(m) - 1;
                  m &= m - 1;
                  wq[pos++] = ((unsigned long long)row << 32) | (unsigned long long)(col0 + c);
                }
                qn += total;
              }
            }
          }
        };
        uint32_t va[32], vb[32];
        constexpr int kChunks = BN / 32 / 2;  // chunks per half
        const int c0 = half * kChunks;
        tmem_ld32(tbase + c0 * 32, va);
#pragma unroll 1
        for (int chunk = 0; chunk < kChunks; chunk += 2) {
          tmem_ld_wait();
          tmem_ld32(tbase + (c0 + chunk + 1) * 32, vb);
          scan(va, c0 + chunk, rl);
          tmem_ld_wait();
          if (chunk + 2 < kChunks) tmem_ld32(tbase + (c0 + chunk + 2) * 32, va);
          scan(vb, c0 + chunk + 1, rl);
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) bar_arrive(&bars->tmem_empty[buf]);
        buf ^= 1;
        if (buf == 0) bphase ^= 1u;
      }
      if (a.mode == 2) flush_queue();
      if (a.mode == 1 && row < a.row_end) {
        float *out = a.lists + ((size_t)row * a.n_slots * 2 + (size_t)(a.slot0 + up.pass) * 2 + half) * a.k1;  // list per (slot, half)
        float mine;
        if (KR > 0) {
#pragma unroll
          for (int s = 0; s < KR; ++s)
            if (s >= KR - a.k1) out[s - (KR - a.k1)] = rl[s];
          mine = rl[KR > 0 ? KR - 1 : 0];
        } else {
          for (int s = 0; s < a.k1; ++s) out[s] = lst[s * 2 * BM + lcol];
          mine = lst[(a.k1 - 1) * 2 * BM + lcol];
        }
        // publish (monotone max; floats >= 0 and < 0 both ordered through the signed/unsigned trick)
        if (mine > -INFINITY) {
          if (mine >= 0.f)
            atomicMax(reinterpret_cast<int *>(a.thr_g + row), __float_as_int(mine));
          else
            atomicMin(reinterpret_cast<unsigned int *>(a.thr_g + row), __float_as_uint(mine));
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (a.mc > 1) cluster_sync_all();  // the peer may still multicast into / arrive on this CTA until it is done too
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}


// ---- pruned search: bounding balls of 256-cell tiles and the column-tile lists --------------------
// The reference's sklearn ball tree prunes by the triangle inequality; here the tree is one level of
// 256-cell tiles of the internal cell order (k-means clusters, Morton curve inside a cluster).  A tile that
// straddles a cluster boundary keeps one ball per side, so the ~n_clusters straddlers cost two ordinary
// tiles instead of being near everything.  Any centre gives a valid bound as long as the radius is the
// maximum distance to that same centre.
constexpr int kTile = BN;  // cells per tile = one column tile = the two row tiles of a 2-CTA cluster

__global__ void __launch_bounds__(kTile) tile_balls_kernel(const double *__restrict__ X, int64_t n, int64_t d,
                                                           const int32_t *__restrict__ cid, double *__restrict__ ball_c,
                                                           double *__restrict__ ball_rho) {
  __shared__ int s_split;
  __shared__ unsigned long long s_rho[2];
  const int64_t t = blockIdx.x, t0 = t * kTile;
  const int tid = threadIdx.x;
  int64_t rem = n - t0;
  const int cnt = rem <= 0 ? 0 : (rem < kTile ? (int)rem : kTile);
  if (tid == 0) {
    s_split = cnt;
    s_rho[0] = s_rho[1] = 0ull;
  }
  __syncthreads();
  if (cid != nullptr && tid < cnt && cid[t0 + tid] != cid[t0]) atomicMin(&s_split, tid);
  __syncthreads();
  const int split = s_split;
  double *c0 = ball_c + (size_t)t * 2 * d, *c1 = c0 + d;
  for (int64_t k = tid; k < d; k += kTile) {
    double s0 = 0.0, s1 = 0.0;
    for (int r = 0; r < split; ++r) s0 += X[(t0 + r) * d + k];
    for (int r = split; r < cnt; ++r) s1 += X[(t0 + r) * d + k];
    c0[k] = split > 0 ? s0 / split : 0.0;
    c1[k] = cnt > split ? s1 / (cnt - split) : 0.0;
  }
  __syncthreads();  // the block's own global writes are visible to it after the barrier
  if (tid < cnt) {
    const int b = tid < split ? 0 : 1;
    const double *c = b ? c1 : c0;
    const double *x = X + (t0 + tid) * d;
    double acc = 0.0;
    for (int64_t k = 0; k < d; ++k) {
      const double diff = x[k] - c[k];
      acc = fma(diff, diff, acc);
    }
    atomicMax(&s_rho[b], (unsigned long long)__double_as_longlong(acc));  // acc >= 0: bit order = value order
  }
  __syncthreads();
  if (tid < 2) {
    const bool has = tid == 0 ? split > 0 : cnt > split;
    const double r2 = __longlong_as_double((long long)s_rho[tid]);
    ball_rho[t * 2 + tid] = has ? sqrt(r2) * (1.0 + 1e-12) + 1e-300 : -1.0;
  }
}

// |c_A|^2 and |c_A - c_B| of the k-means centroids, in double from the float centroids (cen[k * C + c])
__global__ void cluster_tables_kernel(const float *__restrict__ cen, int nd, int C, double *__restrict__ cl_norm,
                                      double *__restrict__ cl_dist) {
  const int a = blockIdx.x;
  for (int b = threadIdx.x; b < C; b += blockDim.x) {
    double acc = 0.0, na = 0.0;
    for (int k = 0; k < nd; ++k) {
      const double ca = (double)cen[(size_t)k * C + a], cb = (double)cen[(size_t)k * C + b];
      const double diff = ca - cb;
      acc = fma(diff, diff, acc);
      na = fma(ca, ca, na);
    }
    cl_dist[(size_t)a * C + b] = sqrt(acc);
    if (b == 0) cl_norm[a] = na;
  }
}

// Projection bound (see SearchState::tile_hi).  With z = (x - mu) on the selected features and g_B(z) = c_B . z,
//   p_AB(x) = w_AB . (z - m_AB) = (g_B - g_A - (|c_B|^2 - |c_A|^2)/2) / |c_B - c_A|,
// so one pass of "all centroids times the cell" (the k-means assignment's contraction, here in float64) gives a
// cell's projection towards every other cluster.  One block per tile; per segment the maximum over its cells.
template <int CPL>
__global__ void __launch_bounds__(kTile) tile_proj_kernel(const double *__restrict__ X, int64_t n, int64_t d,
                                                          const int32_t *__restrict__ cid, const int32_t *__restrict__ sel,
                                                          int nd, const double *__restrict__ mu,
                                                          const float *__restrict__ cen, int C,
                                                          const double *__restrict__ cl_norm,
                                                          const double *__restrict__ cl_dist,
                                                          int32_t *__restrict__ tile_cl, double *__restrict__ tile_hi) {
  extern __shared__ double pj_sm[];
  double *scen = pj_sm;                          // nd * C
  double *shi = scen + (size_t)nd * C;           // [warps][2][C]
  double *smu = shi + (size_t)(kTile / 32) * 2 * C;  // kKmDims
  int *ssel = reinterpret_cast<int *>(smu + kKmDims);
  __shared__ int s_split;
  const int64_t t = blockIdx.x, t0 = t * kTile;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  int64_t rem = n - t0;
  const int cnt = rem <= 0 ? 0 : (rem < kTile ? (int)rem : kTile);
  for (int i = tid; i < nd * C; i += kTile) scen[i] = (double)cen[i];
  for (int i = tid; i < kKmDims; i += kTile) {
    ssel[i] = i < nd ? sel[i] : 0;
    smu[i] = i < nd ? mu[sel[i]] : 0.0;
  }
  if (tid == 0) s_split = cnt;
  __syncthreads();
  if (tid < cnt && cid[t0 + tid] != cid[t0]) atomicMin(&s_split, tid);
  __syncthreads();
  const int split = s_split;
  // segment clusters: cid is sorted, so a segment is pure iff its first and last cells agree
  const int cl0 = split > 0 ? cid[t0] : -1;
  const int cl1 = (cnt > split && cid[t0 + split] == cid[t0 + cnt - 1]) ? cid[t0 + split] : -1;
  double hm[2][CPL];
#pragma unroll
  for (int sgm = 0; sgm < 2; ++sgm)
#pragma unroll
    for (int c = 0; c < CPL; ++c) hm[sgm][c] = -INFINITY;
  for (int r = warp; r < cnt; r += kTile / 32) {
    const int sgm = r < split ? 0 : 1;
    const int A = sgm ? cl1 : cl0;
    if (A < 0) continue;  // warp-uniform
    double z[kKmDims / 32];
#pragma unroll
    for (int j = 0; j < kKmDims / 32; ++j) {
      const int k = j * 32 + lane;
      z[j] = k < nd ? X[(t0 + r) * d + ssel[k]] - smu[k] : 0.0;
    }
    double acc[CPL];
#pragma unroll
    for (int c = 0; c < CPL; ++c) acc[c] = 0.0;
    double zn = 0.0;  // |z|^2 over the selected features (scales the rounding bound of the projections)
#pragma unroll
    for (int j = 0; j < kKmDims / 32; ++j) {
      zn = fma(z[j], z[j], zn);
      if (j * 32 < nd) {
        const int lim = min(32, nd - j * 32);
        for (int l = 0; l < lim; ++l) {
          const double zk = __shfl_sync(0xffffffffu, z[j], l);
          const double *row = scen + (size_t)(j * 32 + l) * C + lane;
#pragma unroll
          for (int c = 0; c < CPL; ++c) acc[c] = fma(zk, row[32 * c], acc[c]);
        }
      }
    }
    for (int o = 16; o > 0; o >>= 1) zn += __shfl_xor_sync(0xffffffffu, zn, o);
    zn = sqrt(zn);
    double gA = 0.0;
#pragma unroll
    for (int c = 0; c < CPL; ++c) {
      const double v = __shfl_sync(0xffffffffu, acc[c], A & 31);
      if (c == (A >> 5)) gA = v;
    }
    const double nA = cl_norm[A];
#pragma unroll
    for (int c = 0; c < CPL; ++c) {
      const int B = lane + 32 * c;
      const double D = cl_dist[(size_t)A * C + B];
      const double nB = cl_norm[B];
      double pr = (acc[c] - gA - 0.5 * (nB - nA)) / D;
      // Rounding of the numerator is ~(nd + 2) u (|z| (|cA| + |cB|) + |cA|^2 + |cB|^2), amplified by 1 / D: the
      // stored maximum is an UPPER bound of the projection, so the bound is added here (8x safety) instead
      // of relying on a slack proportional to |pr| at the point of use.
      pr += 8.0 * (double)(nd + 2) * 1.1102230246251565e-16 * (zn * (sqrt(nA) + sqrt(nB)) + nA + nB) / D;
      // (nearly) coincident centroids: the quotient amplifies rounding -> no bound for this pair
      if (B == A || !(D > 1e-3 * (sqrt(nA) + sqrt(nB) + 1.0)) || !(pr == pr)) pr = INFINITY;
      if (sgm == 0)
        hm[0][c] = fmax(hm[0][c], pr);
      else
        hm[1][c] = fmax(hm[1][c], pr);
    }
  }
#pragma unroll
  for (int sgm = 0; sgm < 2; ++sgm)
#pragma unroll
    for (int c = 0; c < CPL; ++c) shi[((size_t)warp * 2 + sgm) * C + lane + 32 * c] = hm[sgm][c];
  __syncthreads();
  for (int i = tid; i < 2 * C; i += kTile) {
    double m = -INFINITY;
    for (int w = 0; w < kTile / 32; ++w) m = fmax(m, shi[(size_t)w * 2 * C + i]);
    tile_hi[(size_t)t * 2 * C + i] = m;
  }
  if (tid == 0) {
    tile_cl[t * 2] = cl0;
    tile_cl[t * 2 + 1] = cl1;
  }
}

// Largest emit radius of the rows of a tile that belong to this call: r_i^2 = n_i - 2 key2_i is the squared
// distance below which pass 2 emits (error margins included), and an upper bound of every distance the
// search still needs for row i.
__global__ void __launch_bounds__(kTile) tile_radius_kernel(const float *__restrict__ key2, const double *__restrict__ norm,
                                                            int64_t n, int64_t row_begin, int64_t row_end,
                                                            double *__restrict__ tile_rad) {
  __shared__ double sh[kTile / 32];
  const int64_t row = (int64_t)blockIdx.x * kTile + threadIdx.x;
  double t2 = -1.0;  // no local row in this tile
  if (row >= row_begin && row < row_end && row < n) {
    const float k = key2[row];
    t2 = isfinite(k) ? norm[row] - 2.0 * (double)k : INFINITY;
    if (!(t2 >= 0.0)) t2 = (t2 != t2) ? INFINITY : 0.0;
  }
  for (int o = 16; o > 0; o >>= 1) t2 = fmax(t2, __shfl_xor_sync(0xffffffffu, t2, o));
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = t2;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < kTile / 32; ++w) t2 = fmax(t2, sh[w]);
    tile_rad[blockIdx.x] = t2 < 0.0 ? -1.0 : sqrt(t2) * (1.0 + 1e-9);
  }
}

// One block per 256-row group: test every column tile, keep the survivors in ascending order.
// kind 0: tiles within +-window (no test); 1: radius test without the window tiles; 2: radius test.
__global__ void __launch_bounds__(kTile) tile_lists_kernel(const double *__restrict__ ball_c,
                                                           const double *__restrict__ ball_rho,
                                                           const double *__restrict__ tile_rad,
                                                           const int32_t *__restrict__ tile_cl,
                                                           const double *__restrict__ tile_hi, int C, int64_t d,
                                                           int n_tiles, int g0, int kind, int window, int rt_begin,
                                                           int rt_end, int sort_pow2,
                                                           int32_t *__restrict__ list, int32_t *__restrict__ len,
                                                           unsigned long long *__restrict__ steps) {
  // sort_pow2 > 0: the kept tiles are re-ordered by their lower bound, closest first (dynamic shared memory:
  // sort_pow2 float keys + sort_pow2 int tiles).  Pass 1 then meets a row's true neighbours in its first chunk
  // and the bound it publishes keeps the later chunks off the insertion path.
  extern __shared__ unsigned char tl_sm[];
  float *skey = reinterpret_cast<float *>(tl_sm);
  int *sval = reinterpret_cast<int *>(tl_sm + (size_t)sort_pow2 * sizeof(float));
  __shared__ int s_warp[kTile / 32];
  __shared__ int s_base;
  const int gl = blockIdx.x, g = g0 + gl;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  int32_t *out = list + (size_t)gl * n_tiles;
  if (tid == 0) s_base = 0;
  __syncthreads();
  const double rad = kind == 0 ? 0.0 : tile_rad[g];
  const double rr0 = ball_rho[(size_t)g * 2], rr1 = ball_rho[(size_t)g * 2 + 1];
  const double *cr = ball_c + (size_t)g * 2 * d;
  const int rA0 = tile_cl ? tile_cl[(size_t)g * 2] : -1, rA1 = tile_cl ? tile_cl[(size_t)g * 2 + 1] : -1;
  for (int base = 0; base < n_tiles; base += kTile) {
    const int c = base + tid;
    bool keep = false;
    float lbf = 0.f;
    if (c < n_tiles) {
      const bool in_window = c >= g - window && c <= g + window;
      if (kind == 0) {
        keep = in_window;
      } else if (rad >= 0.0 && !(kind == 1 && in_window)) {
        const double *cc = ball_c + (size_t)c * 2 * d;
        const double rc0 = ball_rho[(size_t)c * 2], rc1 = ball_rho[(size_t)c * 2 + 1];
        for (int a = 0; a < 2 && !keep; ++a) {
          const double ra = a ? rr1 : rr0;
          if (ra < 0.0) continue;
          for (int b = 0; b < 2 && !keep; ++b) {
            const double rb = b ? rc1 : rc0;
            if (rb < 0.0) continue;
            // cells of different clusters A != B: |x - y| >= -p_AB(x) - p_BA(y) (projection on the centroid axis)
            const int A = a ? rA1 : rA0;
            const int B = tile_cl ? tile_cl[(size_t)c * 2 + b] : -1;
            if (A >= 0 && B >= 0 && A != B) {
              const double h1 = tile_hi[((size_t)g * 2 + a) * C + B], h2 = tile_hi[((size_t)c * 2 + b) * C + A];
              const double lbp = -h1 - h2;
              // proven too far for this pair of segments (slack: float64 rounding of the projections)
              if (lbp - 1e-9 * (fabs(h1) + fabs(h2) + 1.0) > rad) continue;
            }
            const double *pa = cr + (size_t)a * d, *pb = cc + (size_t)b * d;
            double acc = 0.0;
            for (int64_t k = 0; k < d; ++k) {
              const double diff = pa[k] - pb[k];
              acc = fma(diff, diff, acc);
            }
            // rows x of ball a, columns y of ball b: |x - y| >= |ca - cb| - ra - rb
            const double lb = sqrt(acc) * (1.0 - 1e-12) - ra - rb;
            keep = !(lb > rad);  // rad = inf keeps everything
            lbf = (float)lb;
          }
        }
      }
    }
    const unsigned m = __ballot_sync(0xffffffffu, keep);
    if (lane == 0) s_warp[warp] = __popc(m);
    __syncthreads();
    int off = s_base;
    for (int w = 0; w < warp; ++w) off += s_warp[w];
    if (keep) {
      const int pos = off + __popc(m & ((1u << lane) - 1u));
      out[pos] = c;
      if (sort_pow2 > 0) {
        skey[pos] = lbf;
        sval[pos] = c;
      }
    }
    __syncthreads();
    if (tid == 0) {
      int tot = 0;
      for (int w = 0; w < kTile / 32; ++w) tot += s_warp[w];
      s_base += tot;
    }
    __syncthreads();
  }
  if (sort_pow2 > 0 && s_base > 1) {  // block-uniform
    const int nk = s_base;
    int size = 1;
    while (size < nk) size <<= 1;
    for (int i = nk + tid; i < size; i += kTile) {
      skey[i] = INFINITY;
      sval[i] = 0x7fffffff;
    }
    __syncthreads();
    for (int k = 2; k <= size; k <<= 1) {  // bitonic sort, ascending by (bound, tile)
      for (int j = k >> 1; j > 0; j >>= 1) {
        for (int i = tid; i < size; i += kTile) {
          const int ixj = i ^ j;
          if (ixj > i) {
            const float ka = skey[i], kb = skey[ixj];
            const int va = sval[i], vb = sval[ixj];
            const bool gt = ka > kb || (ka == kb && va > vb);
            const bool up = (i & k) == 0;
            if (gt == up) {
              skey[i] = kb;
              skey[ixj] = ka;
              sval[i] = vb;
              sval[ixj] = va;
            }
          }
        }
        __syncthreads();
      }
    }
    for (int i = tid; i < nk; i += kTile) out[i] = sval[i];
  }
  if (tid == 0) {
    len[gl] = s_base;
    // row tiles of this group inside the call's range (flop accounting)
    int nrt = 0;
    for (int rt = g * (BN / BM); rt < (g + 1) * (BN / BM); ++rt) nrt += (rt >= rt_begin && rt < rt_end) ? 1 : 0;
    atomicAdd(steps, (unsigned long long)s_base * (unsigned long long)nrt);
  }
}

// ---- host side ------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static int encode_operand_map(void *base, int64_t rows, int kp, int box_rows, unsigned char *out) {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void *p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    MELD_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres));
    if (!p || qres != cudaDriverEntryPointSuccess) {
      set_error("cuTensorMapEncodeTiled is not available from this driver");
      return MELD_B200_ERR_CUDA;
    }
    fn = (EncodeTiledFn)p;
  }
  CUtensorMap m;
  const cuuint64_t gdim[2] = {(cuuint64_t)kp, (cuuint64_t)rows};
  const cuuint64_t gstride[1] = {(cuuint64_t)kp * 2};
  const cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)box_rows};
  const cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, base, gdim, gstride, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
    return MELD_B200_ERR_CUDA;
  }
  static_assert(sizeof(CUtensorMap) == 128, "CUtensorMap size");
  memcpy(out, &m, sizeof(m));
  return 0;
}

int tc_plan(int64_t n, int64_t d, int k1, int64_t row_begin, int64_t row_end, SearchPlan *plan) {
  MELD_REQUIRE(row_begin % BM == 0 && (row_end % BM == 0 || row_end == n) && row_begin < row_end && row_end <= n,
               "candidate search: row range [%lld, %lld) must be aligned to %d rows", (long long)row_begin,
               (long long)row_end, BM);
  plan->simt = false;
  plan->row_begin = row_begin;
  plan->row_end = row_end;
  plan->n = n;
  plan->d = d;
  plan->k1 = k1;
  plan->terms = 3;
  const int64_t kp = round_up(3 * d + 3, BK);
  MELD_REQUIRE(kp <= (1 << 20), "knn_graph_build: d=%lld too large for the tensor-core search", (long long)d);
  plan->kp = (int)kp;
  plan->kp_used = (int)kp;
  plan->n_pad = round_up(n, 4 * BM);  // row tiles come in multiples of the largest cluster size
  plan->n_pad_cols = plan->n_pad;
  const int64_t row_tiles = ceil_div(row_end - row_begin, BM), col_tiles = plan->n_pad / BN;
  // enough (row tile, segment) units for ~4 waves of the SMs, never more segments than column tiles;
  // and a segment's slice of the column operand (n_seg x K' bf16) should stay L2 resident while the
  // CTAs of that segment stream it (units are issued segment-major)
  int64_t nseg = ceil_div(4 * (int64_t)sm_count(), row_tiles);
  const int64_t l2_slice = (int64_t)40 << 20;
  const int64_t nseg_l2 = ceil_div(plan->n_pad * kp * 2, l2_slice);
  if (nseg < nseg_l2) nseg = nseg_l2;
  if (nseg > kMaxLists / 2) nseg = kMaxLists / 2;  // two lists per segment
  if (nseg > col_tiles) nseg = col_tiles;
  if (nseg < 1) nseg = 1;
  plan->nseg = (int)nseg;
  plan->nlists = (int)nseg;
  if (tuning().p1_segments > 0 && tuning().p1_segments < plan->nlists) plan->nlists = tuning().p1_segments;
  plan->nlists *= 2;  // two epilogue threads (column halves) per row keep separate lists
  plan->prune = tuning().prune != 0;
  if (plan->prune) {
    // list-driven passes: the window pass writes list slot 0, the list pass slots 1 .. nchunk
    // units per row tile: pruned lists are short (~100 tiles at 500k cells), and every unit pays an A-tile
    // reload and a pipeline refill, so fewer, longer units than the unpruned segment count
    int64_t nchunk = nseg;
    const int64_t cap = tuning().tl_chunks > 0 ? tuning().tl_chunks : 2;
    if (nchunk > cap) nchunk = cap;
    if (nchunk > kMaxLists / 2 - 1) nchunk = kMaxLists / 2 - 1;
    plan->nchunk = (int)nchunk;
    plan->nlists = 2 * (1 + (int)nchunk);
    plan->window = tuning().prune_window > 0 ? tuning().prune_window : 0;
  }
  // bf16 split error of x.y (~2^-15.8 |x||y|) plus fp32 accumulation, x2 for d^2, 4x safety
  plan->margin_c = ldexp(1.0, -11);
  return 0;
}

int tc_prepare(const SearchPlan &plan, const double *X, const double *mu, const double *norm, cudaStream_t stream,
               SearchState *st) {
  MELD_CHECK(st->thr_g.alloc((size_t)plan.n_pad));
  {
    // -inf bit pattern 0xff800000 in every float
    MELD_CUDA(cudaMemsetAsync(st->thr_g.p, 0, (size_t)plan.n_pad * sizeof(float), stream));
    fill_float_kernel<<<sm_count() * 4, 256, 0, stream>>>(st->thr_g.p, plan.n_pad, -INFINITY);
    MELD_LAUNCH_CHECK();
  }
  MELD_CHECK(st->a_op.alloc((size_t)plan.n_pad * plan.kp));
  MELD_CHECK(st->b_op.alloc((size_t)plan.n_pad * plan.kp));
  tc_prep_kernel<<<sm_count() * 8, 256, 0, stream>>>(X, mu, norm, plan.n, plan.n_pad, plan.d, plan.kp,
                                                     reinterpret_cast<__nv_bfloat16 *>(st->a_op.p),
                                                     reinterpret_cast<__nv_bfloat16 *>(st->b_op.p));
  MELD_LAUNCH_CHECK();
  MELD_CHECK(encode_operand_map(st->a_op.p, plan.n_pad, plan.kp, BM, st->tmap_a));
  return 0;
}


int tc_tile_balls(const SearchPlan &plan, const double *X, const CellClusters *cl, cudaStream_t stream, SearchState *st) {
  const int32_t *cid = (cl && cl->cid.p) ? cl->cid.p : nullptr;
  const int64_t T = plan.n_pad / kTile;
  st->n_tiles = T;
  st->g0 = plan.row_begin / kTile;
  const int64_t row_hi = plan.row_end == plan.n ? plan.n_pad : plan.row_end;
  st->n_groups = ceil_div(row_hi, kTile) - st->g0;
  MELD_REQUIRE(T < (int64_t)1 << 30 && st->n_groups >= 1, "tc_search: bad tile counts");
  MELD_CHECK(st->ball_c.alloc((size_t)T * 2 * plan.d));
  MELD_CHECK(st->ball_rho.alloc((size_t)T * 2));
  MELD_CHECK(st->tile_rad.alloc((size_t)T));
  MELD_CHECK(st->tl_list.alloc((size_t)st->n_groups * T));
  MELD_CHECK(st->tl_len.alloc((size_t)st->n_groups));
  MELD_CHECK(st->tl_steps.alloc(4));
  MELD_CUDA(cudaMemsetAsync(st->tl_steps.p, 0, 4 * sizeof(unsigned long long), stream));
  tile_balls_kernel<<<(unsigned)T, kTile, 0, stream>>>(X, plan.n, plan.d, cid, st->ball_c.p, st->ball_rho.p);
  MELD_LAUNCH_CHECK();
  st->n_clusters = 0;
  if (cid && cl->C >= 32 && cl->C % 32 == 0 && cl->C <= kKmMaxC && tuning().prune_proj) {
    const int C = cl->C, nd = cl->nd;
    MELD_CHECK(st->cl_norm.alloc((size_t)C));
    MELD_CHECK(st->cl_dist.alloc((size_t)C * C));
    MELD_CHECK(st->tile_cl.alloc((size_t)T * 2));
    MELD_CHECK(st->tile_hi.alloc((size_t)T * 2 * C));
    cluster_tables_kernel<<<C, 128, 0, stream>>>(cl->cen.p, nd, C, st->cl_norm.p, st->cl_dist.p);
    MELD_LAUNCH_CHECK();
    const size_t smem = ((size_t)nd * C + (size_t)(kTile / 32) * 2 * C + kKmDims) * sizeof(double) + kKmDims * sizeof(int);
    auto proj = C == 32 ? tile_proj_kernel<1> : C == 64 ? tile_proj_kernel<2> : C == 96 ? tile_proj_kernel<3>
                                                                                       : tile_proj_kernel<4>;
    MELD_CUDA(cudaFuncSetAttribute(proj, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    proj<<<(unsigned)T, kTile, smem, stream>>>(X, plan.n, plan.d, cid, cl->sel.p, nd, cl->mu.p, cl->cen.p, C,
                                               st->cl_norm.p, st->cl_dist.p, st->tile_cl.p, st->tile_hi.p);
    MELD_LAUNCH_CHECK();
    st->n_clusters = C;
  }
  return 0;
}

int tc_tile_lists(const SearchPlan &plan, SearchState &st, int kind, const float *key2, const double *norm, int counter,
                  cudaStream_t stream, TileLists *out) {
  MELD_REQUIRE(st.n_tiles > 0 && counter >= 0 && counter < 4, "tc_search: tile balls missing");
  if (kind != 0) {
    MELD_REQUIRE(key2 && norm, "tc_search: tile lists need the emit keys");
    tile_radius_kernel<<<(unsigned)st.n_tiles, kTile, 0, stream>>>(key2, norm, plan.n, plan.row_begin, plan.row_end,
                                                                   st.tile_rad.p);
    MELD_LAUNCH_CHECK();
  }
  const int rt_begin = (int)(plan.row_begin / BM);
  const int rt_end =
      rt_begin + (int)((plan.row_end == plan.n ? plan.n_pad - plan.row_begin : plan.row_end - plan.row_begin) / BM);
  MELD_CUDA(cudaMemsetAsync(st.tl_steps.p + counter, 0, sizeof(unsigned long long), stream));
  int sort_pow2 = 0;
  if (kind == 1 && tuning().tl_sort) {
    sort_pow2 = 1;
    while (sort_pow2 < st.n_tiles) sort_pow2 <<= 1;
    if ((size_t)sort_pow2 * 8 > 200 * 1024) sort_pow2 = 0;  // does not fit shared memory: keep the index order
  }
  const size_t smem = (size_t)sort_pow2 * 8;
  if (smem > 48 * 1024)
    MELD_CUDA(cudaFuncSetAttribute(tile_lists_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  tile_lists_kernel<<<(unsigned)st.n_groups, kTile, smem, stream>>>(
      st.ball_c.p, st.ball_rho.p, st.tile_rad.p, st.n_clusters ? st.tile_cl.p : nullptr,
      st.n_clusters ? st.tile_hi.p : nullptr, st.n_clusters, plan.d, (int)st.n_tiles, (int)st.g0, kind, plan.window,
      rt_begin, rt_end, sort_pow2, st.tl_list.p, st.tl_len.p, st.tl_steps.p + counter);
  MELD_LAUNCH_CHECK();
  out->list = st.tl_list.p;
  out->len = st.tl_len.p;
  out->stride = (int)st.n_tiles;
  out->g0 = (int)st.g0;
  return 0;
}

int tc_pass(const SearchPlan &plan, SearchState &st, int mode, float *lists, const float *key2,
            unsigned long long *pairs, unsigned long long *count, int64_t cap, cudaStream_t stream,
            const TileLists *tl) {
  TcArgs a{};
  a.mode = mode;
  a.n = plan.n;
  a.rt_begin = (int)(plan.row_begin / BM);
  a.row_end = (long long)plan.row_end;
  // whole-data calls cover the padded tiles too (keeps tile counts even for clusters)
  a.n_row_tiles = (int)((plan.row_end == plan.n ? plan.n_pad - plan.row_begin : plan.row_end - plan.row_begin) / BM);
  a.n_col_tiles = (int)(plan.n_pad / BN);
  a.nseg = plan.nseg;
  a.nkb = plan.kp / BK;
  a.a_resident = a.nkb <= kMaxResidentKb ? 1 : 0;
  a.k1 = plan.k1;
  a.lists = lists;
  a.thr_g = st.thr_g.p;
  a.key2 = key2;
  a.pairs = pairs;
  a.count = count;
  a.cap = (long long)cap;
  const size_t list_bytes = mode == 1 ? (size_t)plan.k1 * 2 * BM * 4 : (size_t)kEpiWarps * kQueueCap * 8 + 64;
  const size_t fixed = 1024 + list_bytes + sizeof(Bars);
  const size_t budget = 227 * 1024;
  if (a.a_resident) {
    a.a_region = a.nkb * kABytes;
    a.n_stages = (int)((budget - fixed - a.a_region) / kBBytes);
  } else {
    a.n_stages = (int)((budget - fixed) / (kBBytes + kABytes));
    a.a_region = a.n_stages * kABytes;
  }
  if (a.n_stages > kMaxStages) a.n_stages = kMaxStages;
  MELD_REQUIRE(a.n_stages >= 2, "tc_search: shared memory too small for a pipeline (knn too large)");
  const size_t smem = fixed + a.a_region + (size_t)a.n_stages * kBBytes;
  typedef void (*TcKernel)(const CUtensorMap, const CUtensorMap, const TcArgs);
  TcKernel kern = tc_search_kernel<0>;
  if (mode == 1 && tuning().reg_topk) {
    if (plan.k1 <= 8) kern = tc_search_kernel<8>;
    else if (plan.k1 <= 16) kern = tc_search_kernel<16>;
    else if (plan.k1 <= 32) kern = tc_search_kernel<32>;
  }
  MELD_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  a.n_passes = mode == 1 ? plan.nlists / 2 : a.nseg;  // pass 1 visits nlists/2 segments per row (own first)
  a.n_slots = plan.nlists / 2;
  a.slot0 = 0;
  if (tl) {
    MELD_REQUIRE(tl->list && tl->len && tl->chunks >= 1 && tl->slot0 + (mode == 1 ? tl->chunks : 0) <= a.n_slots,
                 "tc_search: bad tile lists");
    a.tl_list = tl->list;
    a.tl_len = tl->len;
    a.tl_stride = tl->stride;
    a.tl_g0 = tl->g0;
    a.n_passes = tl->chunks;
    a.slot0 = tl->slot0;
    a.interleave = mode == 1 ? (tuning().tl_interleave & 1) : ((tuning().tl_interleave >> 1) & 1);
  }
  const int units = a.n_row_tiles * a.n_passes;
  // cluster size: B tiles are loaded once per cluster (TMA multicast), which divides the L2 -> SM traffic of
  // the streamed operand -- the kernel's real bound at K' = 320 -- by the cluster size
  int cs = tuning().tc_multicast;
  if (cs != 2 && cs != 4) cs = 1;
  if (tl && cs > 2) cs = 2;  // lists are per 256-row group = the two row tiles of a 2-CTA cluster
  while (cs > 1 && (a.n_row_tiles % cs != 0 || a.rt_begin % cs != 0 || units < cs)) cs >>= 1;
  cudaLaunchConfig_t cfg{};
  cfg.blockDim = dim3(kTcThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = (unsigned)cs;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = cs > 1 ? 1 : 0;
  int grid = sm_count();
  if (cs > 1) {
    // a persistent kernel must be fully co-resident: ask how many clusters fit at once (GPC shapes can
    // strand SMs for cluster size 4)
    cfg.gridDim = dim3((unsigned)(sm_count() / cs * cs));
    int max_clusters = 0;
    MELD_CUDA(cudaOccupancyMaxActiveClusters(&max_clusters, kern, &cfg));
    if (max_clusters < 1) {
      cs = 1;
      cfg.numAttrs = 0;
    } else {
      grid = max_clusters * cs;
      if (grid > sm_count() / cs * cs) grid = sm_count() / cs * cs;
    }
  }
  if (grid > units) grid = units / cs * cs;
  if (grid < cs) grid = cs;
  a.mc = cs;
  cfg.gridDim = dim3((unsigned)grid);
  CUtensorMap ma, mb;
  memcpy(&ma, st.tmap_a, sizeof(ma));
  {
    unsigned char tm[128];
    MELD_CHECK(encode_operand_map(st.b_op.p, plan.n_pad, plan.kp, BN / cs, tm));  // one slice of a B tile per CTA
    memcpy(&mb, tm, sizeof(mb));
  }
  MELD_CUDA(cudaLaunchKernelEx(&cfg, kern, ma, mb, a));
  MELD_LAUNCH_CHECK();
  return 0;
}

}  // namespace meld
