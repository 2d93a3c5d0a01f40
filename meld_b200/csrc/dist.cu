// Peer-memory plumbing of the row-partitioned (multi-GPU) filter and row slices of a graph.
//
// The reference has no distributed code (setup.py:45); the operation that is sharded is the
// recurrence behind meld/filter.py:59.  One process per GPU: every rank allocates one block with cudaMalloc,
// exports it with cudaIpcGetMemHandle, the host side (torch.distributed, plumbing) all-gathers the 64-byte
// handles, and every rank maps the other blocks with cudaIpcOpenMemHandle.  After that the data path never
// touches the host or NCCL: the step kernel stores its rows of T_k into every peer's buffer (cheby.cu, PEER
// variant) and the ranks meet through flag words in each other's memory.
#include "common.cuh"

#include <string.h>

using namespace meld;

namespace meld {

__global__ void slice_row_ptr_kernel(const int32_t *__restrict__ row_ptr, int64_t row_begin, int64_t n_rows,
                                     int32_t *__restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i <= n_rows) out[i] = row_ptr[row_begin + i] - row_ptr[row_begin];
}

__global__ void mark_columns_kernel(const int32_t *__restrict__ col, int64_t nnz, uint8_t *__restrict__ ref) {
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < nnz; e += (int64_t)gridDim.x * blockDim.x)
    ref[col[e]] = 1;
}

// recv[w * chunk + i] = 1 when rank w references this rank's row i; bit k of mask[i] <- the k-th peer (rank skipped)
__global__ void halo_mask_kernel(const uint8_t *__restrict__ recv, int64_t chunk, int64_t nloc, int world, int rank,
                                 uint8_t *__restrict__ mask) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nloc; i += (int64_t)gridDim.x * blockDim.x) {
    unsigned m = 0;
    int k = 0;
    for (int w = 0; w < world; ++w) {
      if (w == rank) continue;
      if (recv[(int64_t)w * chunk + i]) m |= 1u << k;
      ++k;
    }
    mask[i] = (uint8_t)m;
  }
}

}  // namespace meld

extern "C" {

int meld_b200_graph_mark_columns(const meld_b200_graph_t *g, uint8_t *ref, void *stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  meld::use_stream(stream);
  MELD_REQUIRE(g && ref, "graph_mark_columns: NULL argument");
  if (g->nnz > 0) {
    mark_columns_kernel<<<sm_count() * 4, 256, 0, stream>>>(g->col.p, g->nnz, ref);
    MELD_LAUNCH_CHECK();
  }
  return 0;
}

int meld_b200_graph_set_halo(meld_b200_graph_t *g, const uint8_t *recv, int64_t chunk, int world, int rank, void *stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  meld::use_stream(stream);
  MELD_REQUIRE(g && recv && world >= 1 && world <= 8 && rank >= 0 && rank < world && chunk >= g->n_rows,
               "graph_set_halo: bad argument");
  MELD_CHECK(g->halo.alloc((size_t)(g->n_rows > 0 ? g->n_rows : 1)));
  if (g->n_rows > 0) {
    halo_mask_kernel<<<sm_count() * 2, 256, 0, stream>>>(recv, chunk, g->n_rows, world, rank, g->halo.p);
    MELD_LAUNCH_CHECK();
  }
  return 0;
}

int meld_b200_graph_row_slice(const meld_b200_graph_t *g, int64_t row_begin, int64_t row_end, void *stream_,
                              meld_b200_graph_t **out) {
  cudaStream_t stream = (cudaStream_t)stream_;
  meld::use_stream(stream);
  MELD_REQUIRE(g && out, "graph_row_slice: NULL argument");
  *out = nullptr;
  MELD_REQUIRE(g->row0 == 0 && g->n_rows == g->n_cols, "graph_row_slice: needs the full operator");
  MELD_REQUIRE(row_begin >= 0 && row_begin <= row_end && row_end <= g->n_rows, "graph_row_slice: bad range [%lld, %lld)",
               (long long)row_begin, (long long)row_end);
  const int64_t nloc = row_end - row_begin;
  int32_t h_ends[2] = {0, 0};
  MELD_CUDA(cudaMemcpyAsync(&h_ends[0], g->row_ptr.p + row_begin, sizeof(int32_t), cudaMemcpyDeviceToHost, stream));
  MELD_CUDA(cudaMemcpyAsync(&h_ends[1], g->row_ptr.p + row_end, sizeof(int32_t), cudaMemcpyDeviceToHost, stream));
  MELD_SYNC(stream);
  const int64_t e0 = h_ends[0], nnz = (int64_t)h_ends[1] - h_ends[0];
  meld_b200_graph *s = new (std::nothrow) meld_b200_graph();
  if (!s) {
    set_error("graph_row_slice: host allocation failed");
    return MELD_B200_ERR_NOMEM;
  }
  struct Guard {
    meld_b200_graph *g;
    ~Guard() { delete g; }
  } guard{s};
  s->n_rows = nloc;
  s->n_cols = g->n_cols;
  s->row0 = row_begin;
  s->nnz = nnz;
  MELD_CHECK(s->row_ptr.alloc((size_t)nloc + 1 + kCsrPad));
  MELD_CHECK(s->col.alloc((size_t)nnz + kCsrPad));
  MELD_CHECK(s->val.alloc((size_t)nnz + kCsrPad));
  MELD_CUDA(cudaMemsetAsync(s->row_ptr.p + nloc + 1, 0, kCsrPad * sizeof(int32_t), stream));
  MELD_CUDA(cudaMemsetAsync(s->col.p + nnz, 0, kCsrPad * sizeof(int32_t), stream));
  MELD_CUDA(cudaMemsetAsync(s->val.p + nnz, 0, kCsrPad * sizeof(double), stream));
  slice_row_ptr_kernel<<<(unsigned)ceil_div(nloc + 1, 256), 256, 0, stream>>>(g->row_ptr.p, row_begin, nloc, s->row_ptr.p);
  MELD_LAUNCH_CHECK();
  if (nnz > 0) {
    // the flat kernel reads values / columns with 32- / 16-byte vector loads at multiples of 4 entries: the slice
    // starts its own arrays at entry 0, so alignment is that of the allocation
    MELD_CUDA(cudaMemcpyAsync(s->col.p, g->col.p + e0, (size_t)nnz * sizeof(int32_t), cudaMemcpyDeviceToDevice, stream));
    MELD_CUDA(cudaMemcpyAsync(s->val.p, g->val.p + e0, (size_t)nnz * sizeof(double), cudaMemcpyDeviceToDevice, stream));
  }
  if (g->perm.p) {  // the slice converts full-length signals with the full graph's cell order
    MELD_CHECK(s->perm.alloc((size_t)g->n_cols));
    MELD_CUDA(cudaMemcpyAsync(s->perm.p, g->perm.p, (size_t)g->n_cols * sizeof(int32_t), cudaMemcpyDeviceToDevice, stream));
  }
  MELD_CHECK(graph_finalize(s, stream));
  guard.g = nullptr;
  *out = s;
  return 0;
}

int meld_b200_dist_create(int rank, int world, int64_t n_rows_total, int p_max, void *stream_, meld_b200_dist_t **out) {
  meld::use_stream((cudaStream_t)stream_);
  MELD_REQUIRE(out != nullptr, "dist_create: NULL argument");
  *out = nullptr;
  MELD_REQUIRE(world >= 1 && world <= 8 && rank >= 0 && rank < world, "dist_create: rank %d of %d (one box: <= 8 ranks)",
               rank, world);
  MELD_REQUIRE(n_rows_total > 0 && p_max >= 1 && p_max <= 8, "dist_create: n=%lld p_max=%d", (long long)n_rows_total, p_max);
  meld_b200_dist *d = new (std::nothrow) meld_b200_dist();
  if (!d) {
    set_error("dist_create: host allocation failed");
    return MELD_B200_ERR_NOMEM;
  }
  d->rank = rank;
  d->world = world;
  d->n = n_rows_total;
  d->p_max = p_max;
  d->buf_len = (((size_t)n_rows_total * p_max + 2 + 3) & ~(size_t)3) + 32;
  d->bytes = meld_b200_dist::kBufOff + 2 * d->buf_len * sizeof(double);
  cudaError_t e = cudaMalloc((void **)&d->base, d->bytes);  // plain cudaMalloc: pool memory cannot be IPC-exported
  if (e != cudaSuccess) {
    set_error("dist_create: cudaMalloc(%zu) failed: %s", d->bytes, cudaGetErrorString(e));
    cudaGetLastError();
    delete d;
    return MELD_B200_ERR_NOMEM;
  }
  e = cudaMemset(d->base, 0, d->bytes);
  if (e != cudaSuccess) {
    set_error("dist_create: cudaMemset failed: %s", cudaGetErrorString(e));
    cudaFree(d->base);
    delete d;
    return MELD_B200_ERR_CUDA;
  }
  cudaDeviceSynchronize();
  d->peer_base[rank] = d->base;
  d->connected = world == 1;
  *out = d;
  return 0;
}

int meld_b200_dist_handle_bytes(void) { return (int)sizeof(cudaIpcMemHandle_t); }

int meld_b200_dist_export(const meld_b200_dist_t *d, void *blob_host) {
  MELD_REQUIRE(d && blob_host, "dist_export: NULL argument");
  cudaIpcMemHandle_t h;
  MELD_CUDA(cudaIpcGetMemHandle(&h, d->base));
  memcpy(blob_host, &h, sizeof(h));
  return 0;
}

int meld_b200_dist_connect(meld_b200_dist_t *d, const void *all_blobs_host) {
  MELD_REQUIRE(d && all_blobs_host, "dist_connect: NULL argument");
  MELD_REQUIRE(!d->connected || d->world == 1, "dist_connect: already connected");
  for (int r = 0; r < d->world; ++r) {
    if (r == d->rank) continue;
    cudaIpcMemHandle_t h;
    memcpy(&h, (const char *)all_blobs_host + (size_t)r * sizeof(h), sizeof(h));
    void *ptr = nullptr;
    cudaError_t e = cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) {
      set_error("dist_connect: cudaIpcOpenMemHandle(rank %d) failed: %s", r, cudaGetErrorString(e));
      cudaGetLastError();
      return MELD_B200_ERR_CUDA;
    }
    d->peer_base[r] = (char *)ptr;
  }
  d->connected = true;
  return 0;
}

int meld_b200_dist_connect_local(meld_b200_dist_t *d, meld_b200_dist_t *const *all, int count) {
  MELD_REQUIRE(d && all && count == d->world, "dist_connect_local: bad argument");
  for (int r = 0; r < count; ++r) {
    MELD_REQUIRE(all[r] && all[r]->rank == r && all[r]->world == d->world && all[r]->n == d->n &&
                     all[r]->p_max == d->p_max,
                 "dist_connect_local: context %d does not belong to this group", r);
    d->peer_base[r] = all[r]->base;
  }
  d->connected = true;
  return 0;
}

int meld_b200_dist_error(const meld_b200_dist_t *d, int *err_host) {
  MELD_REQUIRE(d && err_host, "dist_error: NULL argument");
  MELD_CUDA(cudaMemcpy(err_host, d->err(), sizeof(int), cudaMemcpyDeviceToHost));
  return 0;
}

int meld_b200_dist_destroy(meld_b200_dist_t *d) {
  if (!d) return 0;
  cudaDeviceSynchronize();
  for (int r = 0; r < d->world; ++r)
    if (r != d->rank && d->peer_base[r]) cudaIpcCloseMemHandle(d->peer_base[r]);
  if (d->base) cudaFree(d->base);
  cudaGetLastError();
  delete d;
  return 0;
}

}  // extern "C"
