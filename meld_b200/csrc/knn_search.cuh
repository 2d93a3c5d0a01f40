// Candidate search interface shared by the tcgen05 and the SIMT implementations (knn.cu drives it).
//
// Both work in "s-space": s_ij = x_i . x_j - n_j / 2 on centred data (n_j = |x_j|^2), so that
// d_ij^2 = n_i - 2 s_ij and a larger s means a closer neighbour.
//   pass 1: for every row, nlists sorted (descending) lists of the k1 largest s seen by each
//           (column segment x epilogue half) -> merged by merge_lists_kernel into key2_i;
//   pass 2: append a pair (i << 32 | j) for every column j with s_ij >= key2_i to one global buffer
//           (unordered; *count keeps growing past the capacity so the caller can size a retry).
// margin_c bounds the implementation's error: |d2_approx - d2_exact| <= margin_c (n_i + max_j n_j).
#pragma once
#include "common.cuh"

namespace meld {

constexpr int kMaxK1 = 64;     // knn + 1 limit (shared-memory top-k lists)
constexpr int kMaxLists = 128;  // lists per row merged after pass 1

struct SearchPlan {
  bool simt = false;
  int64_t n = 0, d = 0;
  int64_t row_begin = 0, row_end = 0;  // query rows handled by this call (columns are always all n cells)
  int k1 = 0;
  int nseg = 1;    // column segments (independent work units per row tile)
  int nlists = 1;  // lists per row written by pass 1
  int64_t n_pad_cols = 0;  // columns actually multiplied per pass (padded) and K used, for flop accounting
  int kp_used = 0;
  double margin_c = 0.0;
  // tcgen05 path
  int64_t n_pad = 0;  // rows padded to the tile size
  int kp = 0;         // padded K of the bf16 operand (3 d + 3 rounded up to 64)
  int terms = 3;      // bf16 split terms (3: hi*hi + hi*lo + lo*hi)
  // pruned (ball-tree style) search: column tiles per 256-row group come from lists built with the triangle
  // inequality on per-tile bounding balls; pass 1 = a window pass (slot 0) + a list pass (slots 1..nchunk)
  bool prune = false;
  int nchunk = 1;   // units per row tile in a list-driven pass
  int window = 2;   // the window pass scans the column tiles within +-window of the row group's own tile
};

// k-means clusters of the internal cell order (filled by cell_order in knn.cu; empty when clustering is off)
constexpr int kKmDims = 128;   // features used for clustering (highest variance first)
constexpr int kKmMaxC = 128;
struct CellClusters {
  DevBuf<int32_t> cid;  // [n] cluster of every position of the internal order, non-decreasing
  DevBuf<float> cen;    // [nd][C] centroids of the centred data on the selected features
  DevBuf<int32_t> sel;  // [kKmDims] the selected features
  DevBuf<double> mu;    // [d] column means used for centring
  int nd = 0, C = 0;
};

// Which column tiles a list-driven tc_pass visits (device arrays owned by SearchState).
struct TileLists {
  const int32_t *list = nullptr;  // [groups][stride], ascending tile indices
  const int32_t *len = nullptr;   // [groups]
  int stride = 0, g0 = 0;         // g0: first 256-row group of this call's row range
  int chunks = 1;                 // units per row tile
  int slot0 = 0;                  // pass 1: first list slot this launch writes
};

struct SearchState {
  // SIMT
  DevBuf<float> xc32;  // n x d centred float32
  DevBuf<float> hn32;  // -n_j / 2
  // tcgen05
  DevBuf<float> thr_g;    // n_pad running pass-1 bounds
  DevBuf<uint16_t> a_op;  // n_pad x kp bf16, row operand
  DevBuf<uint16_t> b_op;  // n_pad x kp bf16, column operand (carries -n_j/2 in its tail columns)
  alignas(64) unsigned char tmap_a[128];
  alignas(64) unsigned char tmap_b[128];
  // pruned search
  DevBuf<double> ball_c;     // [tiles][2][d] centres of the (up to) two bounding balls of a 256-cell tile
  DevBuf<double> ball_rho;   // [tiles][2] radii (-1: empty ball)
  DevBuf<double> tile_rad;   // [tiles] largest emit radius of a tile's rows
  // projection bound between cells of different clusters A != B: with w = (c_B - c_A)/|c_B - c_A| and m the
  // midpoint, |x - y| >= w.(y - x) = -p_AB(x) - p_BA(y), p_AB(x) = w.(x - m); tile_hi holds max p per segment
  DevBuf<int32_t> tile_cl;   // [tiles][2] cluster of a tile's segment (-1: empty or mixed)
  DevBuf<double> tile_hi;    // [tiles][2][C]
  DevBuf<double> cl_norm;    // [C] |c_A|^2
  DevBuf<double> cl_dist;    // [C][C] |c_A - c_B|
  int n_clusters = 0;
  DevBuf<int32_t> tl_list, tl_len;
  DevBuf<unsigned long long> tl_steps;  // [4] (row tile, column tile) products issued by passes 0, 1, 2
  int64_t n_tiles = 0, g0 = 0, n_groups = 0;
};

int search_plan(bool simt, int64_t n, int64_t d, int k1, int64_t row_begin, int64_t row_end, SearchPlan *plan);
int search_prepare(const SearchPlan &plan, const double *X, const double *mu, const double *norm, cudaStream_t stream,
                   SearchState *st);
int search_pass1(const SearchPlan &plan, SearchState &st, float *lists, cudaStream_t stream);
int search_pass2(const SearchPlan &plan, SearchState &st, const float *key2, unsigned long long *pairs,
                 unsigned long long *count, int64_t cap, cudaStream_t stream);
void search_release(SearchState *st);

// tcgen05 implementation (knn_tc.cu)
int tc_plan(int64_t n, int64_t d, int k1, int64_t row_begin, int64_t row_end, SearchPlan *plan);
int tc_prepare(const SearchPlan &plan, const double *X, const double *mu, const double *norm, cudaStream_t stream,
               SearchState *st);
int tc_pass(const SearchPlan &plan, SearchState &st, int mode, float *lists, const float *key2,
            unsigned long long *pairs, unsigned long long *count, int64_t cap, cudaStream_t stream,
            const TileLists *tl = nullptr);
// pruned search helpers (knn_tc.cu).  X is in internal cell order; cl (optional) describes the k-means clusters of
// that order -- a tile that straddles a cluster boundary gets one ball per side.
int tc_tile_balls(const SearchPlan &plan, const double *X, const CellClusters *cl, cudaStream_t stream, SearchState *st);
// kind 0: window lists; 1: radius test from key2 without the window tiles; 2: radius test, all tiles.
// counter: which tl_steps slot accumulates the (row tile x column tile) products of the pass that will use the lists.
int tc_tile_lists(const SearchPlan &plan, SearchState &st, int kind, const float *key2, const double *norm, int counter,
                  cudaStream_t stream, TileLists *out);

}  // namespace meld
