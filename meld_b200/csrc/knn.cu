// placeholder
#include "common.cuh"
using namespace meld;
extern "C" {
int meld_b200_knn_graph_build(const double *X, int64_t n, int64_t d, int knn, double decay, double thresh,
                              double anisotropy, double bandwidth_scale, int flags, void *stream,
                              meld_b200_graph_t **graph_out) {
  set_error("knn_graph_build: not built yet");
  return MELD_B200_ERR_UNSUPPORTED;
}
int meld_b200_graph_knn_kernel_nnz(const meld_b200_graph_t *g, int64_t *nnz_host) { return MELD_B200_ERR_UNSUPPORTED; }
int meld_b200_graph_export_knn_kernel(const meld_b200_graph_t *g, int64_t *indptr, int32_t *indices, double *data,
                                      void *stream) { return MELD_B200_ERR_UNSUPPORTED; }
}
