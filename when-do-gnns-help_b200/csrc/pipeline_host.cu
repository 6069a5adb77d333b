// End-to-end entry with HOST buffers: what a reference user calls when the graph lives in host
// memory (homophily_tests.py moves everything with `.to(device)` first).  H2D copies, plan,
// normaliser, A_hat X, label statistics and the D2H result copies all run on one stream and are
// inside the caller's timed region (bench.py "e2e").
#include "common.cuh"

extern "C" int wdgh_spmm_csr(const int64_t *, const int32_t *, const float *, int64_t, const float *, int64_t,
                             int64_t, float *, int64_t, int, int, const float *, const uint8_t *, const int64_t *,
                             const int64_t *, float *, int64_t, void *);

namespace wdgh {

struct HostPipelineCache {
  int64_t n = -1, nnz = -1, d = -1, cap = -1, partial_elems = 0;
  int C = -1;
  int64_t *rowptr = nullptr;
  int32_t *col = nullptr;
  float *x = nullptr, *y = nullptr, *dinv = nullptr, *partial = nullptr;
  int32_t *labels = nullptr, *deg = nullptr, *match = nullptr;
  uint8_t *labels8 = nullptr, *deg_code = nullptr;
  int64_t *plan = nullptr, *counters = nullptr;
  double *node_sum = nullptr;
  cudaStream_t st = nullptr;
  void release() {
    cudaFree(rowptr); cudaFree(col); cudaFree(x); cudaFree(y); cudaFree(dinv); cudaFree(partial);
    cudaFree(labels); cudaFree(labels8); cudaFree(deg_code); cudaFree(deg); cudaFree(match); cudaFree(plan); cudaFree(counters); cudaFree(node_sum);
    if (st) cudaStreamDestroy(st);
    *this = HostPipelineCache();
  }
};
static HostPipelineCache g_cache;
constexpr int64_t kHostPipelineThreshold = 512;

}  // namespace wdgh

using namespace wdgh;

extern "C" int wdgh_pipeline_host_release(void) {
  g_cache.release();
  return 0;
}

extern "C" int wdgh_pipeline_host(const int64_t *rowptr_host, const int32_t *col_host, int64_t n, int64_t nnz,
                                  const float *x_host, int64_t d, const int32_t *labels_host, int32_t num_classes,
                                  int norm, int add_self_loop, float *y_host, int64_t *counters_host,
                                  double *node_sum_host) {
  WDGH_REQUIRE(rowptr_host && x_host && labels_host && counters_host && node_sum_host, "wdgh_pipeline_host: null pointer");
  WDGH_REQUIRE(n > 0 && nnz >= 0 && d > 0 && num_classes >= 1 && (col_host || nnz == 0), "wdgh_pipeline_host: bad shape");
  HostPipelineCache &c = g_cache;
  const int C = num_classes;
  const size_t n_counters = WDGH_SC_WORDS((size_t)C);
  const int64_t cap = 2 * nnz / kHostPipelineThreshold + 2;
  if (c.n != n || c.nnz != nnz || c.d != d || c.C != C) {
    c.release();
    WDGH_CUDA(cudaStreamCreateWithFlags(&c.st, cudaStreamNonBlocking));
    WDGH_CUDA(cudaMalloc(&c.rowptr, (n + 1) * sizeof(int64_t)));
    WDGH_CUDA(cudaMalloc(&c.col, (nnz > 0 ? nnz : 1) * sizeof(int32_t)));
    WDGH_CUDA(cudaMalloc(&c.x, n * d * sizeof(float)));
    WDGH_CUDA(cudaMalloc(&c.y, n * d * sizeof(float)));
    WDGH_CUDA(cudaMalloc(&c.dinv, n * sizeof(float)));
    WDGH_CUDA(cudaMalloc(&c.labels, n * sizeof(int32_t)));
    WDGH_CUDA(cudaMalloc(&c.labels8, n));
    WDGH_CUDA(cudaMalloc(&c.deg_code, n));
    WDGH_CUDA(cudaMalloc(&c.deg, n * sizeof(int32_t)));
    WDGH_CUDA(cudaMalloc(&c.match, n * sizeof(int32_t)));
    WDGH_CUDA(cudaMalloc(&c.plan, WDGH_PLAN_WORDS(cap, nnz) * sizeof(int64_t)));
    WDGH_CUDA(cudaMalloc(&c.counters, n_counters * sizeof(int64_t)));
    WDGH_CUDA(cudaMalloc(&c.node_sum, 2 * sizeof(double)));
    c.n = n; c.nnz = nnz; c.d = d; c.C = C; c.cap = cap;
  }
  cudaStream_t st = c.st;
  WDGH_CUDA(cudaMemcpyAsync(c.rowptr, rowptr_host, (n + 1) * sizeof(int64_t), cudaMemcpyHostToDevice, st));
  if (nnz) WDGH_CUDA(cudaMemcpyAsync(c.col, col_host, nnz * sizeof(int32_t), cudaMemcpyHostToDevice, st));
  WDGH_CUDA(cudaMemcpyAsync(c.labels, labels_host, n * sizeof(int32_t), cudaMemcpyHostToDevice, st));
  WDGH_CUDA(cudaMemcpyAsync(c.x, x_host, n * d * sizeof(float), cudaMemcpyHostToDevice, st));
  int64_t plan_host[8];
  int rc = wdgh_plan_build(c.rowptr, n, nnz, kHostPipelineThreshold, c.plan, cap, plan_host, st);
  if (rc) return rc;
  const int64_t ldp = (d + 3) & ~int64_t(3);
  const int64_t n_part = plan_host[1] > 2 * plan_host[5] ? plan_host[1] : 2 * plan_host[5];
  if (n_part * ldp > c.partial_elems) {
    cudaFree(c.partial);
    c.partial = nullptr;
    c.partial_elems = n_part * ldp;
    WDGH_CUDA(cudaMalloc(&c.partial, c.partial_elems * sizeof(float)));
  }
  if (norm != WDGH_NORM_NONE) {
    rc = wdgh_degree_scale(c.rowptr, nullptr, n, norm, add_self_loop, c.dinv, nullptr, c.deg_code, st);
    if (rc) return rc;
  }
  rc = wdgh_spmm_structure_fused(c.rowptr, c.col, n, nnz, c.x, d, d, c.y, d, norm, add_self_loop,
                                 norm != WDGH_NORM_NONE ? c.dinv : nullptr,
                                 norm != WDGH_NORM_NONE ? c.deg_code : nullptr, c.labels, C, c.plan, plan_host,
                                 c.partial, c.counters, c.node_sum, c.deg, c.match, c.labels8, n, 0, 0, st);
  if (rc) return rc;
  WDGH_CUDA(cudaMemcpyAsync(counters_host, c.counters, n_counters * sizeof(int64_t), cudaMemcpyDeviceToHost, st));
  WDGH_CUDA(cudaMemcpyAsync(node_sum_host, c.node_sum, 2 * sizeof(double), cudaMemcpyDeviceToHost, st));
  if (y_host) WDGH_CUDA(cudaMemcpyAsync(y_host, c.y, n * d * sizeof(float), cudaMemcpyDeviceToHost, st));
  WDGH_CUDA(cudaStreamSynchronize(st));
  return 0;
}
