// End-to-end entry with HOST buffers: what a reference user calls when the graph lives in host
// memory (homophily_tests.py moves everything with `.to(device)` first).  H2D copies, plan,
// normaliser, A_hat X, label statistics and the D2H result copies are all inside the caller's
// timed region (bench.py "e2e").
//
// The pass is PCIe-bound (the feature matrix is 85% of the bytes), so the entry is a two-stream
// pipeline: the CSR and the labels go first; the feature matrix follows in 16 row blocks on
// a copy stream while the compute stream builds the plan, runs the label pass and then aggregates,
// per arriving block, the entries whose source node lies in that block (wdgh_spmm_csr_ranged:
// column ranges of every row, y +=, self loop + scale with the last block).  Only the last
// block's phase is exposed after the copies end.
//
// When Y = A_hat X has to come back to the host as well (y_host != NULL) the D2H copy of Y would follow the whole
// H2D stream -- Y depends on every feature ROW block -- and the step would cost H2D + D2H back to back (1070 ms for
// 30.1 + 25.6 GB).  PCIe is full duplex, and a COLUMN block of Y needs only the same column block of X, so this form
// splits the feature matrix by columns instead: pitched copies (cudaMemcpy2DAsync, 256-byte rows out of the 512-byte
// pitch keep the full link rate: 55.6 / 53.5 GB/s alone, 48 + 50 GB/s both directions at once, tools/pcie_probe.cu),
// aggregation of a column block on strided views (ldx = ldy = d; bit-identical to the one-launch result,
// tools/colblock_probe.py) phased by arriving row chunk exactly like the counters-only form, so that only the last
// chunk's phase separates the end of a block's H2D copy from the start of its D2H copy, and the D2H copy of block k
// runs under the H2D copy of block k + 1.
#include <stdlib.h>

#include <mutex>

#include "common.cuh"

namespace wdgh {

struct HostPipelineCache {
  int64_t n = -1, nnz = -1, d = -1, cap = -1, partial_elems = 0;
  int C = -1, device = -1;
  int64_t *rowptr = nullptr;
  int32_t *col = nullptr;
  float *x = nullptr, *y = nullptr, *dinv = nullptr, *partial = nullptr;
  int32_t *labels = nullptr, *deg = nullptr, *match = nullptr;
  uint8_t *labels8 = nullptr, *deg_code = nullptr;
  int64_t *plan = nullptr, *counters = nullptr;
  double *node_sum = nullptr;
  int64_t *seg = nullptr, *bounds = nullptr;  // column segments of the feature row blocks
  uint8_t *skip = nullptr;
  cudaStream_t st = nullptr, st_copy = nullptr, st_back = nullptr;
  cudaEvent_t ev_csr = nullptr, ev_x[64] = {nullptr}, ev_y[64] = {nullptr};
  void release() {
    cudaFree(seg); cudaFree(bounds); cudaFree(skip);
    if (st_copy) cudaStreamDestroy(st_copy);
    if (st_back) cudaStreamDestroy(st_back);
    if (ev_csr) cudaEventDestroy(ev_csr);
    for (auto &e : ev_x) if (e) cudaEventDestroy(e);
    for (auto &e : ev_y) if (e) cudaEventDestroy(e);
    cudaFree(rowptr); cudaFree(col); cudaFree(x); cudaFree(y); cudaFree(dinv); cudaFree(partial);
    cudaFree(labels); cudaFree(labels8); cudaFree(deg_code); cudaFree(deg); cudaFree(match); cudaFree(plan); cudaFree(counters); cudaFree(node_sum);
    if (st) cudaStreamDestroy(st);
    *this = HostPipelineCache();
  }
};
static HostPipelineCache g_cache;  // one cached pipeline per process, guarded by g_cache_mu
static std::mutex g_cache_mu;
constexpr int64_t kHostPipelineThreshold = 512;

// feature row blocks of the pipelined copy (1 = copy everything, then compute); WDGH_E2E_CHUNKS overrides
static int e2e_chunks() {
  static int cached = 0;
  if (cached == 0) {
    int v = 16;  // measured (30.1 GB per pass): 1 -> 635 ms, 4 -> 587, 8 -> 579, 16 -> 574 ms; the copies alone take ~548 ms
    if (const char *e = getenv("WDGH_E2E_CHUNKS")) v = atoi(e);
    cached = (v >= 1 && v <= 64) ? v : 16;
  }
  return cached;
}

// Width (floats) of the column blocks of the Y-returning form, 0 = row-block form.  64 floats = 256-byte rows: the
// narrowest pitched copy that still runs at the full PCIe rate in both directions (128-byte rows: 34 + 40 GB/s), and a
// width the row-group aggregation kernel runs natively.  WDGH_E2E_COLW overrides (0, 32 or 64).
static int64_t e2e_col_width(int64_t d) {
  static int cached = -1;
  if (cached < 0) {
    int v = 64;
    if (const char *e = getenv("WDGH_E2E_COLW")) v = atoi(e);
    cached = (v == 0 || v == 32 || v == 64) ? v : 64;
  }
  return (cached > 0 && d % cached == 0 && d / cached >= 2 && d / cached <= 64) ? cached : 0;
}

// Row chunks per column block of the Y-returning form: the block's aggregation is phased by arriving row chunk (as in
// the counters-only form), so that only the last chunk's phase -- not a whole-graph launch -- stands between the end of
// a block's H2D copy and the start of its D2H copy.  A phase over 50M rows costs 20-23 ms at 64 columns whatever the
// range length (tools/phase_probe.py; 52.8 ms for the whole block in one launch), so the phases of a block must fit under
// its 230 ms H2D copy.  Measured per step: 858.9 ms (1 chunk), 831.7 ms (4), 828.9 ms (8); 4 keeps the larger margin.
// WDGH_E2E_COLROWS overrides (1 = one launch per column block).
static int e2e_col_rows() {
  static int cached = 0;
  if (cached == 0) {
    int v = 4;
    if (const char *e = getenv("WDGH_E2E_COLROWS")) v = atoi(e);
    cached = (v >= 1 && v <= 32) ? v : 4;
  }
  return cached;
}

}  // namespace wdgh

using namespace wdgh;

extern "C" int wdgh_pipeline_host_release(void) {
  std::lock_guard<std::mutex> lock(g_cache_mu);
  g_cache.release();
  return 0;
}

extern "C" int wdgh_pipeline_host(const int64_t *rowptr_host, const int32_t *col_host, int64_t n, int64_t nnz,
                                  const float *x_host, int64_t d, const int32_t *labels_host, int32_t num_classes,
                                  int norm, int add_self_loop, float *y_host, int64_t *counters_host,
                                  double *node_sum_host) {
  WDGH_REQUIRE(rowptr_host && x_host && labels_host && counters_host && node_sum_host, "wdgh_pipeline_host: null pointer");
  WDGH_REQUIRE(n > 0 && nnz >= 0 && d > 0 && num_classes >= 1 && (col_host || nnz == 0), "wdgh_pipeline_host: bad shape");
  std::lock_guard<std::mutex> lock(g_cache_mu);  // the cached buffers and streams serve one call at a time
  HostPipelineCache &c = g_cache;
  const int C = num_classes;
  int device = -1;
  WDGH_CUDA(cudaGetDevice(&device));
  const size_t n_counters = WDGH_SC_WORDS((size_t)C);
  const int64_t cap = 2 * nnz / kHostPipelineThreshold + 2;
  if (c.n != n || c.nnz != nnz || c.d != d || c.C != C || c.device != device) {
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
    WDGH_CUDA(cudaMalloc(&c.plan, WDGH_PLAN_WORDS(cap) * sizeof(int64_t)));
    WDGH_CUDA(cudaMalloc(&c.counters, n_counters * sizeof(int64_t)));
    WDGH_CUDA(cudaMalloc(&c.node_sum, 2 * sizeof(double)));
    WDGH_CUDA(cudaStreamCreateWithFlags(&c.st_copy, cudaStreamNonBlocking));
    WDGH_CUDA(cudaStreamCreateWithFlags(&c.st_back, cudaStreamNonBlocking));
    WDGH_CUDA(cudaEventCreateWithFlags(&c.ev_csr, cudaEventDisableTiming));
    for (int k = 0; k < 64; ++k) WDGH_CUDA(cudaEventCreateWithFlags(&c.ev_x[k], cudaEventDisableTiming));
    for (int k = 0; k < 64; ++k) WDGH_CUDA(cudaEventCreateWithFlags(&c.ev_y[k], cudaEventDisableTiming));
    WDGH_CUDA(cudaMalloc(&c.seg, (size_t)(e2e_chunks() + 1) * (size_t)n * sizeof(int64_t)));
    WDGH_CUDA(cudaMalloc(&c.bounds, 65 * sizeof(int64_t)));
    WDGH_CUDA(cudaMalloc(&c.skip, (size_t)n));
    c.n = n; c.nnz = nnz; c.d = d; c.C = C; c.cap = cap; c.device = device;
  }
  cudaStream_t st = c.st, sc = c.st_copy, sb = c.st_back;
  const bool ranged_ok = (d % 4 == 0) && (d >= 128 || d == 64 || d == 32);
  const int64_t colw = y_host ? e2e_col_width(d) : 0;   // > 0: column-block form (Y travels back under the X copies)
  const int KC = colw ? (int)(d / colw) : 0;
  int R = colw ? e2e_col_rows() : 1;                    // row chunks per column block
  while (R > 1 && (KC * R > 64 || R > e2e_chunks())) R /= 2;
  const int K = colw ? R : (ranged_ok ? e2e_chunks() : 1);   // row blocks whose column segments are needed
  // 1. graph + labels first (the compute stream needs them at once) ...
  WDGH_CUDA(cudaMemcpyAsync(c.rowptr, rowptr_host, (n + 1) * sizeof(int64_t), cudaMemcpyHostToDevice, sc));
  if (nnz) WDGH_CUDA(cudaMemcpyAsync(c.col, col_host, nnz * sizeof(int32_t), cudaMemcpyHostToDevice, sc));
  WDGH_CUDA(cudaMemcpyAsync(c.labels, labels_host, n * sizeof(int32_t), cudaMemcpyHostToDevice, sc));
  WDGH_CUDA(cudaEventRecord(c.ev_csr, sc));
  // 2. ... then the feature matrix on the copy stream: column blocks (pitched, each in R row chunks) when Y goes back
  //    to the host, row blocks otherwise
  const int64_t blk = (n + K - 1) / K;
  int64_t bounds_h[65];
  for (int k = 0; k < K; ++k) bounds_h[k] = (int64_t)k * blk < n ? (int64_t)k * blk : n;
  for (int kc = 0; kc < KC; ++kc) {
    for (int r = 0; r < R; ++r) {
      const int64_t r0 = bounds_h[r], r1 = (r0 + blk < n) ? r0 + blk : n;
      if (r1 > r0)
        WDGH_CUDA(cudaMemcpy2DAsync(c.x + r0 * d + kc * colw, d * sizeof(float), x_host + r0 * d + kc * colw,
                                    d * sizeof(float), colw * sizeof(float), r1 - r0, cudaMemcpyHostToDevice, sc));
      WDGH_CUDA(cudaEventRecord(c.ev_x[kc * R + r], sc));
    }
  }
  for (int k = 0; k < K && !colw; ++k) {
    const int64_t r0 = bounds_h[k], r1 = (r0 + blk < n) ? r0 + blk : n;
    if (r1 > r0)
      WDGH_CUDA(cudaMemcpyAsync(c.x + r0 * d, x_host + r0 * d, (size_t)(r1 - r0) * d * sizeof(float),
                                cudaMemcpyHostToDevice, sc));
    WDGH_CUDA(cudaEventRecord(c.ev_x[k], sc));
  }
  bounds_h[K] = n;
  WDGH_CUDA(cudaStreamWaitEvent(st, c.ev_csr, 0));
  int64_t plan_host[8];
  int rc = wdgh_plan_build(c.rowptr, n, nnz, kHostPipelineThreshold, c.plan, cap, plan_host, st);
  if (rc) return rc;
  const int64_t ldp = (d + 3) & ~int64_t(3);
  const int64_t n_part = plan_host[1];
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
  const float *dinv = norm != WDGH_NORM_NONE ? c.dinv : nullptr;
  const uint8_t *code = norm != WDGH_NORM_NONE ? c.deg_code : nullptr;
  // 3. label pass: needs only the graph and the labels, runs under the feature copy
  rc = wdgh_structure_counts(c.rowptr, c.col, n, nnz, c.labels, C, c.plan, plan_host, c.counters, c.node_sum, c.deg,
                             c.match, c.labels8, n, 0, st);
  if (rc) return rc;
  if (colw) {
    // 4'. per column block: one aggregation phase per arriving row chunk (R = 1: one full-graph launch); the Y block
    //     leaves on the third stream as soon as its last phase is done
    if (R > 1) {
      WDGH_CUDA(cudaMemcpyAsync(c.bounds, bounds_h, (R + 1) * sizeof(int64_t), cudaMemcpyHostToDevice, st));
      rc = wdgh_column_segments(c.rowptr, c.col, n, c.bounds, R + 1, c.seg, st);
      if (rc) return rc;
      rc = wdgh_plan_heavy_flags(c.plan, plan_host, n, c.skip, st);
      if (rc) return rc;
    }
    for (int kc = 0; kc < KC; ++kc) {
      const float *xb = c.x + kc * colw;
      float *yb = c.y + kc * colw;
      for (int r = 0; r < R; ++r) {
        WDGH_CUDA(cudaStreamWaitEvent(st, c.ev_x[kc * R + r], 0));
        if (R == 1) {
          rc = wdgh_spmm_csr(c.rowptr, c.col, nullptr, n, xb, colw, d, yb, d, norm, add_self_loop, dinv, code, c.plan,
                             plan_host, c.partial, 0, st);
        } else {
          const int last = (r == R - 1);
          rc = wdgh_spmm_csr_ranged(c.rowptr, c.seg + (size_t)r * n, c.seg + (size_t)(r + 1) * n, c.col, nullptr, n, xb,
                                    colw, d, yb, d, norm, add_self_loop, dinv, code, c.skip, r > 0, last, last, nullptr,
                                    0, 0, 0, 0, 0, c.plan, plan_host, c.partial, 0, st);
        }
        if (rc) return rc;
      }
      WDGH_CUDA(cudaEventRecord(c.ev_y[kc], st));
      WDGH_CUDA(cudaStreamWaitEvent(sb, c.ev_y[kc], 0));
      WDGH_CUDA(cudaMemcpy2DAsync(y_host + kc * colw, d * sizeof(float), yb, d * sizeof(float), colw * sizeof(float), n,
                                  cudaMemcpyDeviceToHost, sb));
    }
  } else if (K == 1) {
    WDGH_CUDA(cudaStreamWaitEvent(st, c.ev_x[0], 0));
    rc = wdgh_spmm_csr(c.rowptr, c.col, nullptr, n, c.x, d, d, c.y, d, norm, add_self_loop, dinv, code, c.plan,
                       plan_host, c.partial, 0, st);
    if (rc) return rc;
  } else {
    // 4. one aggregation phase per arriving feature block
    WDGH_CUDA(cudaMemcpyAsync(c.bounds, bounds_h, (K + 1) * sizeof(int64_t), cudaMemcpyHostToDevice, st));
    rc = wdgh_column_segments(c.rowptr, c.col, n, c.bounds, K + 1, c.seg, st);
    if (rc) return rc;
    rc = wdgh_plan_heavy_flags(c.plan, plan_host, n, c.skip, st);
    if (rc) return rc;
    for (int k = 0; k < K; ++k) {
      WDGH_CUDA(cudaStreamWaitEvent(st, c.ev_x[k], 0));
      const int last = (k == K - 1);
      rc = wdgh_spmm_csr_ranged(c.rowptr, c.seg + (size_t)k * n, c.seg + (size_t)(k + 1) * n, c.col, nullptr, n, c.x, d,
                                d, c.y, d, norm, add_self_loop, dinv, code, c.skip, k > 0, last, last, nullptr, 0, 0, 0, 0, 0,
                                c.plan, plan_host, c.partial, 0, st);
      if (rc) return rc;
    }
  }
  // Y leaves on the copy stream as soon as the last phase is done; the counters follow on the compute stream
  if (y_host && !colw) {
    WDGH_CUDA(cudaEventRecord(c.ev_csr, st));
    WDGH_CUDA(cudaStreamWaitEvent(sc, c.ev_csr, 0));
    WDGH_CUDA(cudaMemcpyAsync(y_host, c.y, n * d * sizeof(float), cudaMemcpyDeviceToHost, sc));
  }
  WDGH_CUDA(cudaMemcpyAsync(counters_host, c.counters, n_counters * sizeof(int64_t), cudaMemcpyDeviceToHost, st));
  WDGH_CUDA(cudaMemcpyAsync(node_sum_host, c.node_sum, 2 * sizeof(double), cudaMemcpyDeviceToHost, st));
  WDGH_CUDA(cudaStreamSynchronize(st));
  if (counters_host[WDGH_SC_N_MULTI_NEG] != 0) {
    // several distinct negative labels: the 1-byte label copy folds them together, but the reference compares raw
    // labels (hm.py:51) -- redo the label pass on the int32 labels
    rc = wdgh_structure_counts(c.rowptr, c.col, n, nnz, c.labels, C, c.plan, plan_host, c.counters, c.node_sum, c.deg,
                               c.match, nullptr, n, 0, st);
    if (rc) return rc;
    WDGH_CUDA(cudaMemcpyAsync(counters_host, c.counters, n_counters * sizeof(int64_t), cudaMemcpyDeviceToHost, st));
    WDGH_CUDA(cudaMemcpyAsync(node_sum_host, c.node_sum, 2 * sizeof(double), cudaMemcpyDeviceToHost, st));
    WDGH_CUDA(cudaStreamSynchronize(st));
  }
  if (y_host) WDGH_CUDA(cudaStreamSynchronize(colw ? sb : sc));
  return 0;
}
