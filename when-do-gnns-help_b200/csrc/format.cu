// Library core + data-format kernels: torch COO indices -> CSR, label packing, the
// degree-binned load-balance plan, prefix sums and (A + I) construction.
//
// Reference behaviour restated: `A.coalesce().indices()` (utils/homophily_metrics.py:50,63,127)
// yields row-major sorted (row, col) pairs; `adj + sp.eye(n)` (utils/util_funcs.py:385,420)
// merges or inserts the diagonal.
#include <limits.h>

#include <atomic>

#include "common.cuh"

namespace wdgh {

thread_local char g_err[512] = "";
std::atomic<uint64_t> g_launches{0};

int sm_count() {
  static int cached[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  if (cached[dev] == 0) {
    int v = 0;
    if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || v <= 0) v = 148;
    cached[dev] = v;
  }
  return cached[dev];
}

// ---------------------------------------------------------------------------
// COO (sorted) -> CSR
// ---------------------------------------------------------------------------
__global__ void coo_to_csr_kernel(const int64_t *__restrict__ indices, int64_t nnz, int64_t n,
                                  int64_t *__restrict__ rowptr, int32_t *__restrict__ col) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < nnz; e += stride) {
    const int64_t r = indices[e];
    col[e] = (int32_t)indices[nnz + e];
    const int64_t prev = (e == 0) ? -1 : indices[e - 1];
    for (int64_t q = prev + 1; q <= r; ++q) rowptr[q] = e;  // first entry of every row in (prev, r]
    if (e == nnz - 1)
      for (int64_t q = r + 1; q <= n; ++q) rowptr[q] = nnz;
  }
}

__global__ void fill_i64_kernel(int64_t *p, int64_t count, int64_t v) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += stride) p[i] = v;
}

__global__ void csr_to_coo_rows_kernel(const int64_t *__restrict__ rowptr, int64_t n, int64_t *__restrict__ row) {
  // one warp per row, grid-stride
  const int lane = threadIdx.x & 31;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t r = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; r < n; r += nwarps) {
    const int64_t s = rowptr[r], e = rowptr[r + 1];
    for (int64_t i = s + lane; i < e; i += 32) row[i] = r;
  }
}

// ---------------------------------------------------------------------------
// labels
// ---------------------------------------------------------------------------
__global__ void pack_labels_kernel(const int64_t *__restrict__ labels, int64_t n, int32_t *__restrict__ out,
                                   int32_t *__restrict__ max_label) {
  int m = INT_MIN;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const int32_t v = (int32_t)labels[i];
    out[i] = v;
    m = max(m, v);
  }
  m = warp_max(m);
  if ((threadIdx.x & 31) == 0 && m != INT_MIN) atomicMax(max_label, m);
}

__global__ void argmax_rows_kernel(const float *__restrict__ m, int64_t n, int64_t c, int64_t ld,
                                   int32_t *__restrict__ out) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const float *r = m + i * ld;
    float best = r[0];
    int32_t arg = 0;
    for (int64_t k = 1; k < c; ++k) {
      const float v = r[k];
      if (v > best) {  // first maximum wins, as torch.argmax on CPU
        best = v;
        arg = (int32_t)k;
      }
    }
    out[i] = arg;
  }
}

// ---------------------------------------------------------------------------
// load-balance plan
// ---------------------------------------------------------------------------
__global__ void plan_find_heavy_kernel(const int64_t *__restrict__ rowptr, int64_t n, int64_t T, int64_t cap,
                                       int64_t *__restrict__ plan) {
  unsigned long long *hdr = reinterpret_cast<unsigned long long *>(plan);
  int64_t *heavy_row = plan + WDGH_PLAN_HEADER;
  int64_t *heavy_chunk0 = plan + WDGH_PLAN_HEADER + cap;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const int64_t deg = rowptr[i + 1] - rowptr[i];
    if (deg > T) {
      const int64_t nch = (deg + T - 1) / T;
      const int64_t k = (int64_t)atomicAdd(&hdr[kPlanNHeavy], 1ull);
      const int64_t c0 = (int64_t)atomicAdd(&hdr[kPlanNChunks], (unsigned long long)nch);
      if (k < cap && c0 + nch <= cap) {
        heavy_row[k] = i;
        heavy_chunk0[k] = c0;
      }
    }
  }
}

__global__ void plan_fill_owner_kernel(const int64_t *__restrict__ rowptr, int64_t T, int64_t cap,
                                       int64_t *__restrict__ plan) {
  const int64_t n_heavy = plan[kPlanNHeavy];
  const int64_t *heavy_row = plan + WDGH_PLAN_HEADER;
  const int64_t *heavy_chunk0 = plan + WDGH_PLAN_HEADER + cap;
  int64_t *owner = plan + WDGH_PLAN_HEADER + 2 * cap;
  if (n_heavy > cap || plan[kPlanNChunks] > cap) return;  // reported as WDGH_ESTATE by the host
  // one warp per split row
  const int lane = threadIdx.x & 31;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t k = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; k < n_heavy; k += nwarps) {
    const int64_t r = heavy_row[k];
    const int64_t nch = (rowptr[r + 1] - rowptr[r] + T - 1) / T;
    const int64_t c0 = heavy_chunk0[k];
    for (int64_t p = lane; p < nch; p += 32) owner[c0 + p] = k;
  }
}

// ---------------------------------------------------------------------------
// exclusive prefix sum over int64 (three small kernels; n up to 2^31)
// ---------------------------------------------------------------------------
constexpr int kScanBlock = 1024;

__device__ __forceinline__ int64_t block_exclusive_scan(int64_t v, int64_t *total) {
  __shared__ int64_t warp_tot[32];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  int64_t inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int64_t t = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += t;
  }
  if (lane == 31) warp_tot[w] = inc;
  __syncthreads();
  if (w == 0) {
    int64_t t = (lane < (blockDim.x >> 5)) ? warp_tot[lane] : 0;
    int64_t ti = t;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int64_t u = __shfl_up_sync(0xffffffffu, ti, o);
      if (lane >= o) ti += u;
    }
    warp_tot[lane] = ti - t;  // exclusive warp offsets
    if (lane == 31) *total = ti;
  }
  __syncthreads();
  return warp_tot[w] + inc - v;
}

// in/out: data[i] <- exclusive prefix within its 1024-block; block_sum[b] = block total
__global__ void scan_blocks_kernel(int64_t *__restrict__ data, int64_t n, int64_t *__restrict__ block_sum) {
  __shared__ int64_t total;
  const int64_t i = (int64_t)blockIdx.x * kScanBlock + threadIdx.x;
  const int64_t v = (i < n) ? data[i] : 0;
  const int64_t ex = block_exclusive_scan(v, &total);
  if (i < n) data[i] = ex;
  if (threadIdx.x == 0) block_sum[blockIdx.x] = total;
}
// single CTA: exclusive scan of block sums (loops over 1024-wide tiles); writes the grand total to *grand
__global__ void scan_block_sums_kernel(int64_t *__restrict__ block_sum, int64_t nb, int64_t *__restrict__ grand) {
  __shared__ int64_t total;
  __shared__ int64_t carry;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (int64_t base = 0; base < nb; base += kScanBlock) {
    const int64_t i = base + threadIdx.x;
    const int64_t v = (i < nb) ? block_sum[i] : 0;
    const int64_t ex = block_exclusive_scan(v, &total);
    const int64_t c = carry;
    if (i < nb) block_sum[i] = ex + c;
    __syncthreads();
    if (threadIdx.x == 0) carry = c + total;
    __syncthreads();
  }
  if (threadIdx.x == 0) *grand = carry;
}

// ---------------------------------------------------------------------------
// A + I
// ---------------------------------------------------------------------------
__device__ __forceinline__ int64_t lower_bound_col(const int32_t *col, int64_t s, int64_t e, int32_t key) {
  while (s < e) {
    const int64_t m = (s + e) >> 1;
    if (col[m] < key) s = m + 1;
    else e = m;
  }
  return s;
}

// need[i] = 1 if row i has no stored diagonal (a new entry is inserted), else 0
__global__ void selfloop_need_kernel(const int64_t *__restrict__ rowptr, const int32_t *__restrict__ col, int64_t n,
                                     int64_t *__restrict__ need) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const int64_t s = rowptr[i], e = rowptr[i + 1];
    const int64_t p = lower_bound_col(col, s, e, (int32_t)i);
    need[i] = (p < e && col[p] == (int32_t)i) ? 0 : 1;
  }
}

// out_rowptr[i] = rowptr[i] + (#insertions before row i); then copy / merge the entries (one warp per row)
__global__ void selfloop_fill_kernel(const int64_t *__restrict__ rowptr, const int32_t *__restrict__ col,
                                     const float *__restrict__ val, int64_t n,
                                     const int64_t *__restrict__ need_ex,     // exclusive scan within 1024-blocks
                                     const int64_t *__restrict__ block_off,   // exclusive scan of block totals
                                     const int64_t *__restrict__ grand,
                                     int64_t *__restrict__ out_rowptr, int32_t *__restrict__ out_col,
                                     float *__restrict__ out_val) {
  const int lane = threadIdx.x & 31;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; i < n; i += nwarps) {
    const int64_t s = rowptr[i], e = rowptr[i + 1];
    const int64_t ins_before = need_ex[i] + block_off[i / kScanBlock];
    const int64_t os = s + ins_before;
    const int64_t p = lower_bound_col(col, s, e, (int32_t)i);  // position of / for the diagonal
    const bool has = (p < e && col[p] == (int32_t)i);
    if (lane == 0) {
      out_rowptr[i] = os;
      if (i == n - 1) out_rowptr[n] = rowptr[n] + *grand;
    }
    for (int64_t q = s + lane; q < e; q += 32) {
      const int64_t dst = os + (q - s) + ((!has && q >= p) ? 1 : 0);
      out_col[dst] = col[q];
      float v = val ? val[q] : 1.f;
      if (has && q == p) v += 1.f;
      out_val[dst] = v;
    }
    if (!has && lane == 0) {
      out_col[os + (p - s)] = (int32_t)i;
      out_val[os + (p - s)] = 1.f;
    }
  }
}

}  // namespace wdgh

using namespace wdgh;

extern "C" int wdgh_version(void) { return WDGH_VERSION; }
extern "C" const char *wdgh_last_error(void) { return g_err; }
extern "C" uint64_t wdgh_launch_count(void) { return g_launches.load(); }

extern "C" int wdgh_device_info(int *sm, int *major, int *minor) {
  int dev = 0;
  WDGH_CUDA(cudaGetDevice(&dev));
  int a = 0, b = 0, c = 0;
  WDGH_CUDA(cudaDeviceGetAttribute(&a, cudaDevAttrMultiProcessorCount, dev));
  WDGH_CUDA(cudaDeviceGetAttribute(&b, cudaDevAttrComputeCapabilityMajor, dev));
  WDGH_CUDA(cudaDeviceGetAttribute(&c, cudaDevAttrComputeCapabilityMinor, dev));
  if (sm) *sm = a;
  if (major) *major = b;
  if (minor) *minor = c;
  if (b != 10) return fail(WDGH_ENODEV, "wdgh_b200 is built for sm_100a only; this device is not compute capability 10.x");
  return 0;
}

extern "C" int wdgh_coo_to_csr(const int64_t *indices, int64_t nnz, int64_t n, int64_t *rowptr, int32_t *col,
                               void *stream) {
  WDGH_REQUIRE(rowptr && n >= 0 && nnz >= 0 && n < (int64_t)INT_MAX, "wdgh_coo_to_csr: bad arguments");
  cudaStream_t st = as_stream(stream);
  if (nnz == 0) {
    fill_i64_kernel<<<persistent_grid(ceil_div(n + 1, 256), 8), 256, 0, st>>>(rowptr, n + 1, 0);
    WDGH_LAUNCHED("fill_i64_kernel");
    return 0;
  }
  WDGH_REQUIRE(indices && col, "wdgh_coo_to_csr: null pointer");
  coo_to_csr_kernel<<<persistent_grid(ceil_div(nnz, 256), 16), 256, 0, st>>>(indices, nnz, n, rowptr, col);
  WDGH_LAUNCHED("coo_to_csr_kernel");
  return 0;
}

extern "C" int wdgh_csr_to_coo_rows(const int64_t *rowptr, int64_t n, int64_t nnz, int64_t *row, void *stream) {
  WDGH_REQUIRE(rowptr && (row || nnz == 0), "wdgh_csr_to_coo_rows: null pointer");
  if (n == 0 || nnz == 0) return 0;
  csr_to_coo_rows_kernel<<<persistent_grid(ceil_div(n, 8), 16), 256, 0, as_stream(stream)>>>(rowptr, n, row);
  WDGH_LAUNCHED("csr_to_coo_rows_kernel");
  return 0;
}

extern "C" int wdgh_pack_labels(const int64_t *labels, int64_t n, int32_t *out, int32_t *max_label, void *stream) {
  WDGH_REQUIRE(labels && out && max_label && n >= 0, "wdgh_pack_labels: bad arguments");
  cudaStream_t st = as_stream(stream);
  WDGH_CUDA(cudaMemsetAsync(max_label, 0x80, sizeof(int32_t), st));  // 0x80808080: below any label
  if (n == 0) return 0;
  pack_labels_kernel<<<persistent_grid(ceil_div(n, 256), 8), 256, 0, st>>>(labels, n, out, max_label);
  WDGH_LAUNCHED("pack_labels_kernel");
  return 0;
}

extern "C" int wdgh_argmax_rows(const float *m, int64_t n, int64_t c, int64_t ld, int32_t *out, void *stream) {
  WDGH_REQUIRE(m && out && n >= 0 && c >= 1 && ld >= c, "wdgh_argmax_rows: bad arguments");
  if (n == 0) return 0;
  argmax_rows_kernel<<<persistent_grid(ceil_div(n, 256), 8), 256, 0, as_stream(stream)>>>(m, n, c, ld, out);
  WDGH_LAUNCHED("argmax_rows_kernel");
  return 0;
}

extern "C" int wdgh_plan_build(const int64_t *rowptr, int64_t n, int64_t nnz, int64_t heavy_threshold,
                               int64_t *plan_i64, int64_t capacity, int64_t *plan_host, void *stream) {
  WDGH_REQUIRE(rowptr && plan_i64 && plan_host && n >= 0 && nnz >= 0 && heavy_threshold >= 32 && capacity >= 1,
               "wdgh_plan_build: bad arguments");
  cudaStream_t st = as_stream(stream);
  // header: counters, parameters, and the self-resetting scheduling words (tickets) of the row-group kernels
  int64_t hdr[WDGH_PLAN_HEADER] = {0, 0, heavy_threshold, capacity, 0, 0, nnz, 0, 0, 0, 0, 0, 0, 0, 0, 0};
  WDGH_CUDA(cudaMemcpyAsync(plan_i64, hdr, sizeof(hdr), cudaMemcpyHostToDevice, st));
  if (n > 0) {
    plan_find_heavy_kernel<<<persistent_grid(ceil_div(n, 256), 8), 256, 0, st>>>(rowptr, n, heavy_threshold,
                                                                                capacity, plan_i64);
    WDGH_LAUNCHED("plan_find_heavy_kernel");
    plan_fill_owner_kernel<<<persistent_grid(ceil_div(capacity, 8), 4), 256, 0, st>>>(rowptr, heavy_threshold,
                                                                                     capacity, plan_i64);
    WDGH_LAUNCHED("plan_fill_owner_kernel");
  }
  WDGH_CUDA(cudaMemcpyAsync(hdr, plan_i64, sizeof(hdr), cudaMemcpyDeviceToHost, st));
  WDGH_CUDA(cudaStreamSynchronize(st));
  plan_host[0] = hdr[kPlanNHeavy];
  plan_host[1] = hdr[kPlanNChunks];
  plan_host[2] = heavy_threshold;
  plan_host[3] = capacity;
  plan_host[4] = 0;
  plan_host[5] = 0;
  plan_host[6] = nnz;
  plan_host[7] = 0;
  if (hdr[kPlanNHeavy] > capacity || hdr[kPlanNChunks] > capacity)
    return fail(WDGH_ESTATE, "wdgh_plan_build: capacity too small (need >= 2*nnz/threshold + 2)");
  return 0;
}

extern "C" int wdgh_add_self_loops(const int64_t *rowptr, const int32_t *col, const float *val, int64_t n,
                                   int64_t *out_rowptr, int32_t *out_col, float *out_val, int64_t *scratch,
                                   void *stream) {
  WDGH_REQUIRE(rowptr && out_rowptr && out_col && out_val && scratch && n >= 0, "wdgh_add_self_loops: bad arguments");
  cudaStream_t st = as_stream(stream);
  if (n == 0) {
    WDGH_CUDA(cudaMemsetAsync(out_rowptr, 0, sizeof(int64_t), st));
    return 0;
  }
  const int64_t nb = ceil_div(n, kScanBlock);
  int64_t *need = scratch;            // [n]
  int64_t *block_sum = scratch + n;   // [nb]
  int64_t *grand = scratch + n + nb;  // [1]
  selfloop_need_kernel<<<persistent_grid(ceil_div(n, 256), 8), 256, 0, st>>>(rowptr, col, n, need);
  WDGH_LAUNCHED("selfloop_need_kernel");
  scan_blocks_kernel<<<(unsigned)nb, kScanBlock, 0, st>>>(need, n, block_sum);
  WDGH_LAUNCHED("scan_blocks_kernel");
  scan_block_sums_kernel<<<1, kScanBlock, 0, st>>>(block_sum, nb, grand);
  WDGH_LAUNCHED("scan_block_sums_kernel");
  selfloop_fill_kernel<<<persistent_grid(ceil_div(n, 8), 16), 256, 0, st>>>(rowptr, col, val, n, need, block_sum,
                                                                           grand, out_rowptr, out_col, out_val);
  WDGH_LAUNCHED("selfloop_fill_kernel");
  return 0;
}
