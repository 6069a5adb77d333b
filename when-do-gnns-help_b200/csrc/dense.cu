// Dense side of the path: row gathers, the SIMT fp32 Gram kernel (the cross-check of the tcgen05
// kernel in gram_tc.cu; both run on the GPU, the caller picks one explicitly), class-wise column sums and the
// aggregation-similarity score (utils/homophily_metrics.py:190-229), the arccos (GNTK) kernel
// transform (:232-257) and the per-edge feature cosine of generalised edge homophily (:164-187).
#include <math_constants.h>

#include "common.cuh"

namespace wdgh {

// ---------------------------------------------------------------------------
// out[k][:] = x[ids[k]][:]
// ---------------------------------------------------------------------------
__global__ void gather_rows_kernel(const float *__restrict__ x, int64_t d, int64_t ldx,
                                   const int64_t *__restrict__ ids, int64_t m, float *__restrict__ out,
                                   int64_t ldo) {
  const int lane = threadIdx.x & 31;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t k = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; k < m; k += nwarps) {
    const float *src = x + ids[k] * ldx;
    float *dst = out + k * ldo;
    for (int64_t c = lane; c < d; c += 32) dst[c] = __ldg(src + c);
  }
}

// ---------------------------------------------------------------------------
// g = z z^T, SIMT fp32: 64x64 output tile per CTA, 16x16 threads, 4x4 micro-tile, K-tile 16
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
gram_simt_kernel(const float *__restrict__ z, int64_t m, int64_t d, int64_t ldz, float *__restrict__ g,
                 int64_t ldg) {
  __shared__ float sa[16][64 + 1];
  __shared__ float sb[16][64 + 1];
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int64_t i0 = (int64_t)blockIdx.y * 64, j0 = (int64_t)blockIdx.x * 64;
  float acc[4][4] = {};
  for (int64_t k0 = 0; k0 < d; k0 += 16) {
    // 64 rows x 16 k per operand: 1024 elements, 4 per thread; consecutive threads read consecutive k
    for (int t = threadIdx.x; t < 1024; t += 256) {
      const int r = t >> 4, kk = t & 15;
      const int64_t k = k0 + kk;
      const int64_t ia = i0 + r, jb = j0 + r;
      sa[kk][r] = (ia < m && k < d) ? __ldg(z + ia * ldz + k) : 0.f;
      sb[kk][r] = (jb < m && k < d) ? __ldg(z + jb * ldz + k) : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < 16; ++kk) {
      float a[4], b[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        a[u] = sa[kk][ty * 4 + u];
        b[u] = sb[kk][tx * 4 + u];
      }
#pragma unroll
      for (int u = 0; u < 4; ++u)
#pragma unroll
        for (int v = 0; v < 4; ++v) acc[u][v] = fmaf(a[u], b[v], acc[u][v]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    const int64_t i = i0 + ty * 4 + u;
    if (i >= m) continue;
#pragma unroll
    for (int v = 0; v < 4; ++v) {
      const int64_t j = j0 + tx * 4 + v;
      if (j < m) g[i * ldg + j] = acc[u][v];
    }
  }
}

// ---------------------------------------------------------------------------
// w[i][c] = sum / mean over columns j with labels[j] == c of g[i][j]   (hm.py:201-206)
// one CTA per row; warp w owns classes w, w+nwarps, ...; fixed summation order (deterministic)
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
class_colsum_kernel(const float *__restrict__ g, int64_t m, int64_t ldg, const int32_t *__restrict__ labels, int C,
                    int is_mean, float *__restrict__ w) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
  for (int64_t i = blockIdx.x; i < m; i += gridDim.x) {
    const float *row = g + i * ldg;
    for (int c = wid; c < C; c += nw) {
      float s = 0.f;
      int cnt = 0;
      for (int64_t j = lane; j < m; j += 32) {
        if (__ldg(labels + j) == c) {
          s += __ldg(row + j);
          cnt += 1;
        }
      }
      s = warp_sum(s);
      cnt = (int)warp_sum((long long)cnt);
      if (lane == 0) w[i * C + c] = is_mean ? s / (float)cnt : s;  // empty class: 0/0 = NaN like torch.mean
    }
  }
}

// column sums of the label matrix: lsum[c] = sum_j label_rows[j][c]   (for degs_label, hm.py:210)
__global__ void label_colsum_kernel(const float *__restrict__ lab, int64_t m, int C, float *__restrict__ lsum) {
  const int c = blockIdx.x;
  __shared__ float part[32];
  float s = 0.f;
  for (int64_t j = threadIdx.x; j < m; j += blockDim.x) s += lab[j * C + c];
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x < 32) {
    float t = (threadIdx.x < (blockDim.x >> 5)) ? part[threadIdx.x] : 0.f;
    t = warp_sum(t);
    if (threadIdx.x == 0) lsum[c] = t;
  }
}

// aggregation-similarity indicator per node, counted into *count   (hm.py:207-229)
__global__ void las_score_kernel(const float *__restrict__ w, const int32_t *__restrict__ labels,
                                 const float *__restrict__ lab, const float *__restrict__ lsum, int64_t m, int C,
                                 int hard, int lp, int is_sum, unsigned long long *__restrict__ count) {
  long long hit = 0;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < m; i += stride) {
    const float *wi = w + i * C;
    const float *li = lab + i * C;
    const int y = labels[i];
    bool ok;
    if (!hard) {
      if (lp == 1) {
        float nnodes, degs;
        if (is_sum) {
          nnodes = (float)m;
          degs = 0.f;
          for (int c = 0; c < C; ++c) degs += li[c] * lsum[c];  // (label @ label.T).sum(1)
        } else {
          nnodes = (float)C;
          degs = 1.f;
        }
        float tot = 0.f;
        for (int c = 0; c < C; ++c) tot += wi[c];
        const float own = wi[y];
        float ratio = (own / degs) / ((tot - own) / (nnodes - degs));
        if (isnan(ratio)) ratio = 0.f;
        ok = ratio >= 1.f;
      } else {
        float off = 0.f, on = 0.f;
        for (int c = 0; c < C; ++c) {
          off += wi[c] - wi[c] * li[c];
          on += wi[c] * li[c];
        }
        ok = (off <= 0.f) && (on >= 0.f);
      }
    } else {
      if (lp == 1) {
        int arg = 0;
        float best = wi[0];
        bool best_nan = isnan(best);
        for (int c = 1; c < C; ++c) {  // torch.argmax: first maximum, NaN counts as maximum
          const float v = wi[c];
          if (!best_nan && (isnan(v) || v > best)) {
            best = v;
            arg = c;
            best_nan = isnan(v);
          }
        }
        ok = (arg == y);
      } else {
        float mx = -CUDART_INF_F, on = 0.f;
        for (int c = 0; c < C; ++c) {
          mx = fmaxf(mx, wi[c] - wi[c] * li[c]);
          on += wi[c] * li[c];
        }
        ok = (mx <= 0.f) && (on >= 0.f);
      }
    }
    hit += ok;
  }
  hit = warp_sum(hit);
  if ((threadIdx.x & 31) == 0 && hit) atomicAdd(count, (unsigned long long)hit);
}

// ---------------------------------------------------------------------------
// GNTK transform (hm.py:236-244): dsq = sqrt(diag g); then elementwise in place
// ---------------------------------------------------------------------------
__global__ void gntk_diag_kernel(const float *__restrict__ g, int64_t m, int64_t ldg, float *__restrict__ dsq) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < m; i += stride) dsq[i] = sqrtf(g[i * ldg + i]);
}
// The reference evaluates the arccos kernel op by op on float32 tensors (hm.py:236-244): every product, square,
// difference and sum is rounded separately.  `sqrt(norm^2 - g^2)` cancels catastrophically wherever two rows are
// (nearly) parallel -- the diagonal first of all -- so a fused multiply-add there changes K by up to ~3e-5 relative,
// and the KR metric's pinv(rcond=1e-15) turns such differences into flipped predictions.  The kernel therefore
// mirrors torch's rounding sequence with explicit round-to-nearest intrinsics (no FMA contraction), and takes acos
// in float64 rounded once to float32 (0.5 ulp; torch's vectorised float32 acos is a 1-ulp routine that agrees
// with the correctly rounded value on almost every input).
__global__ void gntk_apply_kernel(float *__restrict__ g, int64_t m, int64_t ldg, const float *__restrict__ dsq,
                                  int n_layers) {
  const float eps = 1e-8f;
  const float pi = 3.14159265358979323846f;
  const float inv_pi = (float)(1.0 / 3.14159265358979323846);
  const int64_t total = m * m;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += stride) {
    const int64_t i = t / m, j = t - i * m;
    const float v = g[i * ldg + j];
    float k;
    if (n_layers == 1) {
      const float raw = __fmul_rn(dsq[i], dsq[j]);
      // (norm > eps) * norm + eps * (norm <= eps): NaN fails both comparisons -> 0 * NaN + eps * 0 = NaN in torch
      float norm = (raw > eps) ? raw : eps;
      if (isnan(raw)) norm = CUDART_NAN_F;
      float ac = (float)acos((double)__fdiv_rn(v, norm));
      float root = __fsqrt_rn(__fsub_rn(__fmul_rn(norm, norm), __fmul_rn(v, v)));
      if (isnan(ac)) ac = 0.f;
      if (isnan(root)) root = 0.f;
      k = __fmul_rn(inv_pi, __fadd_rn(__fmul_rn(v, __fsub_rn(pi, ac)), root));
    } else {
      k = v;
    }
    g[i * ldg + j] = k / 2.f;
  }
}

// NTK branch of homophily_plot.py similarity (:191-193): k = clamp(g,0,1); g = k (pi - acos k) / (2 pi)
__global__ void ntk_clamp_kernel(float *__restrict__ g, int64_t m, int64_t ldg) {
  const float pi = 3.14159265358979323846f;
  const int64_t total = m * m;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += stride) {
    const int64_t i = t / m, j = t - i * m;
    const float k = fminf(fmaxf(g[i * ldg + j], 0.f), 1.f);
    g[i * ldg + j] = (k * (pi - acosf(k))) / (2.f * pi);
  }
}

// ---------------------------------------------------------------------------
// per-entry feature cosine (hm.py:167-172 / 181-186): one warp per entry
// ---------------------------------------------------------------------------
__device__ __forceinline__ int64_t row_of_entry(const int64_t *rowptr, int64_t n, int64_t e) {
  // largest r with rowptr[r] <= e
  int64_t lo = 0, hi = n;
  while (hi - lo > 1) {
    const int64_t mid = (lo + hi) >> 1;
    if (rowptr[mid] <= e) lo = mid;
    else hi = mid;
  }
  return lo;
}

__global__ void __launch_bounds__(256)
edge_cosine_kernel(const int64_t *__restrict__ rowptr, const int32_t *__restrict__ col,
                   const float *__restrict__ val, int64_t n, const float *__restrict__ x, int64_t d, int64_t ldx,
                   int mode, const int64_t *__restrict__ ids, int64_t n_ids, double *__restrict__ out_sum,
                   unsigned long long *__restrict__ out_cnt) {
  const int lane = threadIdx.x & 31;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  double sum = 0.0;
  long long cnt = 0;
  for (int64_t t = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; t < n_ids; t += nwarps) {
    const int64_t e = ids ? ids[t] : t;
    const int64_t i = row_of_entry(rowptr, n, e);
    const int64_t j = col[e];
    const bool raw_dot = (mode & 2) != 0;
    if ((mode & 1) == 0) {
      if (i == j) continue;                       // adj - diag(adj)   (hm.py:170)
      if (val != nullptr && !(val[e] > 0.f)) continue;  // (adj > 0)      (hm.py:171)
    }
    const float *a = x + i * ldx, *b = x + j * ldx;
    float dot = 0.f, na = 0.f, nb = 0.f;
    for (int64_t k = lane; k < d; k += 32) {
      const float u = __ldg(a + k), v = __ldg(b + k);
      dot = fmaf(u, v, dot);
      na = fmaf(u, u, na);
      nb = fmaf(v, v, nb);
    }
    dot = warp_sum(dot);
    na = warp_sum(na);
    nb = warp_sum(nb);
    float sim = raw_dot ? dot : dot / (sqrtf(na) * sqrtf(nb));
    if (isnan(sim)) sim = 0.f;
    sum += (double)sim;
    cnt += 1;
  }
  if (lane == 0) {
    if (sum != 0.0) atomicAdd(out_sum, sum);
    if (cnt) atomicAdd(out_cnt, (unsigned long long)cnt);
  }
}

}  // namespace wdgh

using namespace wdgh;

int wdgh_gram_tc_launch(const float *z, int64_t m, int64_t d, int64_t ldz, float *g, int64_t ldg, float *workspace,
                        int chunk, cudaStream_t st);

extern "C" int wdgh_gather_rows(const float *x, int64_t d, int64_t ldx, const int64_t *ids, int64_t m, float *out,
                                int64_t ldo, void *stream) {
  WDGH_REQUIRE(x && ids && out && d >= 1 && ldx >= d && ldo >= d && m >= 0, "wdgh_gather_rows: bad arguments");
  if (m == 0) return 0;
  gather_rows_kernel<<<persistent_grid(ceil_div(m, 8), 8), 256, 0, as_stream(stream)>>>(x, d, ldx, ids, m, out, ldo);
  WDGH_LAUNCHED("gather_rows_kernel");
  return 0;
}

extern "C" int wdgh_gram(const float *z, int64_t m, int64_t d, int64_t ldz, float *g, int64_t ldg,
                         int mode, float *workspace, void *stream) {
  WDGH_REQUIRE(z && g && m >= 0 && d >= 1 && ldz >= d && ldg >= m, "wdgh_gram: bad arguments");
  WDGH_REQUIRE(mode == WDGH_GRAM_SIMT || mode == WDGH_GRAM_TC || mode == WDGH_GRAM_TC_FAITHFUL, "wdgh_gram: bad mode");
  if (m == 0) return 0;
  cudaStream_t st = as_stream(stream);
  // K <= 32 (the similarity of one-hot labels, homophily_tests.py:131): a single k-block cannot feed the tensor
  // pipe and the 128 x 256 tiles are all epilogue -- the SIMT kernel is faster there (round 1: 0.22 vs 0.46 ms at
  // m = 10000, d = 10), so the tensor-core modes route it themselves.
  if (mode != WDGH_GRAM_SIMT && d > 32)
    return wdgh_gram_tc_launch(z, m, d, ldz, g, ldg, workspace, mode == WDGH_GRAM_TC_FAITHFUL ? 1 : 4, st);
  dim3 grid((unsigned)ceil_div(m, 64), (unsigned)ceil_div(m, 64));
  gram_simt_kernel<<<grid, 256, 0, st>>>(z, m, d, ldz, g, ldg);
  WDGH_LAUNCHED("gram_simt_kernel");
  return 0;
}

extern "C" int wdgh_class_colsum(const float *g, int64_t m, int64_t ldg, const int32_t *labels, int32_t num_classes,
                                 int is_mean, float *w, void *stream) {
  WDGH_REQUIRE(g && labels && w && m >= 0 && ldg >= m && num_classes >= 1, "wdgh_class_colsum: bad arguments");
  if (m == 0) return 0;
  class_colsum_kernel<<<persistent_grid(m, 8), 256, 0, as_stream(stream)>>>(g, m, ldg, labels, num_classes, is_mean,
                                                                           w);
  WDGH_LAUNCHED("class_colsum_kernel");
  return 0;
}

extern "C" int wdgh_las_score(const float *w, const int32_t *labels, const float *label_rows, int64_t m,
                              int32_t num_classes, int hard, int lp, int is_sum, float *scratch_c,
                              unsigned long long *count, void *stream) {
  WDGH_REQUIRE(w && labels && label_rows && scratch_c && count && m >= 0 && num_classes >= 1,
               "wdgh_las_score: bad arguments");
  cudaStream_t st = as_stream(stream);
  WDGH_CUDA(cudaMemsetAsync(count, 0, sizeof(unsigned long long), st));
  if (m == 0) return 0;
  label_colsum_kernel<<<(unsigned)num_classes, 256, 0, st>>>(label_rows, m, num_classes, scratch_c);
  WDGH_LAUNCHED("label_colsum_kernel");
  las_score_kernel<<<persistent_grid(ceil_div(m, 256), 4), 256, 0, st>>>(w, labels, label_rows, scratch_c, m,
                                                                        num_classes, hard, lp, is_sum, count);
  WDGH_LAUNCHED("las_score_kernel");
  return 0;
}

extern "C" int wdgh_gntk_transform(float *g, int64_t m, int64_t ldg, int n_layers, float *scratch_m, void *stream) {
  WDGH_REQUIRE(g && scratch_m && m >= 0 && ldg >= m && (n_layers == 0 || n_layers == 1),
               "wdgh_gntk_transform: bad arguments");
  if (m == 0) return 0;
  cudaStream_t st = as_stream(stream);
  gntk_diag_kernel<<<persistent_grid(ceil_div(m, 256), 4), 256, 0, st>>>(g, m, ldg, scratch_m);
  WDGH_LAUNCHED("gntk_diag_kernel");
  gntk_apply_kernel<<<persistent_grid(ceil_div(m * m, 256), 8), 256, 0, st>>>(g, m, ldg, scratch_m, n_layers);
  WDGH_LAUNCHED("gntk_apply_kernel");
  return 0;
}

extern "C" int wdgh_ntk_clamp_transform(float *g, int64_t m, int64_t ldg, void *stream) {
  WDGH_REQUIRE(g && m >= 0 && ldg >= m, "wdgh_ntk_clamp_transform: bad arguments");
  if (m == 0) return 0;
  ntk_clamp_kernel<<<persistent_grid(ceil_div(m * m, 256), 8), 256, 0, as_stream(stream)>>>(g, m, ldg);
  WDGH_LAUNCHED("ntk_clamp_kernel");
  return 0;
}

extern "C" int wdgh_edge_cosine(const int64_t *rowptr, const int32_t *col, const float *val, int64_t n,
                                const float *x, int64_t d, int64_t ldx, int mode, const int64_t *entry_ids,
                                int64_t n_ids, double *out_sum, unsigned long long *out_cnt, void *stream) {
  WDGH_REQUIRE(rowptr && x && out_sum && out_cnt && n >= 0 && d >= 1 && ldx >= d && n_ids >= 0,
               "wdgh_edge_cosine: bad arguments");
  WDGH_REQUIRE(mode >= 0 && mode <= 3 && ((mode & 1) == 0 || entry_ids != nullptr), "wdgh_edge_cosine: bad mode");
  cudaStream_t st = as_stream(stream);
  WDGH_CUDA(cudaMemsetAsync(out_sum, 0, sizeof(double), st));
  WDGH_CUDA(cudaMemsetAsync(out_cnt, 0, sizeof(unsigned long long), st));
  if (n_ids == 0 || n == 0) return 0;
  edge_cosine_kernel<<<persistent_grid(ceil_div(n_ids, 8), 8), 256, 0, st>>>(rowptr, col, val, n, x, d, ldx, mode,
                                                                            entry_ids, n_ids, out_sum, out_cnt);
  WDGH_LAUNCHED("edge_cosine_kernel");
  return 0;
}
