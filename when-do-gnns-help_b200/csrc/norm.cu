// Normalisers of utils/util_funcs.py: degree scalings (sys_normalized_adjacency :418-426,
// row_normalized_adjacency :383-390), materialised normalised values, and the dense
// normalize_tensor (:365-380).
#include "common.cuh"

namespace wdgh {

// dinv[i] = f32( (rowsum_i + self)^p ), computed in float64 like scipy does, inf -> 0.
__global__ void degree_scale_kernel(const int64_t *__restrict__ rowptr, const float *__restrict__ val, int64_t n,
                                    int norm, int self_loop, float *__restrict__ dinv,
                                    double *__restrict__ dinv64, uint8_t *__restrict__ deg_code) {
  if (val == nullptr) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
      double rs = (double)(rowptr[i + 1] - rowptr[i]) + (self_loop ? 1.0 : 0.0);
      double r;
      if (norm == WDGH_NORM_SYM) {
        if (rs == 0.0) rs = 1.0;  // util_funcs.py:422
        r = 1.0 / sqrt(rs);
      } else if (norm == WDGH_NORM_SYM_RAW) {
        r = (rs == 0.0) ? 0.0 : 1.0 / sqrt(rs);  // util_funcs.py:433-434: power(0, -0.5) = inf -> 0
      } else {
        r = (rs == 0.0) ? 0.0 : 1.0 / rs;  // sk_normalize leaves all-zero rows untouched; :31-32 inf -> 0
      }
      if (dinv) dinv[i] = (float)r;
      if (dinv64) dinv64[i] = r;
      if (deg_code) {  // stored entries per row, saturated: 255 = "look the scale up in dinv"
        const int64_t len = rowptr[i + 1] - rowptr[i];
        deg_code[i] = (uint8_t)(len < 255 ? len : 255);
      }
    }
    return;
  }
  // weighted: one warp per row sums the stored values in float64
  const int lane = threadIdx.x & 31;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; i < n; i += nwarps) {
    const int64_t s = rowptr[i], e = rowptr[i + 1];
    double rs = 0.0, ra = 0.0;
    for (int64_t q = s + lane; q < e; q += 32) {
      const double v = (double)val[q];
      rs += v;
      ra += fabs(v);
    }
    rs = warp_sum(rs);
    ra = warp_sum(ra);
    if (lane == 0) {
      double r;
      if (norm == WDGH_NORM_SYM) {
        rs += self_loop ? 1.0 : 0.0;
        if (rs == 0.0) rs = 1.0;
        r = pow(rs, -0.5);
        if (isinf(r)) r = 0.0;  // util_funcs.py:424
      } else if (norm == WDGH_NORM_SYM_RAW) {
        rs += self_loop ? 1.0 : 0.0;
        r = pow(rs, -0.5);  // util_funcs.py:433 (NaN for a negative row sum, like numpy)
        if (isinf(r)) r = 0.0;  // :434
      } else if (norm == WDGH_NORM_RW_SUM) {
        rs += self_loop ? 1.0 : 0.0;
        r = 1.0 / rs;  // util_funcs.py:31 / :41, signed row sum
        if (isinf(r)) r = 0.0;  // :32 / :42
      } else {
        ra += self_loop ? 1.0 : 0.0;  // l1 norm (sklearn normalize, util_funcs.py:386)
        r = (ra == 0.0) ? 0.0 : 1.0 / ra;
      }
      if (dinv) dinv[i] = (float)r;
      if (dinv64) dinv64[i] = r;
    }
  }
}

// Values of the normalised matrix for a CSR that already holds its diagonal; float64 products cast to
// float32 exactly like d_mat_inv_sqrt.dot(adj).dot(d_mat_inv_sqrt) -> astype(float32) (util_funcs.py:402,426).
__global__ void scale_values_kernel(const int64_t *__restrict__ rowptr, const int32_t *__restrict__ col,
                                    const float *__restrict__ val, int64_t n, int norm,
                                    const double *__restrict__ dinv64, float *__restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; i < n; i += nwarps) {
    const int64_t s = rowptr[i], e = rowptr[i + 1];
    const double di = dinv64[i];
    for (int64_t q = s + lane; q < e; q += 32) {
      const double a = val ? (double)val[q] : 1.0;
      double v = di * a;
      if (norm == WDGH_NORM_SYM || norm == WDGH_NORM_SYM_RAW) v = v * dinv64[col[q]];
      out[q] = (float)v;
    }
  }
}

// r[i] = rowsum_i^p with inf -> 0 (float32 arithmetic, torch.pow semantics)
__global__ void dense_rowscale_kernel(const float *__restrict__ x, int64_t n, int64_t d, int64_t ld, int symmetric,
                                      float *__restrict__ r) {
  const int lane = threadIdx.x & 31;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; i < n; i += nwarps) {
    float s = 0.f;
    for (int64_t k = lane; k < d; k += 32) s += x[i * ld + k];
    s = warp_sum(s);
    if (lane == 0) {
      float v = symmetric ? (1.0f / sqrtf(s)) : (1.0f / s);
      if (isinf(v)) v = 0.f;  // util_funcs.py:370,377
      r[i] = v;
    }
  }
}
__global__ void dense_apply_scale_kernel(const float *__restrict__ x, int64_t n, int64_t d, int64_t ld,
                                         int symmetric, const float *__restrict__ r, float *__restrict__ out,
                                         int64_t ldo) {
  const int64_t total = n * d;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += stride) {
    const int64_t i = t / d, k = t - i * d;
    float v = r[i] * x[i * ld + k];
    if (symmetric) v *= r[k];
    out[i * ldo + k] = v;
  }
}

}  // namespace wdgh

using namespace wdgh;

extern "C" int wdgh_degree_scale(const int64_t *rowptr, const float *val, int64_t n, int norm, int add_self_loop,
                                 float *dinv, double *dinv64, uint8_t *deg_code, void *stream) {
  WDGH_REQUIRE(rowptr && (dinv || dinv64) && n >= 0, "wdgh_degree_scale: bad arguments");
  WDGH_REQUIRE(norm >= WDGH_NORM_RW && norm <= WDGH_NORM_SYM_RAW, "wdgh_degree_scale: norm must be RW, SYM, RW_SUM or SYM_RAW");
  WDGH_REQUIRE(deg_code == nullptr || val == nullptr, "wdgh_degree_scale: degree codes need a binary adjacency");
  if (n == 0) return 0;
  const int64_t ctas = val ? ceil_div(n, 8) : ceil_div(n, 256);
  degree_scale_kernel<<<persistent_grid(ctas, 8), 256, 0, as_stream(stream)>>>(rowptr, val, n, norm,
                                                                             add_self_loop ? 1 : 0, dinv, dinv64,
                                                                             deg_code);
  WDGH_LAUNCHED("degree_scale_kernel");
  return 0;
}

extern "C" int wdgh_scale_values(const int64_t *rowptr, const int32_t *col, const float *val, int64_t n, int norm,
                                 const double *dinv64, float *out, void *stream) {
  WDGH_REQUIRE(rowptr && col && out && dinv64 && n >= 0, "wdgh_scale_values: bad arguments");
  WDGH_REQUIRE(norm >= WDGH_NORM_RW && norm <= WDGH_NORM_SYM_RAW, "wdgh_scale_values: norm must be RW, SYM, RW_SUM or SYM_RAW");
  if (n == 0) return 0;
  scale_values_kernel<<<persistent_grid(ceil_div(n, 8), 8), 256, 0, as_stream(stream)>>>(rowptr, col, val, n, norm,
                                                                                       dinv64, out);
  WDGH_LAUNCHED("scale_values_kernel");
  return 0;
}

extern "C" int wdgh_normalize_dense(const float *x, int64_t n, int64_t d, int64_t ld, int symmetric, float *scratch_n,
                                    float *out, int64_t ldo, void *stream) {
  WDGH_REQUIRE(x && out && scratch_n && n >= 0 && d >= 1 && ld >= d && ldo >= d, "wdgh_normalize_dense: bad arguments");
  WDGH_REQUIRE(!symmetric || n == d, "wdgh_normalize_dense: symmetric needs a square matrix");
  if (n == 0) return 0;
  cudaStream_t st = as_stream(stream);
  dense_rowscale_kernel<<<persistent_grid(ceil_div(n, 8), 8), 256, 0, st>>>(x, n, d, ld, symmetric, scratch_n);
  WDGH_LAUNCHED("dense_rowscale_kernel");
  dense_apply_scale_kernel<<<persistent_grid(ceil_div(n * d, 256), 8), 256, 0, st>>>(x, n, d, ld, symmetric,
                                                                                    scratch_n, out, ldo);
  WDGH_LAUNCHED("dense_apply_scale_kernel");
  return 0;
}
