// A_hat X aggregation: CSR SpMM in float32 with on-the-fly D^-1/2 / D^-1 scaling.
//
// Replaces torch.spmm(adj, features) / torch.mm(adj, features) of the reference
// (utils/homophily_metrics.py:192,199,234,299,315) together with the materialised
// normalisers of utils/util_funcs.py:383-390 (row) and :418-426 (symmetric):
//     y_i = s_i * ( sum_{j in row i} w_ij * t_j * x_j  +  [self loop] t_i * x_i )
// with s = t = D^-1/2 (SYM), s = D^-1, t = 1 (RW), s = t = 1 (NONE).  A_hat itself is
// never written to memory.
//
// HBM-bound gather kernel (no tensor cores: there is no dense operand reuse):
//   * a group of G lanes owns one row; lane g holds VEC contiguous floats of NCH
//     column chunks, so one neighbour row is fetched with coalesced 16-byte loads
//     (d=128: one warp, one float4 per lane, 512 B per neighbour);
//   * the group loads G column indices (+ weights) with one coalesced load, keeps
//     them in registers ("row segment staged on chip") and broadcasts them with
//     shuffles; the inner loop issues U independent row gathers before the FMAs so
//     every warp keeps U*512 B in flight;
//   * degree-binned load balance: rows longer than the plan's threshold are skipped
//     here and split into fixed-size chunks, one warp per chunk, summed in a fixed
//     order by a second small kernel (deterministic, no float atomics).
#include "common.cuh"

namespace wdgh {

template <int VEC>
struct Vec;
template <>
struct Vec<4> {
  float4 v;
  __device__ __forceinline__ void zero() { v = make_float4(0.f, 0.f, 0.f, 0.f); }
  __device__ __forceinline__ void load(const float *p) { v = ldg_na(reinterpret_cast<const float4 *>(p)); }
  __device__ __forceinline__ void fma(float w, const Vec &o) {
    v.x = fmaf(w, o.v.x, v.x);
    v.y = fmaf(w, o.v.y, v.y);
    v.z = fmaf(w, o.v.z, v.z);
    v.w = fmaf(w, o.v.w, v.w);
  }
  __device__ __forceinline__ void add(const Vec &o) {
    v.x += o.v.x; v.y += o.v.y; v.z += o.v.z; v.w += o.v.w;
  }
  __device__ __forceinline__ void scale(float s) { v.x *= s; v.y *= s; v.z *= s; v.w *= s; }
  __device__ __forceinline__ void store_stream(float *p) const { st_cs(reinterpret_cast<float4 *>(p), v); }
  __device__ __forceinline__ void store(float *p) const { *reinterpret_cast<float4 *>(p) = v; }
};
template <>
struct Vec<1> {
  float v;
  __device__ __forceinline__ void zero() { v = 0.f; }
  __device__ __forceinline__ void load(const float *p) { v = __ldg(p); }
  __device__ __forceinline__ void fma(float w, const Vec &o) { v = fmaf(w, o.v, v); }
  __device__ __forceinline__ void add(const Vec &o) { v += o.v; }
  __device__ __forceinline__ void scale(float s) { v *= s; }
  __device__ __forceinline__ void store_stream(float *p) const { __stcs(p, v); }
  __device__ __forceinline__ void store(float *p) const { *p = v; }
};

// Accumulate entries [s, e) of one row into acc[NCH].  All G lanes of the group call this together.
template <int G, int VEC, int NCH, bool HAS_VAL>
__device__ __forceinline__ void accumulate_range(Vec<VEC> (&acc)[NCH], int64_t s, int64_t e,
                                                 const int32_t *__restrict__ col, const float *__restrict__ val,
                                                 const float *__restrict__ tscale,  // t_j or nullptr
                                                 const float *__restrict__ x, int64_t ldx, int cbase, int d,
                                                 int gl, unsigned gmask) {
  constexpr int U0 = (NCH >= 4) ? 2 : (NCH == 2 ? 4 : 8);  // independent gathers in flight per lane
  constexpr int U = U0 < G ? U0 : G;
  bool live[NCH];
#pragma unroll
  for (int t = 0; t < NCH; ++t) live[t] = cbase + (t * G + gl) * VEC < d;

  for (int64_t base = s; base < e; base += G) {
    const int64_t idx = base + gl;
    int j = 0;
    float w = 0.f;
    if (idx < e) {
      j = __ldg(col + idx);
      w = HAS_VAL ? __ldg(val + idx) : 1.f;
      if (tscale != nullptr) w *= __ldg(tscale + j);
    }
    const int cnt = (int)min((int64_t)G, e - base);
    for (int k = 0; k < cnt; k += U) {
      Vec<VEC> v[U][NCH];
      float ww[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        // slots past the end of the segment are predicated off (group-uniform condition)
        const bool on = k + u < cnt;
        const int src = on ? k + u : 0;
        const int jj = __shfl_sync(gmask, j, src, G);
        ww[u] = __shfl_sync(gmask, w, src, G);
        const float *xr = x + (int64_t)jj * ldx + cbase;
#pragma unroll
        for (int t = 0; t < NCH; ++t) {
          if (on && live[t]) v[u][t].load(xr + (t * G + gl) * VEC);
          else v[u][t].zero();
        }
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
#pragma unroll
        for (int t = 0; t < NCH; ++t) acc[t].fma(ww[u], v[u][t]);
      }
    }
  }
}

// One group of G lanes per row; rows longer than `threshold` are left to the chunk kernels.
template <int G, int VEC, int NCH, bool HAS_VAL>
__global__ void __launch_bounds__(256)
spmm_rows_kernel(const int64_t *__restrict__ rowptr, const int32_t *__restrict__ col,
                 const float *__restrict__ val, int64_t n, const float *__restrict__ x, int d, int64_t ldx,
                 float *__restrict__ y, int64_t ldy, int norm, int self_loop, const float *__restrict__ dinv,
                 int64_t threshold, int64_t row_offset) {
  constexpr int RPW = 32 / G;
  const int lane = threadIdx.x & 31;
  const int gl = lane % G;
  const int grp = lane / G;
  const unsigned gmask = (G == 32) ? 0xffffffffu : (((1u << G) - 1u) << (grp * G));
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t row = warp * RPW + grp;
  const int cbase = blockIdx.y * (G * VEC * NCH);
  if (row >= n) return;
  const int64_t s = __ldg(rowptr + row), e = __ldg(rowptr + row + 1);
  if (e - s > threshold) return;

  Vec<VEC> acc[NCH];
#pragma unroll
  for (int t = 0; t < NCH; ++t) acc[t].zero();
  accumulate_range<G, VEC, NCH, HAS_VAL>(acc, s, e, col, val, norm == WDGH_NORM_SYM ? dinv : nullptr, x, ldx,
                                         cbase, d, gl, gmask);
  const int64_t grow = row + row_offset;  // id of this row in the global (column) index space
  const float si = (norm != WDGH_NORM_NONE) ? __ldg(dinv + grow) : 1.f;
  const float self_w = (norm == WDGH_NORM_SYM) ? si : 1.f;
#pragma unroll
  for (int t = 0; t < NCH; ++t) {
    const int c = cbase + (t * G + gl) * VEC;
    if (c < d) {
      if (self_loop) {
        Vec<VEC> xi;
        xi.load(x + grow * ldx + c);
        acc[t].fma(self_w, xi);
      }
      acc[t].scale(si);
      acc[t].store_stream(y + row * ldy + c);
    }
  }
}

// One warp per chunk of a split row: partial[chunk][:] = sum over the chunk's entries (unscaled by s_i).
template <int VEC, int NCH, bool HAS_VAL>
__global__ void __launch_bounds__(256)
spmm_chunks_kernel(const int64_t *__restrict__ rowptr, const int32_t *__restrict__ col,
                   const float *__restrict__ val, const float *__restrict__ x, int d, int64_t ldx, int norm,
                   const float *__restrict__ dinv, const int64_t *__restrict__ plan, int64_t n_chunks,
                   float *__restrict__ partial, int64_t ldp) {
  const int lane = threadIdx.x & 31;
  const int64_t chunk = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int cbase = blockIdx.y * (32 * VEC * NCH);
  if (chunk >= n_chunks) return;
  const int64_t cap = plan[kPlanCapacity], T = plan[kPlanThreshold];
  const int64_t k = plan_chunk_owner(plan, cap)[chunk];
  const int64_t row = plan_heavy_row(plan)[k];
  const int64_t part = chunk - plan_heavy_chunk0(plan, cap)[k];
  const int64_t s = __ldg(rowptr + row) + part * T;
  const int64_t e = min(s + T, __ldg(rowptr + row + 1));

  Vec<VEC> acc[NCH];
#pragma unroll
  for (int t = 0; t < NCH; ++t) acc[t].zero();
  accumulate_range<32, VEC, NCH, HAS_VAL>(acc, s, e, col, val, norm == WDGH_NORM_SYM ? dinv : nullptr, x, ldx,
                                          cbase, d, lane, 0xffffffffu);
#pragma unroll
  for (int t = 0; t < NCH; ++t) {
    const int c = cbase + (t * 32 + lane) * VEC;
    if (c < d) acc[t].store(partial + chunk * ldp + c);
  }
}

// One CTA per split row: fixed-order sum of its chunk partials, self loop, s_i scaling.
__global__ void __launch_bounds__(128)
spmm_heavy_finish_kernel(const int64_t *__restrict__ rowptr, const float *__restrict__ x, int d, int64_t ldx,
                         float *__restrict__ y, int64_t ldy, int norm, int self_loop,
                         const float *__restrict__ dinv, const int64_t *__restrict__ plan,
                         const float *__restrict__ partial, int64_t ldp, int64_t row_offset) {
  const int64_t k = blockIdx.x;
  const int64_t cap = plan[kPlanCapacity], T = plan[kPlanThreshold];
  const int64_t row = plan_heavy_row(plan)[k];
  const int64_t c0 = plan_heavy_chunk0(plan, cap)[k];
  const int64_t deg = __ldg(rowptr + row + 1) - __ldg(rowptr + row);
  const int64_t nch = (deg + T - 1) / T;
  const int64_t grow = row + row_offset;
  const float si = (norm != WDGH_NORM_NONE) ? __ldg(dinv + grow) : 1.f;
  const float self_w = (norm == WDGH_NORM_SYM) ? si : 1.f;
  for (int c = threadIdx.x; c < d; c += blockDim.x) {
    float acc = 0.f;
    for (int64_t p = 0; p < nch; ++p) acc += partial[(c0 + p) * ldp + c];
    if (self_loop) acc = fmaf(self_w, __ldg(x + grow * ldx + c), acc);
    y[row * ldy + c] = acc * si;
  }
}

struct SpmmArgs {
  const int64_t *rowptr;
  const int32_t *col;
  const float *val;
  int64_t n;
  const float *x;
  int d;
  int64_t ldx;
  float *y;
  int64_t ldy;
  int norm, self_loop;
  const float *dinv;
  const int64_t *plan;
  int64_t threshold, n_heavy, n_chunks, row_offset;
  float *partial;
  int64_t ldp;
  cudaStream_t st;
};

template <int G, int VEC, int NCH, bool HAS_VAL>
static int launch_rows(const SpmmArgs &a) {
  constexpr int RPW = 32 / G;
  const int64_t rows_per_cta = (256 / 32) * RPW;
  const int tile = G * VEC * NCH;
  dim3 grid((unsigned)ceil_div(a.n, rows_per_cta), (unsigned)ceil_div(a.d, tile));
  spmm_rows_kernel<G, VEC, NCH, HAS_VAL><<<grid, 256, 0, a.st>>>(a.rowptr, a.col, a.val, a.n, a.x, a.d, a.ldx, a.y,
                                                                a.ldy, a.norm, a.self_loop, a.dinv, a.threshold,
                                                                a.row_offset);
  WDGH_LAUNCHED("spmm_rows_kernel");
  return 0;
}

template <int VEC, int NCH, bool HAS_VAL>
static int launch_heavy(const SpmmArgs &a) {
  if (a.n_chunks == 0) return 0;
  const int tile = 32 * VEC * NCH;
  dim3 grid((unsigned)ceil_div(a.n_chunks, 8), (unsigned)ceil_div(a.d, tile));
  spmm_chunks_kernel<VEC, NCH, HAS_VAL><<<grid, 256, 0, a.st>>>(a.rowptr, a.col, a.val, a.x, a.d, a.ldx, a.norm,
                                                               a.dinv, a.plan, a.n_chunks, a.partial, a.ldp);
  WDGH_LAUNCHED("spmm_chunks_kernel");
  spmm_heavy_finish_kernel<<<(unsigned)a.n_heavy, 128, 0, a.st>>>(a.rowptr, a.x, a.d, a.ldx, a.y, a.ldy, a.norm,
                                                                  a.self_loop, a.dinv, a.plan, a.partial, a.ldp,
                                                                  a.row_offset);
  WDGH_LAUNCHED("spmm_heavy_finish_kernel");
  return 0;
}

template <bool HAS_VAL>
static int dispatch(const SpmmArgs &a, bool vec4) {
  int rc;
  const int d = a.d;
  if (vec4) {
    if (d <= 4) rc = launch_rows<1, 4, 1, HAS_VAL>(a);
    else if (d <= 8) rc = launch_rows<2, 4, 1, HAS_VAL>(a);
    else if (d <= 16) rc = launch_rows<4, 4, 1, HAS_VAL>(a);
    else if (d <= 32) rc = launch_rows<8, 4, 1, HAS_VAL>(a);
    else if (d <= 64) rc = launch_rows<16, 4, 1, HAS_VAL>(a);
    else if (d <= 128) rc = launch_rows<32, 4, 1, HAS_VAL>(a);
    else if (d <= 256) rc = launch_rows<32, 4, 2, HAS_VAL>(a);
    else rc = launch_rows<32, 4, 4, HAS_VAL>(a);
    if (rc) return rc;
    if (d <= 128) return launch_heavy<4, 1, HAS_VAL>(a);
    if (d <= 256) return launch_heavy<4, 2, HAS_VAL>(a);
    return launch_heavy<4, 4, HAS_VAL>(a);
  }
  if (d <= 1) rc = launch_rows<1, 1, 1, HAS_VAL>(a);
  else if (d <= 2) rc = launch_rows<2, 1, 1, HAS_VAL>(a);
  else if (d <= 4) rc = launch_rows<4, 1, 1, HAS_VAL>(a);
  else if (d <= 8) rc = launch_rows<8, 1, 1, HAS_VAL>(a);
  else if (d <= 16) rc = launch_rows<16, 1, 1, HAS_VAL>(a);
  else if (d <= 32) rc = launch_rows<32, 1, 1, HAS_VAL>(a);
  else if (d <= 64) rc = launch_rows<32, 1, 2, HAS_VAL>(a);
  else rc = launch_rows<32, 1, 4, HAS_VAL>(a);
  if (rc) return rc;
  if (d <= 32) return launch_heavy<1, 1, HAS_VAL>(a);
  if (d <= 64) return launch_heavy<1, 2, HAS_VAL>(a);
  return launch_heavy<1, 4, HAS_VAL>(a);
}

}  // namespace wdgh

using namespace wdgh;

extern "C" int wdgh_spmm_csr(const int64_t *rowptr, const int32_t *col, const float *val, int64_t n,
                             const float *x, int64_t d, int64_t ldx, float *y, int64_t ldy, int norm,
                             int add_self_loop, const float *dinv, const int64_t *plan_i64,
                             const int64_t *plan_host, float *partial, int64_t row_offset, void *stream) {
  WDGH_REQUIRE(rowptr && x && y && plan_i64 && plan_host, "wdgh_spmm_csr: null pointer");  // col may be NULL iff nnz == 0
  WDGH_REQUIRE(n >= 0 && d > 0 && d <= (1 << 24) && ldx >= d && ldy >= d, "wdgh_spmm_csr: bad shape");
  WDGH_REQUIRE(norm == WDGH_NORM_NONE || norm == WDGH_NORM_RW || norm == WDGH_NORM_SYM, "wdgh_spmm_csr: bad norm");
  WDGH_REQUIRE(norm == WDGH_NORM_NONE || dinv != nullptr, "wdgh_spmm_csr: norm requires dinv");
  WDGH_REQUIRE(x != y, "wdgh_spmm_csr: in-place aggregation is not supported");
  WDGH_REQUIRE(row_offset >= 0, "wdgh_spmm_csr: negative row_offset");
  if (n == 0) return 0;
  SpmmArgs a;
  a.rowptr = rowptr; a.col = col; a.val = val; a.n = n; a.x = x; a.d = (int)d; a.ldx = ldx; a.y = y; a.ldy = ldy;
  a.norm = norm; a.self_loop = add_self_loop ? 1 : 0; a.dinv = dinv; a.plan = plan_i64;
  a.n_heavy = plan_host[0]; a.n_chunks = plan_host[1]; a.threshold = plan_host[2];
  a.partial = partial; a.ldp = (d + 3) & ~int64_t(3);
  a.row_offset = row_offset;
  a.st = as_stream(stream);
  WDGH_REQUIRE(a.n_chunks == 0 || partial != nullptr, "wdgh_spmm_csr: split rows need the partial buffer");
  const bool vec4 = (d % 4 == 0) && (ldx % 4 == 0) && (ldy % 4 == 0) &&
                    (reinterpret_cast<uintptr_t>(x) % 16 == 0) && (reinterpret_cast<uintptr_t>(y) % 16 == 0) &&
                    (partial == nullptr || reinterpret_cast<uintptr_t>(partial) % 16 == 0);
  return val ? dispatch<true>(a, vec4) : dispatch<false>(a, vec4);
}
