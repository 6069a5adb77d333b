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
#include <limits.h>
#include <stdlib.h>

#include "internal.cuh"

namespace wdgh {

template <int VEC>
struct Vec;
template <>
struct Vec<4> {
  float4 v;
  __device__ __forceinline__ void zero() { v = make_float4(0.f, 0.f, 0.f, 0.f); }
  __device__ __forceinline__ void load(const float *p) { v = ldg_na(reinterpret_cast<const float4 *>(p)); }
  __device__ __forceinline__ void load_plain(const float *p) { v = *reinterpret_cast<const float4 *>(p); }
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
struct Vec<2> {
  float2 v;
  __device__ __forceinline__ void zero() { v = make_float2(0.f, 0.f); }
  __device__ __forceinline__ void load(const float *p) {
    asm("ld.global.nc.L1::no_allocate.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "l"(p));
  }
  __device__ __forceinline__ void load_plain(const float *p) { v = *reinterpret_cast<const float2 *>(p); }
  __device__ __forceinline__ void fma(float w, const Vec &o) {
    v.x = fmaf(w, o.v.x, v.x);
    v.y = fmaf(w, o.v.y, v.y);
  }
  __device__ __forceinline__ void add(const Vec &o) { v.x += o.v.x; v.y += o.v.y; }
  __device__ __forceinline__ void scale(float s) { v.x *= s; v.y *= s; }
  __device__ __forceinline__ void store_stream(float *p) const {
    asm volatile("st.global.cs.v2.f32 [%0], {%1, %2};" ::"l"(p), "f"(v.x), "f"(v.y) : "memory");
  }
  __device__ __forceinline__ void store(float *p) const { *reinterpret_cast<float2 *>(p) = v; }
};
template <>
struct Vec<1> {
  float v;
  __device__ __forceinline__ void zero() { v = 0.f; }
  __device__ __forceinline__ void load(const float *p) { v = __ldg(p); }
  __device__ __forceinline__ void load_plain(const float *p) { v = *p; }
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
                         const float *__restrict__ partial, int64_t ldp, int64_t row_offset, int finalize) {
  const int64_t k = blockIdx.x;
  const int64_t cap = plan[kPlanCapacity], T = plan[kPlanThreshold];
  const int64_t row = plan_heavy_row(plan)[k];
  const int64_t c0 = plan_heavy_chunk0(plan, cap)[k];
  const int64_t deg = __ldg(rowptr + row + 1) - __ldg(rowptr + row);
  const int64_t nch = (deg + T - 1) / T;
  const int64_t grow = row + row_offset;
  const float si = (norm != WDGH_NORM_NONE && finalize) ? __ldg(dinv + grow) : 1.f;
  const float self_w = (norm == WDGH_NORM_SYM) ? si : 1.f;
  for (int c = threadIdx.x; c < d; c += blockDim.x) {
    float acc = 0.f;
    for (int64_t p = 0; p < nch; ++p) acc += partial[(c0 + p) * ldp + c];
    if (self_loop && finalize) acc = fmaf(self_w, __ldg(x + grow * ldx + c), acc);
    y[row * ldy + c] = acc * si;  // finalize == 0: the raw partial sum (2-D partition: reduced across ranks later)
  }
}

// ---------------------------------------------------------------------------
// Persistent, software-pipelined row kernel (d % 4 == 0, d >= 128): one warp per CTA walks rows
// row, row + W, row + 2W, ...  The dependent chain of the plain row kernel
//     rowptr -> column ids -> D^-1/2 gather -> feature rows (U at a time) -> store
// is cut to the feature-row batches: row bounds are fetched two rows ahead, the next row's column
// ids (and scales) right after the current row's first gather batch has been issued.
// For a binary adjacency the per-entry D^-1/2 gather (a second DRAM row activation per entry when
// the float array exceeds L2) is replaced by a 1-byte degree code (n bytes, L2-resident) and a
// 256-entry table in shared memory that holds bit-identical float values.
// ---------------------------------------------------------------------------
// Label statistics fused into the aggregation pass (STATS): the column ids are in registers anyway, so the
// per-entry label gather (1-byte labels, L2-resident), the match counts (warp ballots + popc) and the
// class-pair histogram (match_any fold into a shared-memory C x C table) ride along for free while the
// kernel waits on HBM -- the separate edge pass of wdgh_structure_counts disappears.
struct StatsArgs {
  const uint8_t *labels8;        // 1-byte labels by global node id, 0xFF = unlabelled
  int C;
  unsigned long long *counters;  // WDGH_SC_* layout
  int32_t *deg_nsl, *match_nsl;  // per local row
};

// Ranged / phased aggregation: the kernel normally walks CSR rows [rowptr[r], rowptr[r+1]); with row_end set it
// walks [rowptr[r], row_end[r]) instead (a column range of every row, e.g. the entries whose source node lives on
// one rank), optionally adding to what Y already holds and deferring the self-loop + D^-1/2 scaling to the last phase.
struct RangeArgs {
  const int64_t *row_end;  // nullptr: plain CSR
  const uint8_t *skip;     // nullptr, or 1 for rows that are handled elsewhere (split rows)
  int accumulate;          // y += partial instead of y = partial
  int finalize;            // apply self loop and row scale (0: store the raw partial sum)
};

template <int VEC, int NCH, bool HAS_VAL, bool FULL, int MINB, bool STATS>
__global__ void __launch_bounds__(32, MINB)
spmm_rows_pipelined_kernel(const int64_t *__restrict__ rowptr, const int32_t *__restrict__ col,
                           const float *__restrict__ val, int64_t n, const float *__restrict__ x, int d, int64_t ldx,
                           float *__restrict__ y, int64_t ldy, int norm, int self_loop,
                           const float *__restrict__ dinv, const uint8_t *__restrict__ deg_code, int64_t threshold,
                           int64_t row_offset, StatsArgs sa, RangeArgs ra) {
  extern __shared__ unsigned s_stats[];  // STATS: [C*C] class-pair histogram, then 4 scalar counters
  unsigned *s_cnt = s_stats + (STATS ? sa.C * sa.C : 0);
  if (STATS) {
    for (int b = threadIdx.x; b < sa.C * sa.C + 4; b += 32) s_stats[b] = 0;
    __syncwarp();
  }
  // One warp per row: lane l owns VEC contiguous floats of NCH chunks (VEC = 4: d >= 128; VEC = 2: d = 64;
  // VEC = 1: d = 32 -- the narrow forms serve the column-slab pipelines that overlap transfers with compute).
  // FULL: d == 32 * VEC * NCH, i.e. no column tail and a single column tile -> no per-load predicates
  constexpr int TILE = 32 * VEC * NCH;
  constexpr int U0 = 32 / (VEC * NCH);  // independent row gathers in flight per lane (16 B x 8 for d = 128)
  constexpr int U = U0 > 16 ? 16 : (U0 < 2 ? 2 : U0);
  constexpr unsigned kFull = 0xffffffffu;
  __shared__ float table[256];
  const int lane = threadIdx.x;
  const bool sym = (norm == WDGH_NORM_SYM);
  const bool coded = sym && deg_code != nullptr;
  if (coded) {
    for (int c = lane; c < 256; c += 32) {
      double rs = (double)c + (self_loop ? 1.0 : 0.0);
      if (rs == 0.0) rs = 1.0;
      table[c] = (float)(1.0 / sqrt(rs));  // same expression as degree_scale_kernel -> same bits
    }
    __syncwarp();
  }
  const int cbase = FULL ? 0 : blockIdx.y * TILE;
  bool live[NCH];
#pragma unroll
  for (int t = 0; t < NCH; ++t) live[t] = FULL || (cbase + (t * 32 + lane) * VEC < d);
  const int64_t W = gridDim.x;
  const int ld32 = (int)ldx;                       // row stride in floats (< 2^31): one IMAD.WIDE per gather
  const float *xl = x + cbase + lane * VEC;        // this lane's column slice of row 0

  auto bounds = [&](int64_t r, int64_t &s, int64_t &e) {
    s = 0;
    e = 0;
    if (r < n) {
      s = __ldg(rowptr + r);
      e = ra.row_end ? __ldg(ra.row_end + r) : __ldg(rowptr + r + 1);
      if (ra.skip && __ldg(ra.skip + r)) e = s + threshold + 1;  // marks the row as split ("heavy")
    }
  };
  // column ids + weights of the segment [pos, pos+32) of row `r`.  With STATS the label bytes of the row
  // (lir) and of each neighbour (ljr) are only REQUESTED here; seg_stats() consumes them one pipeline step
  // later, when they have long arrived, so the statistics never stall the gather stream.
  auto load_seg = [&](int64_t pos, int64_t e, int64_t r, int &j, float &w, int &lir, int &ljr) {
    j = 0;
    w = 0.f;
    const int64_t idx = pos + lane;
    const bool valid = idx < e;
    if (valid) {
      j = __ldg(col + idx);
      w = HAS_VAL ? __ldg(val + idx) : 1.f;
      if (STATS) ljr = __ldg(sa.labels8 + j);
      if (coded) {
        const int c = __ldg(deg_code + j);
        w *= (c < 255) ? table[c] : __ldg(dinv + j);
      } else if (sym) {
        w *= __ldg(dinv + j);
      }
    }
    if (STATS && pos < e) lir = __ldg(sa.labels8 + r + row_offset);
  };
  // label statistics of one segment whose (j, lir, ljr) were requested earlier; per-row counts go straight
  // to deg_nsl / match_nsl with one reduction per segment: no row state in registers
  auto seg_stats = [&](int64_t pos, int64_t e, int64_t r, int j, int lir, int ljr) {
    if (pos < e) {  // warp-uniform: the segment is not empty
      const bool valid = pos + lane < e;
      const int li = (lir == 255) ? -1 : lir;
      const int lj = (!valid || ljr == 255) ? -1 : ljr;
      const bool self = valid && ((int64_t)j == r + row_offset);
      const bool same = valid && (li == lj);
      const bool both = valid && (li >= 0) && (lj >= 0);
      const int dn = __popc(__ballot_sync(kFull, valid && !self));
      const int mn = __popc(__ballot_sync(kFull, same && !self));
      const unsigned c0 = __popc(__ballot_sync(kFull, same)), c1 = __popc(__ballot_sync(kFull, same && both));
      const unsigned c2 = __popc(__ballot_sync(kFull, both)), c3 = __popc(__ballot_sync(kFull, self));
      if (lane == 0) {
        if (dn) atomicAdd(&sa.deg_nsl[r], dn);
        if (mn) atomicAdd(&sa.match_nsl[r], mn);
        if (c0) atomicAdd(&s_cnt[0], c0);
        if (c1) atomicAdd(&s_cnt[1], c1);
        if (c2) atomicAdd(&s_cnt[2], c2);
        if (c3) atomicAdd(&s_cnt[3], c3);
      }
      fold_keys((both && !self) ? li * sa.C + lj : -1, s_stats, nullptr, true);
    }
  };

  int64_t row = blockIdx.x, r1 = row + W, r2 = row + 2 * W;
  int64_t s, e, s1, e1, s2, e2;
  bounds(row, s, e);
  bounds(r1, s1, e1);
  bool heavy = (e - s > threshold);   // split rows are handled by the chunk kernels
  if (heavy) e = s;
  bool heavy1 = (e1 - s1 > threshold);
  if (heavy1) e1 = s1;
  int j, nj = 0;
  float w, nw = 0.f;
  int lir = 255, ljr = 255, nlir = 255, nljr = 255;  // raw label bytes of the current / next row's segment
  load_seg(s, e, row, j, w, lir, ljr);

  while (row < n) {
    bounds(r2, s2, e2);  // two rows ahead
    Vec<VEC> acc[NCH];
#pragma unroll
    for (int t = 0; t < NCH; ++t) acc[t].zero();
    bool next_issued = false;
    for (int64_t base = s; base < e; base += 32) {
      if (base != s) load_seg(base, e, row, j, w, lir, ljr);
      const int cnt = (int)min((int64_t)32, e - base);
      const bool last_seg = base + 32 >= e;
      bool stats_done = !STATS;
      int k = 0;
      for (; k + U <= cnt; k += U) {  // full batches: U unpredicated gathers in flight
        Vec<VEC> v[U][NCH];
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const float *xr = xl + (int64_t)__shfl_sync(kFull, j, k + u) * ld32;
#pragma unroll
          for (int t = 0; t < NCH; ++t) {
            if (live[t]) v[u][t].load(xr + t * (32 * VEC));
            else v[u][t].zero();
          }
        }
        if (!stats_done) {  // this segment's label bytes arrived long ago; its gathers are already in flight
          seg_stats(base, e, row, j, lir, ljr);
          stats_done = true;
        }
        if (!next_issued && last_seg) {  // fetch the next row's ids while this row's gathers are in flight
          load_seg(s1, e1, r1, nj, nw, nlir, nljr);
          next_issued = true;
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const float wu = __shfl_sync(kFull, w, k + u);
#pragma unroll
          for (int t = 0; t < NCH; ++t) acc[t].fma(wu, v[u][t]);
        }
      }
      if (k < cnt) {  // tail batch, predicated
        Vec<VEC> v[U][NCH];
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const bool on = k + u < cnt;
          const float *xr = xl + (int64_t)__shfl_sync(kFull, j, on ? k + u : 0) * ld32;
#pragma unroll
          for (int t = 0; t < NCH; ++t) {
            if (on && live[t]) v[u][t].load(xr + t * (32 * VEC));
            else v[u][t].zero();
          }
        }
        if (!stats_done) {
          seg_stats(base, e, row, j, lir, ljr);
          stats_done = true;
        }
        if (!next_issued && last_seg) {
          load_seg(s1, e1, r1, nj, nw, nlir, nljr);
          next_issued = true;
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const float wu = __shfl_sync(kFull, w, (k + u < cnt) ? k + u : 0);
          if (k + u < cnt) {
#pragma unroll
            for (int t = 0; t < NCH; ++t) acc[t].fma(wu, v[u][t]);
          }
        }
      }
    }
    if (!next_issued) load_seg(s1, e1, r1, nj, nw, nlir, nljr);
    if (!heavy) {
      const int64_t grow = row + row_offset;
      const float si = (norm != WDGH_NORM_NONE && ra.finalize) ? __ldg(dinv + grow) : 1.f;
      const float self_w = sym ? si : 1.f;
#pragma unroll
      for (int t = 0; t < NCH; ++t) {
        if (live[t]) {
          const int c = cbase + (t * 32 + lane) * VEC;
          if (ra.accumulate) {  // partial sum of the earlier phases (plain load: written by a previous launch)
            Vec<VEC> prev;
            prev.load_plain(y + row * ldy + c);
            acc[t].add(prev);
          }
          if (ra.finalize) {
            if (self_loop) {
              Vec<VEC> xi;
              xi.load(x + grow * ldx + c);
              acc[t].fma(self_w, xi);
            }
            acc[t].scale(si);
            acc[t].store_stream(y + row * ldy + c);
          } else {
            acc[t].store(y + row * ldy + c);  // re-read by the next phase: keep it cacheable
          }
        }
      }
    }
    row = r1; r1 = r2; r2 += W;
    s = s1; e = e1; heavy = heavy1;
    s1 = s2; e1 = e2;
    heavy1 = (e1 - s1 > threshold);
    if (heavy1) e1 = s1;
    j = nj; w = nw;
    lir = nlir; ljr = nljr;
  }
  if (STATS) {
    __syncwarp();
    unsigned long long *g_hist = sa.counters + WDGH_SC_HEADER + 2 * sa.C;
    for (int b = lane; b < sa.C * sa.C; b += 32) {
      const unsigned v = s_stats[b];
      if (v) atomicAdd(&g_hist[b], (unsigned long long)v);
    }
    if (lane < 4 && s_cnt[lane]) atomicAdd(&sa.counters[lane], (unsigned long long)s_cnt[lane]);  // WDGH_SC_MATCH_ALL..N_SELF
  }
}

// ---------------------------------------------------------------------------
// Row-group kernel: the persistent one-warp CTA walks GROUPS of 32 consecutive rows and treats the stored
// entries of a group as one stream.  The gather batches (U feature rows in flight per lane) are cut from the
// stream, not from a row, so short rows no longer mean short batches: a row of 3 entries shares its batch
// with its neighbours, every batch but the last of a group is full and unpredicated.  Row boundaries are
// found on the consuming side (each entry carries its row slot; a change flushes the accumulator).
//   * lane l owns row 32 g + l of the group: bounds, D^-1/2 scale and skip flag are one coalesced load each;
//   * an inclusive warp scan of the row lengths gives the stream offsets (shared memory, double buffered so
//     that the next group's first segment is requested while the current group's last batch is in flight);
//   * each lane finds the row of its stream position with a 5-step search in the offsets;
//   * the self loop is a virtual LAST entry of its row (column = the row itself, weight = s_i): its feature
//     row rides in the gather batches instead of being fetched after them, and the summation order -- stored
//     entries in CSR order, then the self loop -- is the one of the row kernels, so results are bit-identical.
// Split ("heavy") rows contribute no entries and are not stored; the chunk kernels own them.
// ---------------------------------------------------------------------------
template <int VEC, int NCH, bool HAS_VAL, bool FULL, int MINB>
__global__ void __launch_bounds__(32, MINB)
spmm_rowgroup_kernel(const int64_t *__restrict__ rowptr, const int32_t *__restrict__ col,
                     const float *__restrict__ val, int64_t n, const float *__restrict__ x, int d, int64_t ldx,
                     float *__restrict__ y, int64_t ldy, int norm, int self_loop,
                     const float *__restrict__ dinv, const uint8_t *__restrict__ deg_code, int64_t threshold,
                     int64_t row_offset, RangeArgs ra, unsigned long long *__restrict__ next_group) {
  constexpr int TILE = 32 * VEC * NCH;
  constexpr int U0 = 32 / (VEC * NCH);
  constexpr int U = U0 > 16 ? 16 : (U0 < 2 ? 2 : U0);
  constexpr unsigned kFull = 0xffffffffu;
  __shared__ float table[256];
  __shared__ int s_off[2][33];      // stream offset of every row slot of the group (exclusive scan), [32] = total
  __shared__ int64_t s_beg[2][32];  // first stored entry of every row slot
  const int lane = threadIdx.x;
  const bool sym = (norm == WDGH_NORM_SYM);
  const bool coded = sym && deg_code != nullptr;
  const bool virt = self_loop && ra.finalize;  // self loop = virtual trailing entry of its row
  if (coded) {
    for (int c = lane; c < 256; c += 32) {
      double rs = (double)c + (self_loop ? 1.0 : 0.0);
      if (rs == 0.0) rs = 1.0;
      table[c] = (float)(1.0 / sqrt(rs));  // same expression as degree_scale_kernel -> same bits
    }
    __syncwarp();
  }
  const int cbase = FULL ? 0 : blockIdx.y * TILE;
  bool live[NCH];
#pragma unroll
  for (int t = 0; t < NCH; ++t) live[t] = FULL || (cbase + (t * 32 + lane) * VEC < d);
  const int64_t W = gridDim.x;
  const int64_t n_groups = (n + 31) >> 5;
  const int ld32 = (int)ldx;
  const float *xl = x + cbase + lane * VEC;

  auto load_bounds = [&](int64_t g, int64_t &b, int64_t &e, int &flag) {
    b = 0;
    e = 0;
    flag = 0;
    const int64_t r = (g << 5) + lane;
    if (r < n) {
      b = __ldg(rowptr + r);
      e = ra.row_end ? __ldg(ra.row_end + r) : __ldg(rowptr + r + 1);
      if (ra.skip) flag = __ldg(ra.skip + r);
    }
  };
  // scan the row lengths of group g into s_off[buf] / s_beg[buf]
  auto publish = [&](int buf, int64_t g, int64_t b, int64_t e, int flag, unsigned &nostore, int &total, float &si) {
    const int64_t r = (g << 5) + lane;
    const bool inr = r < n;
    const bool hv = inr && (flag != 0 || e - b > threshold);
    int inc = (!inr || hv) ? 0 : (int)(e - b) + (virt ? 1 : 0);
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int v = __shfl_up_sync(kFull, inc, o);
      if (lane >= o) inc += v;
    }
    s_off[buf][lane + 1] = inc;
    if (lane == 0) s_off[buf][0] = 0;
    s_beg[buf][lane] = b;
    nostore = __ballot_sync(kFull, !inr || hv);
    total = __shfl_sync(kFull, inc, 31);
    si = (norm != WDGH_NORM_NONE && ra.finalize && inr) ? __ldg(dinv + r + row_offset) : 1.f;
    __syncwarp();
  };
  // column id, weight and row slot of stream position t0 + lane of group g
  auto load_seg = [&](int buf, int64_t g, int t0, int total, int &j, float &w, int &rho) {
    j = 0;
    w = 0.f;
    rho = 0;
    const int t = t0 + lane;
    if (t < total) {
      int lo = 0;
#pragma unroll
      for (int step = 16; step >= 1; step >>= 1)
        if (s_off[buf][lo + step] <= t) lo += step;
      rho = lo;
      const int k = t - s_off[buf][lo];
      if (virt && k == s_off[buf][lo + 1] - s_off[buf][lo] - 1) {
        const int64_t grow = (g << 5) + lo + row_offset;
        j = (int)grow;
        w = sym ? __ldg(dinv + grow) : 1.f;
      } else {
        const int64_t idx = s_beg[buf][lo] + k;
        j = __ldg(col + idx);
        w = HAS_VAL ? __ldg(val + idx) : 1.f;
        if (coded) {
          const int c = __ldg(deg_code + j);
          w *= (c < 255) ? table[c] : __ldg(dinv + j);
        } else if (sym) {
          w *= __ldg(dinv + j);
        }
      }
    }
  };

  // group order: static stride W, or (next_group != nullptr) a ticket counter -- row lengths are heavy-tailed,
  // with tickets a warp that drew long rows simply takes fewer groups
  auto take = [&](int64_t prev) -> int64_t {
    if (next_group == nullptr || !FULL) return prev + W;
    unsigned long long t = 0;
    if (lane == 0) t = atomicAdd(next_group, 1ull);
    return (int64_t)__shfl_sync(kFull, t, 0) + W;  // tickets start after the W groups handed out by blockIdx
  };
  int64_t g = blockIdx.x;
  if (g >= n_groups) return;
  int64_t b, e;
  int flag;
  load_bounds(g, b, e, flag);
  int buf = 0;
  unsigned nostore, n_nostore = kFull;
  int total, n_total = 0;
  float si_l, n_si = 1.f;
  publish(0, g, b, e, flag, nostore, total, si_l);
  int64_t gn = take(g);
  int64_t gnn = 0;
  load_bounds(gn, b, e, flag);  // the next group's bounds travel while this group is aggregated
  int j, rho, nj = 0, nrho = 0;
  float w, nw = 0.f;
  load_seg(0, g, 0, total, j, w, rho);
  Vec<VEC> acc[NCH];
#pragma unroll
  for (int t = 0; t < NCH; ++t) acc[t].zero();

  // store row slot `cur` of the current group (accumulator -> y) and clear the accumulator
  auto flush_row = [&](int cur) {
    if (!((nostore >> cur) & 1u)) {
      const int64_t row = (g << 5) + cur;
      const float si = __shfl_sync(kFull, si_l, cur);
      const bool had = s_off[buf][cur + 1] != s_off[buf][cur];
      if (had || ra.finalize || !ra.accumulate) {  // an empty range adds nothing to an earlier phase's sum
#pragma unroll
        for (int t = 0; t < NCH; ++t) {
          if (live[t]) {
            const int c = cbase + (t * 32 + lane) * VEC;
            if (ra.accumulate) {
              Vec<VEC> prev;
              prev.load_plain(y + row * ldy + c);
              acc[t].add(prev);
            }
            if (ra.finalize) {
              acc[t].scale(si);
              acc[t].store_stream(y + row * ldy + c);
            } else {
              acc[t].store(y + row * ldy + c);
            }
          }
        }
      }
    }
#pragma unroll
    for (int t = 0; t < NCH; ++t) acc[t].zero();
  };

  while (true) {
    int cur = 0;
    bool next_ready = false;
    auto prefetch = [&](int t0, bool last_seg) {
      if (!last_seg) {
        load_seg(buf, g, t0 + 32, total, nj, nw, nrho);
      } else {  // next group: offsets, first segment, and the bounds of the group after it
        publish(buf ^ 1, gn, b, e, flag, n_nostore, n_total, n_si);
        load_seg(buf ^ 1, gn, 0, n_total, nj, nw, nrho);
        gnn = take(gn);
        load_bounds(gnn, b, e, flag);
        next_ready = true;
      }
    };
    for (int t0 = 0; t0 < total; t0 += 32) {
      const int cnt = min(32, total - t0);
      const bool last_seg = t0 + 32 >= total;
      bool next_issued = false;
      int k = 0;
      for (; k + U <= cnt; k += U) {  // full batches: U unpredicated gathers in flight
        Vec<VEC> v[U][NCH];
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const float *xr = xl + (int64_t)__shfl_sync(kFull, j, k + u) * ld32;
#pragma unroll
          for (int t = 0; t < NCH; ++t) {
            if (live[t]) v[u][t].load(xr + t * (32 * VEC));
            else v[u][t].zero();
          }
        }
        if (!next_issued) {
          prefetch(t0, last_seg);
          next_issued = true;
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const int ru = __shfl_sync(kFull, rho, k + u);
          while (cur < ru) flush_row(cur++);
          const float wu = __shfl_sync(kFull, w, k + u);
#pragma unroll
          for (int t = 0; t < NCH; ++t) acc[t].fma(wu, v[u][t]);
        }
      }
      if (k < cnt) {  // last batch of the group, predicated
        Vec<VEC> v[U][NCH];
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const bool on = k + u < cnt;
          const float *xr = xl + (int64_t)__shfl_sync(kFull, j, on ? k + u : 0) * ld32;
#pragma unroll
          for (int t = 0; t < NCH; ++t) {
            if (on && live[t]) v[u][t].load(xr + t * (32 * VEC));
            else v[u][t].zero();
          }
        }
        if (!next_issued) {
          prefetch(t0, last_seg);
          next_issued = true;
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
          if (k + u < cnt) {
            const int ru = __shfl_sync(kFull, rho, k + u);
            while (cur < ru) flush_row(cur++);
            const float wu = __shfl_sync(kFull, w, k + u);
#pragma unroll
            for (int t = 0; t < NCH; ++t) acc[t].fma(wu, v[u][t]);
          }
        }
      }
      j = nj; w = nw; rho = nrho;
    }
    if (!next_ready) {  // a group without entries
      prefetch(0, true);
      j = nj; w = nw; rho = nrho;
    }
    while (cur < 32) flush_row(cur++);
    g = gn;
    gn = gnn;
    buf ^= 1;
    nostore = n_nostore; total = n_total; si_l = n_si;
    if (g >= n_groups) break;
  }
}

// ---------------------------------------------------------------------------
// nnz-balanced streaming kernel (d % 4 == 0, d >= 128): one warp per UNIT of WDGH_UNIT consecutive
// stored entries, whatever rows they belong to.  Every warp always has full 32-entry segments and
// U gathers in flight, a hub row is just many units, and short rows cost no idle lanes -- this is
// what lets the gather stream approach the measured random-512 B-row ceiling of the part
// (tools/gather_bw.cu: ~7.0 TB/s).  Rows that cross a unit boundary leave unscaled partial sums in
// `partial` (head = first row of a unit that started earlier, tail = last row that continues) and
// are completed in fixed order by spmm_stream_fixup_kernel: deterministic, no float atomics.
// ---------------------------------------------------------------------------
template <int NCH, bool HAS_VAL>
__global__ void __launch_bounds__(32)
spmm_stream_kernel(const int64_t *__restrict__ rowptr, const int32_t *__restrict__ col,
                   const float *__restrict__ val, int64_t n, int64_t nnz, const float *__restrict__ x, int d,
                   int64_t ldx, float *__restrict__ y, int64_t ldy, int norm, int self_loop,
                   const float *__restrict__ dinv, const int64_t *__restrict__ plan, float *__restrict__ partial,
                   int64_t ldp, int64_t row_offset) {
  // Gathered rows are staged in shared memory with cp.async (LDGSTS): bytes in flight are bounded by
  // the 32 KB ring of this warp, not by registers.  SLOTS rows per stage, 2 stages: stage s+1 is in
  // flight while stage s is consumed.  Every lane reads back exactly the 16-byte pieces it copied
  // itself, so cp.async.wait_group is the only synchronisation needed.
  constexpr int SLOTS = 32 / NCH;
  constexpr int STAGES = 2;
  constexpr unsigned kFull = 0xffffffffu;
  __shared__ float4 ring[STAGES][SLOTS][32 * NCH];
  const int lane = threadIdx.x;
  const int64_t unit = blockIdx.x;
  const int cbase = blockIdx.y * (128 * NCH);
  bool live[NCH];
#pragma unroll
  for (int t = 0; t < NCH; ++t) live[t] = cbase + (t * 32 + lane) * 4 < d;
  const int64_t u0 = unit * WDGH_UNIT;
  const int64_t u1 = min(u0 + (int64_t)WDGH_UNIT, nnz);
  const int64_t r0 = (unit == 0) ? 0 : plan_unit_row(plan, plan[kPlanCapacity])[unit];
  const bool first_partial = __ldg(rowptr + r0) < u0;
  const float *tscale = (norm == WDGH_NORM_SYM) ? dinv : nullptr;

  // All positions below are 32-bit offsets: rows relative to r0, entries relative to u0.
  // window of row ends: lane l holds (rowptr[r0 + rbase + 1 + l] - u0), clipped
  const int n_rel = (int)(n - r0);          // rows from r0 to the end of the matrix
  const int len = (int)(u1 - u0);           // entries in this unit (<= WDGH_UNIT)
  int cur = 0, rbase = 0;                   // current row / window base, relative to r0
  auto load_ends = [&](int base) -> int {
    const int i = base + 1 + lane;
    if (i > n_rel) return INT_MAX;
    const int64_t e = __ldg(rowptr + r0 + i) - u0;
    return e > (int64_t)INT_MAX ? INT_MAX : (int)e;
  };
  int ends = load_ends(0);
  int cur_end = __shfl_sync(kFull, ends, 0);
  // finalisation data of the current row, fetched when the row becomes current
  float si = 1.f;
  Vec<4> xi[NCH];
  auto fetch_row = [&]() {
    if (cur < n_rel) {
      const int64_t gr = r0 + cur + row_offset;
      if (norm != WDGH_NORM_NONE) si = __ldg(dinv + gr);
      if (self_loop) {
#pragma unroll
        for (int t = 0; t < NCH; ++t)
          if (live[t]) xi[t].load(x + gr * ldx + cbase + (t * 32 + lane) * 4);
      }
    }
  };
  fetch_row();
  Vec<4> acc[NCH];
#pragma unroll
  for (int t = 0; t < NCH; ++t) acc[t].zero();

  auto flush = [&]() {  // the current row ends here
    if (cur == 0 && first_partial) {
      float *head = partial + (2 * unit) * ldp + cbase;
#pragma unroll
      for (int t = 0; t < NCH; ++t)
        if (live[t]) acc[t].store(head + (t * 32 + lane) * 4);
    } else {
      const float self_w = (norm == WDGH_NORM_SYM) ? si : 1.f;
      float *yr = y + (r0 + cur) * ldy + cbase;
#pragma unroll
      for (int t = 0; t < NCH; ++t) {
        if (live[t]) {
          if (self_loop) acc[t].fma(self_w, xi[t]);
          acc[t].scale(si);
          acc[t].store_stream(yr + (t * 32 + lane) * 4);
        }
      }
    }
#pragma unroll
    for (int t = 0; t < NCH; ++t) acc[t].zero();
    ++cur;
    if (cur - rbase == 32) {
      rbase = cur;
      ends = load_ends(rbase);
    }
    cur_end = __shfl_sync(kFull, ends, cur - rbase);
    fetch_row();
  };

  // segment = SLOTS consecutive entries; lane l (< SLOTS) holds column id and weight of entry l
  auto load_seg = [&](int seg, int &j, float &w) {
    j = 0;
    w = 0.f;
    const int p = seg * SLOTS + lane;
    if (lane < SLOTS && p < len) {
      const int64_t idx = u0 + p;
      j = __ldg(col + idx);
      w = HAS_VAL ? __ldg(val + idx) : 1.f;
      if (tscale != nullptr) w *= __ldg(tscale + j);
    }
  };
  auto issue_seg = [&](int seg, int j) {  // async copies of the segment's feature rows into its stage
    const int cnt = min(SLOTS, len - seg * SLOTS);
    float4(*stage)[32 * NCH] = ring[seg % STAGES];
#pragma unroll
    for (int k = 0; k < SLOTS; ++k) {
      const int jj = __shfl_sync(kFull, j, k);
      if (k < cnt) {
        const float *xr = x + (int64_t)jj * ldx + cbase;
#pragma unroll
        for (int t = 0; t < NCH; ++t) {
          if (live[t]) {
            const unsigned dst = (unsigned)__cvta_generic_to_shared(&stage[k][t * 32 + lane]);
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(xr + (t * 32 + lane) * 4)
                         : "memory");
          }
        }
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };

  while (cur < n_rel && cur_end <= 0) flush();  // leading empty rows (only unit 0 can have them)

  const int nseg = (len + SLOTS - 1) / SLOTS;
  int j0, j1 = 0;
  float w0, w1 = 0.f;
  load_seg(0, j0, w0);
  if (nseg > 1) load_seg(1, j1, w1);
  issue_seg(0, j0);
  for (int seg = 0; seg < nseg; ++seg) {
    int j2 = 0;
    float w2 = 0.f;
    if (seg + 2 < nseg) load_seg(seg + 2, j2, w2);  // ids two segments ahead: ready when their copies are issued
    if (seg + 1 < nseg) issue_seg(seg + 1, j1);
    else asm volatile("cp.async.commit_group;" ::: "memory");  // keep one group per iteration
    asm volatile("cp.async.wait_group 1;" ::: "memory");
    const int cnt = min(SLOTS, len - seg * SLOTS);
    float4(*stage)[32 * NCH] = ring[seg % STAGES];
#pragma unroll 8
    for (int k = 0; k < cnt; ++k) {
      const float wk = __shfl_sync(kFull, w0, k);
#pragma unroll
      for (int t = 0; t < NCH; ++t) {
        if (live[t]) {
          Vec<4> v;
          v.v = stage[k][t * 32 + lane];
          acc[t].fma(wk, v);
        }
      }
      const int next_pos = seg * SLOTS + k + 1;
      while (cur < n_rel && next_pos == cur_end) flush();  // also walks over empty rows that follow
    }
    j0 = j1; w0 = w1;
    j1 = j2; w1 = w2;
  }
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  if (cur < n_rel) {  // the last row of the unit continues in the next unit
    // head: the unit lies inside one row that started earlier; tail otherwise
    float *dst = partial + (2 * unit + ((cur == 0 && first_partial) ? 0 : 1)) * ldp + cbase;
#pragma unroll
    for (int t = 0; t < NCH; ++t)
      if (live[t]) acc[t].store(dst + (t * 32 + lane) * 4);
  }
}

// One warp per unit: if the unit's first row started in an earlier unit and ends here, sum its
// partials in unit order (tail of the first unit, heads of the units in between, head of this one).
__global__ void __launch_bounds__(128)
spmm_stream_fixup_kernel(const int64_t *__restrict__ rowptr, int64_t nnz, const float *__restrict__ x, int d,
                         int64_t ldx, float *__restrict__ y, int64_t ldy, int norm, int self_loop,
                         const float *__restrict__ dinv, const int64_t *__restrict__ plan,
                         const float *__restrict__ partial, int64_t ldp, int64_t row_offset) {
  const int lane = threadIdx.x & 31;
  const int64_t unit = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (unit < 1 || unit >= plan[kPlanNUnits]) return;
  const int64_t u0 = unit * WDGH_UNIT;
  const int64_t u1 = min(u0 + (int64_t)WDGH_UNIT, nnz);
  const int64_t r0 = plan_unit_row(plan, plan[kPlanCapacity])[unit];
  const int64_t s0 = __ldg(rowptr + r0), e0 = __ldg(rowptr + r0 + 1);
  if (!(s0 < u0 && e0 <= u1)) return;  // the row does not close in this unit
  const int64_t uf = s0 / WDGH_UNIT;
  const float si = (norm != WDGH_NORM_NONE) ? __ldg(dinv + r0 + row_offset) : 1.f;
  const float self_w = (norm == WDGH_NORM_SYM) ? si : 1.f;
  for (int c = lane; c < d; c += 32) {
    float acc = partial[(2 * uf + 1) * ldp + c];
    for (int64_t v = uf + 1; v <= unit; ++v) acc += partial[(2 * v) * ldp + c];
    if (self_loop) acc = fmaf(self_w, __ldg(x + (r0 + row_offset) * ldx + c), acc);
    y[r0 * ldy + c] = acc * si;
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
  const uint8_t *deg_code;
  const int64_t *plan;
  int64_t threshold, n_heavy, n_chunks, row_offset, nnz, n_units;
  bool stats = false;
  StatsArgs sa = {nullptr, 0, nullptr, nullptr, nullptr};
  RangeArgs ra = {nullptr, nullptr, 0, 1};
  bool heavy_pass = true;  // run the split-row chunk kernels
  float *partial;
  int64_t ldp;
  cudaStream_t st;
};

// CTA size of the row kernel.  Small CTAs keep warp slots busy when row lengths are skewed (a CTA's
// slots are only recycled when its longest row finishes).  WDGH_SPMM_BLOCK overrides for experiments.
static int rows_block_threads() {
  static int cached = 0;
  if (cached == 0) {
    int v = 32;  // measured on B200 (1B-entry power-law graph): 256 -> 123.8 ms, 128 -> 114.7, 64 -> 112.4, 32 -> 111.3
    if (const char *e = getenv("WDGH_SPMM_BLOCK")) v = atoi(e);
    cached = (v == 32 || v == 64 || v == 128 || v == 256) ? v : 32;
  }
  return cached;
}

template <int G, int VEC, int NCH, bool HAS_VAL>
static int launch_rows(const SpmmArgs &a) {
  constexpr int RPW = 32 / G;
  const int block = rows_block_threads();
  const int64_t rows_per_cta = (block / 32) * RPW;
  const int tile = G * VEC * NCH;
  dim3 grid((unsigned)ceil_div(a.n, rows_per_cta), (unsigned)ceil_div(a.d, tile));
  spmm_rows_kernel<G, VEC, NCH, HAS_VAL><<<grid, block, 0, a.st>>>(a.rowptr, a.col, a.val, a.n, a.x, a.d, a.ldx, a.y,
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
                                                                  a.row_offset, a.ra.finalize);
  WDGH_LAUNCHED("spmm_heavy_finish_kernel");
  return 0;
}

template <int NCH, bool HAS_VAL>
static int launch_stream(const SpmmArgs &a) {
  // one warp (one unit) per CTA, 32 KB of static shared memory each: up to 7 CTAs = 224 KB in flight per SM
  static bool configured = false;
  if (!configured) {
    cudaFuncSetAttribute(spmm_stream_kernel<NCH, HAS_VAL>, cudaFuncAttributePreferredSharedMemoryCarveout,
                         cudaSharedmemCarveoutMaxShared);
    configured = true;
  }
  dim3 grid((unsigned)a.n_units, (unsigned)ceil_div(a.d, 128 * NCH));
  spmm_stream_kernel<NCH, HAS_VAL><<<grid, 32, 0, a.st>>>(a.rowptr, a.col, a.val, a.n, a.nnz, a.x, a.d, a.ldx, a.y,
                                                         a.ldy, a.norm, a.self_loop, a.dinv, a.plan, a.partial, a.ldp,
                                                         a.row_offset);
  WDGH_LAUNCHED("spmm_stream_kernel");
  if (a.n_units > 1) {
    spmm_stream_fixup_kernel<<<(unsigned)ceil_div(a.n_units, 4), 128, 0, a.st>>>(
        a.rowptr, a.nnz, a.x, a.d, a.ldx, a.y, a.ldy, a.norm, a.self_loop, a.dinv, a.plan, a.partial, a.ldp,
        a.row_offset);
    WDGH_LAUNCHED("spmm_stream_fixup_kernel");
  }
  return 0;
}

// WDGH_SPMM_VARIANT: 0 = plain row kernels, 1 (default) = persistent kernels (row groups, or per-row pipelined with
// WDGH_ROWGROUP=0), 2 = nnz-balanced cp.async stream
static int wide_variant() {
  static int cached = -1;
  if (cached < 0) {
    const char *e = getenv("WDGH_SPMM_VARIANT");
    cached = e ? atoi(e) : 1;
    if (cached < 0 || cached > 2) cached = 1;
  }
  return cached;
}
static int pipelined_minb() {
  static int cached = 0;
  if (cached == 0) {
    int v = 32;  // measured (1B-entry graph, d=128): 16 -> 118.1 ms, 24 -> 100.4 ms, 32 -> 95.6 ms
    if (const char *e = getenv("WDGH_PIPE_MINB")) v = atoi(e);
    cached = (v == 16 || v == 24 || v == 32) ? v : 32;
  }
  return cached;
}

// WDGH_ROWGROUP: 2 (default) = row-group kernel, groups handed out by a ticket counter; 1 = row-group kernel,
// static stride; 0 = the per-row pipelined kernel.  Measured on rows [0, 6.25M) of the 50M-node bench graph, d=128:
//   full rows (19.5 entries/row):      0 -> 11.34 ms, 1 -> 11.62 ms, 2 -> 10.73 ms
//   one 2-D row slice (4.9 entries/row): 0 -> 3.98 ms, 1 -> 3.87 ms, 2 -> 3.63 ms
static int rowgroup_enabled() {
  static int cached = -1;
  if (cached < 0) {
    const char *e = getenv("WDGH_ROWGROUP");
    cached = e ? atoi(e) : 2;
    if (cached < 0 || cached > 2) cached = 2;
  }
  return cached;
}
template <int VEC, int NCH, bool HAS_VAL>
static int launch_rowgroup(const SpmmArgs &a) {
  constexpr int TILE = 32 * VEC * NCH;
  const int minb = (NCH == 1) ? pipelined_minb() : 16;
  const int64_t n_groups = (a.n + 31) / 32;
  int64_t ctas = (int64_t)sm_count() * minb;
  if (ctas > n_groups) ctas = n_groups;
  dim3 grid((unsigned)ctas, (unsigned)ceil_div(a.d, TILE));
  const bool full = (a.d == TILE);
  unsigned long long *tickets = (rowgroup_enabled() == 2 && full) ? ticket_slot(a.st) : nullptr;
#define WDGH_RG_LAUNCH(FULLV, MINB)                                                                                \
  spmm_rowgroup_kernel<VEC, NCH, HAS_VAL, FULLV, MINB><<<grid, 32, 0, a.st>>>(                                       \
      a.rowptr, a.col, a.val, a.n, a.x, a.d, a.ldx, a.y, a.ldy, a.norm, a.self_loop, a.dinv, a.deg_code, a.threshold, \
      a.row_offset, a.ra, tickets)
  if (full) {
    switch (minb) {
      case 32: WDGH_RG_LAUNCH(true, 32); break;
      case 24: WDGH_RG_LAUNCH(true, 24); break;
      default: WDGH_RG_LAUNCH(true, 16); break;
    }
  } else {
    WDGH_RG_LAUNCH(false, 16);
  }
#undef WDGH_RG_LAUNCH
  WDGH_LAUNCHED("spmm_rowgroup_kernel");
  return 0;
}

template <int VEC, int NCH, bool HAS_VAL>
static int launch_pipelined(const SpmmArgs &a) {
  if (!a.stats && rowgroup_enabled()) return launch_rowgroup<VEC, NCH, HAS_VAL>(a);
  constexpr int TILE = 32 * VEC * NCH;
  const int minb = (NCH == 1) ? pipelined_minb() : 16;
  int64_t ctas = (int64_t)sm_count() * minb;
  if (ctas > a.n) ctas = a.n;
  dim3 grid((unsigned)ctas, (unsigned)ceil_div(a.d, TILE));
  const bool full = (a.d == TILE);
#define WDGH_PIPE_LAUNCH(FULLV, MINB)                                                                              \
  spmm_rows_pipelined_kernel<VEC, NCH, HAS_VAL, FULLV, MINB, false><<<grid, 32, 0, a.st>>>(                           \
      a.rowptr, a.col, a.val, a.n, a.x, a.d, a.ldx, a.y, a.ldy, a.norm, a.self_loop, a.dinv, a.deg_code, a.threshold, \
      a.row_offset, a.sa, a.ra)
#define WDGH_PIPE_LAUNCH_STATS(FULLV, MINB)                                                                        \
  spmm_rows_pipelined_kernel<4, NCH, false, FULLV, MINB, true>                                                        \
      <<<grid, 32, (a.sa.C * a.sa.C + 4) * sizeof(unsigned), a.st>>>(a.rowptr, a.col, a.val, a.n, a.x, a.d, a.ldx, a.y, \
                                                                      a.ldy, a.norm, a.self_loop, a.dinv, a.deg_code, \
                                                                      a.threshold, a.row_offset, a.sa, a.ra)
  if (a.stats && VEC == 4) {  // binary adjacency only (checked by the caller)
    if (full && NCH == 1) WDGH_PIPE_LAUNCH_STATS(true, 32);
    else if (full) WDGH_PIPE_LAUNCH_STATS(true, 16);
    else WDGH_PIPE_LAUNCH_STATS(false, 16);
  } else if (full) {
    switch (minb) {
      case 32: WDGH_PIPE_LAUNCH(true, 32); break;
      case 24: WDGH_PIPE_LAUNCH(true, 24); break;
      default: WDGH_PIPE_LAUNCH(true, 16); break;
    }
  } else {
    WDGH_PIPE_LAUNCH(false, 16);
  }
#undef WDGH_PIPE_LAUNCH
#undef WDGH_PIPE_LAUNCH_STATS
  WDGH_LAUNCHED("spmm_rows_pipelined_kernel");
  return 0;
}

template <bool HAS_VAL>
static int dispatch(const SpmmArgs &a, bool vec4) {
  int rc;
  const int d = a.d;
  if (vec4 && (d >= 128 || d == 64 || d == 32) && wide_variant() == 1) {
    if (d == 32) rc = launch_pipelined<1, 1, HAS_VAL>(a);
    else if (d == 64) rc = launch_pipelined<2, 1, HAS_VAL>(a);
    else if (d <= 128) rc = launch_pipelined<4, 1, HAS_VAL>(a);
    else if (d <= 256) rc = launch_pipelined<4, 2, HAS_VAL>(a);
    else rc = launch_pipelined<4, 4, HAS_VAL>(a);
    if (rc || !a.heavy_pass) return rc;
    if (d <= 128) return launch_heavy<4, 1, HAS_VAL>(a);
    if (d <= 256) return launch_heavy<4, 2, HAS_VAL>(a);
    return launch_heavy<4, 4, HAS_VAL>(a);
  }
  if (vec4 && d >= 128 && a.n_units > 0 && wide_variant() == 2) {
    if (d <= 128) return launch_stream<1, HAS_VAL>(a);
    if (d <= 256) return launch_stream<2, HAS_VAL>(a);
    return launch_stream<4, HAS_VAL>(a);
  }
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
                             int add_self_loop, const float *dinv, const uint8_t *deg_code,
                             const int64_t *plan_i64, const int64_t *plan_host, float *partial, int64_t row_offset,
                             void *stream) {
  WDGH_REQUIRE(rowptr && x && y && plan_i64 && plan_host, "wdgh_spmm_csr: null pointer");  // col may be NULL iff nnz == 0
  WDGH_REQUIRE(n >= 0 && d > 0 && d <= (1 << 24) && ldx >= d && ldy >= d, "wdgh_spmm_csr: bad shape");
  WDGH_REQUIRE(norm == WDGH_NORM_NONE || norm == WDGH_NORM_RW || norm == WDGH_NORM_SYM, "wdgh_spmm_csr: bad norm");
  WDGH_REQUIRE(norm == WDGH_NORM_NONE || dinv != nullptr, "wdgh_spmm_csr: norm requires dinv");
  WDGH_REQUIRE(x != y, "wdgh_spmm_csr: in-place aggregation is not supported");
  WDGH_REQUIRE(row_offset >= 0, "wdgh_spmm_csr: negative row_offset");
  if (n == 0) return 0;
  SpmmArgs a;
  a.rowptr = rowptr; a.col = col; a.val = val; a.n = n; a.x = x; a.d = (int)d; a.ldx = ldx; a.y = y; a.ldy = ldy;
  a.norm = norm; a.self_loop = add_self_loop ? 1 : 0; a.dinv = dinv; a.deg_code = (val == nullptr) ? deg_code : nullptr; a.plan = plan_i64;
  a.n_heavy = plan_host[0]; a.n_chunks = plan_host[1]; a.threshold = plan_host[2];
  a.n_units = plan_host[5];
  a.nnz = plan_host[6];
  a.partial = partial; a.ldp = (d + 3) & ~int64_t(3);
  a.row_offset = row_offset;
  a.st = as_stream(stream);
  WDGH_REQUIRE(a.n_chunks == 0 || partial != nullptr, "wdgh_spmm_csr: split rows need the partial buffer");
  const bool vec4 = (d % 4 == 0) && (ldx % 4 == 0) && (ldy % 4 == 0) &&
                    (reinterpret_cast<uintptr_t>(x) % 16 == 0) && (reinterpret_cast<uintptr_t>(y) % 16 == 0) &&
                    (partial == nullptr || reinterpret_cast<uintptr_t>(partial) % 16 == 0);
  return val ? dispatch<true>(a, vec4) : dispatch<false>(a, vec4);
}

extern "C" int wdgh_structure_counts(const int64_t *, const int32_t *, int64_t, int64_t, const int32_t *, int32_t,
                                     const int64_t *, const int64_t *, int64_t *, double *, int32_t *, int32_t *,
                                     uint8_t *, int64_t, int64_t, void *);

extern "C" int wdgh_spmm_structure_fused(const int64_t *rowptr, const int32_t *col, int64_t n, int64_t nnz,
                                         const float *x, int64_t d, int64_t ldx, float *y, int64_t ldy, int norm,
                                         int add_self_loop, const float *dinv, const uint8_t *deg_code,
                                         const int32_t *labels, int32_t num_classes, const int64_t *plan_i64,
                                         const int64_t *plan_host, float *partial, int64_t *counters, double *node_sum,
                                         int32_t *deg_nsl, int32_t *match_nsl, uint8_t *labels_u8_scratch,
                                         int64_t n_labels, int64_t row_offset, int single_kernel, void *stream) {
  WDGH_REQUIRE(rowptr && x && y && labels && plan_i64 && plan_host && counters && node_sum && deg_nsl && match_nsl,
               "wdgh_spmm_structure_fused: null pointer");
  WDGH_REQUIRE(n >= 0 && nnz >= 0 && d > 0 && d <= (1 << 24) && ldx >= d && ldy >= d, "wdgh_spmm_structure_fused: bad shape");
  WDGH_REQUIRE(num_classes >= 1 && num_classes <= 46340, "wdgh_spmm_structure_fused: num_classes out of range");
  WDGH_REQUIRE(norm == WDGH_NORM_NONE || dinv != nullptr, "wdgh_spmm_structure_fused: norm requires dinv");
  const int C = num_classes;
  const bool vec4 = (d % 4 == 0) && (ldx % 4 == 0) && (ldy % 4 == 0) && (reinterpret_cast<uintptr_t>(x) % 16 == 0) &&
                    (reinterpret_cast<uintptr_t>(y) % 16 == 0) &&
                    (partial == nullptr || reinterpret_cast<uintptr_t>(partial) % 16 == 0);
  // Measured on the B200 (1B-entry graph): the label work inside the aggregation kernel costs more than the
  // separate 10 ms edge pass (111.3 vs 105.9 ms per step) because the gather kernel is issue- and
  // power-sensitive, so the single-kernel form is opt-in.
  const bool fusable = single_kernel != 0 && vec4 && d >= 128 && n > 0 && labels_u8_scratch != nullptr && C <= 254 &&
                       C * C <= kHistSmemBins && wide_variant() == 1 && n_labels >= n + row_offset;
  if (!fusable) {  // same results from the two separate passes
    int rc = wdgh_spmm_csr(rowptr, col, nullptr, n, x, d, ldx, y, ldy, norm, add_self_loop, dinv, deg_code, plan_i64,
                           plan_host, partial, row_offset, stream);
    if (rc) return rc;
    return wdgh_structure_counts(rowptr, col, n, nnz, labels, num_classes, plan_i64, plan_host, counters, node_sum,
                                 deg_nsl, match_nsl, labels_u8_scratch, n_labels, row_offset, stream);
  }
  cudaStream_t st = as_stream(stream);
  const uint8_t *labels8 = nullptr;
  int rc = structure_prepare(labels, n_labels, C, labels_u8_scratch, counters, node_sum, &labels8, st);
  if (rc) return rc;
  WDGH_CUDA(cudaMemsetAsync(deg_nsl, 0, n * sizeof(int32_t), st));    // per-row counts are accumulated with reductions
  WDGH_CUDA(cudaMemsetAsync(match_nsl, 0, n * sizeof(int32_t), st));
  SpmmArgs a;
  a.rowptr = rowptr; a.col = col; a.val = nullptr; a.n = n; a.x = x; a.d = (int)d; a.ldx = ldx; a.y = y; a.ldy = ldy;
  a.norm = norm; a.self_loop = add_self_loop ? 1 : 0; a.dinv = dinv; a.deg_code = deg_code; a.plan = plan_i64;
  a.n_heavy = plan_host[0]; a.n_chunks = plan_host[1]; a.threshold = plan_host[2];
  a.n_units = plan_host[5]; a.nnz = plan_host[6];
  a.partial = partial; a.ldp = (d + 3) & ~int64_t(3);
  a.row_offset = row_offset;
  a.st = st;
  a.stats = true;
  a.sa.labels8 = labels8; a.sa.C = C; a.sa.counters = reinterpret_cast<unsigned long long *>(counters);
  a.sa.deg_nsl = deg_nsl; a.sa.match_nsl = match_nsl;
  WDGH_REQUIRE(a.n_chunks == 0 || partial != nullptr, "wdgh_spmm_structure_fused: split rows need the partial buffer");
  rc = dispatch<false>(a, true);
  if (rc) return rc;
  return structure_finish(rowptr, col, n, labels, labels8, C, plan_i64, plan_host, counters, node_sum, deg_nsl,
                          match_nsl, row_offset, st);
}

// ---------------------------------------------------------------------------
// Ranged / phased aggregation (multi-GPU overlap): see RangeArgs.
// ---------------------------------------------------------------------------
namespace wdgh {
__global__ void column_segments_kernel(const int64_t *__restrict__ rowptr, const int32_t *__restrict__ col, int64_t n,
                                       const int64_t *__restrict__ bounds, int nb, int64_t *__restrict__ seg) {
  // seg[b][r] = first entry of row r whose column is >= bounds[b]   (thread per row, binary search per boundary)
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < n; r += stride) {
    const int64_t s = rowptr[r], e = rowptr[r + 1];
    for (int b = 0; b < nb; ++b) {
      const int64_t key = bounds[b];
      int64_t lo = s, hi = e;
      while (lo < hi) {
        const int64_t mid = (lo + hi) >> 1;
        if ((int64_t)col[mid] < key) lo = mid + 1;
        else hi = mid;
      }
      seg[(int64_t)b * n + r] = lo;
    }
  }
}
__global__ void heavy_flags_kernel(const int64_t *__restrict__ plan, uint8_t *__restrict__ flags) {
  const int64_t n_heavy = plan[kPlanNHeavy];
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < n_heavy; k += stride)
    flags[plan_heavy_row(plan)[k]] = 1;
}
}  // namespace wdgh

extern "C" int wdgh_column_segments(const int64_t *rowptr, const int32_t *col, int64_t n, const int64_t *bounds_dev,
                                    int32_t num_bounds, int64_t *seg, void *stream) {
  WDGH_REQUIRE(rowptr && bounds_dev && seg && n >= 0 && num_bounds >= 1, "wdgh_column_segments: bad arguments");
  if (n == 0) return 0;
  column_segments_kernel<<<persistent_grid(ceil_div(n, 256), 8), 256, 0, as_stream(stream)>>>(rowptr, col, n, bounds_dev,
                                                                                            num_bounds, seg);
  WDGH_LAUNCHED("column_segments_kernel");
  return 0;
}

extern "C" int wdgh_plan_heavy_flags(const int64_t *plan_i64, const int64_t *plan_host, int64_t n, uint8_t *flags,
                                     void *stream) {
  WDGH_REQUIRE(plan_i64 && plan_host && flags && n >= 0, "wdgh_plan_heavy_flags: bad arguments");
  cudaStream_t st = as_stream(stream);
  WDGH_CUDA(cudaMemsetAsync(flags, 0, (size_t)n, st));
  if (plan_host[0] > 0) {
    heavy_flags_kernel<<<persistent_grid(ceil_div(plan_host[0], 256), 4), 256, 0, st>>>(plan_i64, flags);
    WDGH_LAUNCHED("heavy_flags_kernel");
  }
  return 0;
}

extern "C" int wdgh_spmm_csr_ranged(const int64_t *rowptr, const int64_t *range_begin, const int64_t *range_end,
                                    const int32_t *col, const float *val, int64_t n, const float *x, int64_t d,
                                    int64_t ldx, float *y, int64_t ldy, int norm, int add_self_loop,
                                    const float *dinv, const uint8_t *deg_code, const uint8_t *skip_rows,
                                    int accumulate, int finalize, int run_split_rows, const int64_t *plan_i64,
                                    const int64_t *plan_host, float *partial, int64_t row_offset, void *stream) {
  WDGH_REQUIRE(rowptr && range_begin && range_end && x && y && plan_i64 && plan_host, "wdgh_spmm_csr_ranged: null pointer");
  WDGH_REQUIRE(n >= 0 && d > 0 && ldx >= d && ldy >= d, "wdgh_spmm_csr_ranged: bad shape");
  WDGH_REQUIRE(norm == WDGH_NORM_NONE || dinv != nullptr, "wdgh_spmm_csr_ranged: norm requires dinv");
  const bool vec4 = (d % 4 == 0) && (ldx % 4 == 0) && (ldy % 4 == 0) && (reinterpret_cast<uintptr_t>(x) % 16 == 0) &&
                    (reinterpret_cast<uintptr_t>(y) % 16 == 0) &&
                    (partial == nullptr || reinterpret_cast<uintptr_t>(partial) % 16 == 0);
  WDGH_REQUIRE(vec4 && (d >= 128 || d == 64 || d == 32) && wide_variant() == 1,
               "wdgh_spmm_csr_ranged: needs 16-byte aligned rows and d in {32, 64} or d >= 128");
  WDGH_REQUIRE(plan_host[1] == 0 || skip_rows != nullptr, "wdgh_spmm_csr_ranged: split rows need skip_rows (wdgh_plan_heavy_flags)");
  if (n == 0) return 0;
  SpmmArgs a;
  a.rowptr = range_begin; a.col = col; a.val = val; a.n = n; a.x = x; a.d = (int)d; a.ldx = ldx; a.y = y; a.ldy = ldy;
  a.norm = norm; a.self_loop = add_self_loop ? 1 : 0; a.dinv = dinv; a.deg_code = (val == nullptr) ? deg_code : nullptr;
  a.plan = plan_i64;
  a.n_heavy = plan_host[0]; a.n_chunks = plan_host[1]; a.threshold = INT64_MAX / 4;  // ranges are never "heavy" by length
  a.n_units = 0; a.nnz = plan_host[6];
  a.partial = partial; a.ldp = (d + 3) & ~int64_t(3);
  a.row_offset = row_offset;
  a.st = as_stream(stream);
  a.ra.row_end = range_end; a.ra.skip = skip_rows; a.ra.accumulate = accumulate ? 1 : 0; a.ra.finalize = finalize ? 1 : 0;
  a.heavy_pass = false;
  int rc = val ? dispatch<true>(a, true) : dispatch<false>(a, true);
  if (rc || !run_split_rows || a.n_chunks == 0) return rc;
  // split rows: always over their full column range, after the last phase (they overwrite their Y rows)
  WDGH_REQUIRE(partial != nullptr, "wdgh_spmm_csr_ranged: split rows need the partial buffer");
  a.rowptr = rowptr; a.threshold = plan_host[2]; a.ra = RangeArgs{nullptr, nullptr, 0, finalize ? 1 : 0};
  if (val) {
    if (d <= 128) return launch_heavy<4, 1, true>(a);
    if (d <= 256) return launch_heavy<4, 2, true>(a);
    return launch_heavy<4, 4, true>(a);
  }
  if (d <= 128) return launch_heavy<4, 1, false>(a);
  if (d <= 256) return launch_heavy<4, 2, false>(a);
  return launch_heavy<4, 4, false>(a);
}

// ---------------------------------------------------------------------------
// 2-D partition: y[r] = s_r * (sum_p part_p[r] + [self loop] t_r * x[r]) for a row slice -- the reduction of the
// partial aggregations of one row group (own partial + the slices pulled from the peers) fused with the epilogue.
// ---------------------------------------------------------------------------
namespace wdgh {
struct PartPtrs {
  const float *p[16];
};
__global__ void __launch_bounds__(256)
reduce_finalize_kernel(PartPtrs parts, int n_parts, int64_t rows, int d4, int64_t ldp4, const float *__restrict__ x,
                       int64_t ldx4, float *__restrict__ y, int64_t ldy4, int norm, int self_loop,
                       const float *__restrict__ dinv, int64_t row_offset) {
  const int64_t total = rows * d4;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += stride) {
    const int64_t r = t / d4;
    const int c = (int)(t - r * d4);
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int p = 0; p < n_parts; ++p) {
      const float4 v = ldg_na(reinterpret_cast<const float4 *>(parts.p[p]) + r * ldp4 + c);
      acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
    const float si = (norm != WDGH_NORM_NONE) ? __ldg(dinv + r + row_offset) : 1.f;
    if (self_loop) {
      const float sw = (norm == WDGH_NORM_SYM) ? si : 1.f;
      const float4 xi = ldg_na(reinterpret_cast<const float4 *>(x) + (r + row_offset) * ldx4 + c);
      acc.x = fmaf(sw, xi.x, acc.x); acc.y = fmaf(sw, xi.y, acc.y); acc.z = fmaf(sw, xi.z, acc.z); acc.w = fmaf(sw, xi.w, acc.w);
    }
    acc.x *= si; acc.y *= si; acc.z *= si; acc.w *= si;
    st_cs(reinterpret_cast<float4 *>(y) + r * ldy4 + c, acc);
  }
}
}  // namespace wdgh

extern "C" int wdgh_reduce_finalize(const float *const *parts_host, int32_t n_parts, int64_t rows, int64_t d,
                                    int64_t ld_parts, const float *x, int64_t ldx, float *y, int64_t ldy, int norm,
                                    int add_self_loop, const float *dinv, int64_t row_offset, void *stream) {
  WDGH_REQUIRE(parts_host && n_parts >= 1 && n_parts <= 16 && y && rows >= 0 && d > 0, "wdgh_reduce_finalize: bad arguments");
  WDGH_REQUIRE(d % 4 == 0 && ld_parts % 4 == 0 && ldx % 4 == 0 && ldy % 4 == 0, "wdgh_reduce_finalize: needs 16-byte rows");
  WDGH_REQUIRE(norm == WDGH_NORM_NONE || dinv != nullptr, "wdgh_reduce_finalize: norm requires dinv");
  WDGH_REQUIRE(!add_self_loop || x != nullptr, "wdgh_reduce_finalize: self loop requires x");
  if (rows == 0) return 0;
  PartPtrs pp;
  for (int i = 0; i < 16; ++i) pp.p[i] = i < n_parts ? parts_host[i] : nullptr;
  reduce_finalize_kernel<<<persistent_grid(ceil_div(rows * (d / 4), 256), 8), 256, 0, as_stream(stream)>>>(
      pp, n_parts, rows, (int)(d / 4), ld_parts / 4, x, ldx / 4, y, ldy / 4, norm, add_self_loop ? 1 : 0, dinv, row_offset);
  WDGH_LAUNCHED("reduce_finalize_kernel");
  return 0;
}
