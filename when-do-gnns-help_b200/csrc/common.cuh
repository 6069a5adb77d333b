// Shared helpers for the wdgh_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include <atomic>

#include "../../include/wdgh_b200.h"

namespace wdgh {

constexpr int kWarp = 32;

// ---- error plumbing -------------------------------------------------------
extern thread_local char g_err[512];
extern std::atomic<uint64_t> g_launches;

inline int fail(int code, const char *msg) {
  snprintf(g_err, sizeof(g_err), "%s", msg);
  return code;
}
inline int fail_cuda(cudaError_t e, const char *where) {
  snprintf(g_err, sizeof(g_err), "%s: %s", where, cudaGetErrorString(e));
  return (int)e;
}
#define WDGH_CUDA(call)                                      \
  do {                                                       \
    cudaError_t e__ = (call);                                \
    if (e__ != cudaSuccess) return ::wdgh::fail_cuda(e__, #call); \
  } while (0)
// after a kernel launch: count it and surface launch-configuration errors
#define WDGH_LAUNCHED(name)                                  \
  do {                                                       \
    ::wdgh::g_launches.fetch_add(1, std::memory_order_relaxed); \
    cudaError_t e__ = cudaGetLastError();                    \
    if (e__ != cudaSuccess) return ::wdgh::fail_cuda(e__, name); \
  } while (0)
#define WDGH_REQUIRE(cond, msg) \
  do {                          \
    if (!(cond)) return ::wdgh::fail(WDGH_EINVAL, msg); \
  } while (0)

inline cudaStream_t as_stream(void *s) { return reinterpret_cast<cudaStream_t>(s); }

int sm_count();  // cached multiProcessorCount of the current device

inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }

// grid for a grid-stride kernel: enough CTAs to fill the machine `waves` times, never more than needed
inline unsigned persistent_grid(int64_t work_ctas, int ctas_per_sm) {
  int64_t cap = (int64_t)sm_count() * ctas_per_sm;
  int64_t g = work_ctas < cap ? work_ctas : cap;
  return (unsigned)(g < 1 ? 1 : g);
}

// ---- device helpers -------------------------------------------------------
__device__ __forceinline__ int lane_id() { return threadIdx.x & 31; }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ long long warp_sum(long long v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ int warp_max(int v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = max(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// streaming (evict-first) 16-byte store: outputs that are never re-read by this kernel
__device__ __forceinline__ void st_cs(float4 *p, float4 v) {
  asm volatile("st.global.cs.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w)
               : "memory");
}
// read-only 16-byte gather that does not pollute L1 (random feature rows have no L1 reuse)
__device__ __forceinline__ float4 ldg_na(const float4 *p) {
  float4 r;
  asm("ld.global.nc.L1::no_allocate.v4.f32 {%0, %1, %2, %3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
               : "l"(p));
  return r;
}

// plan layout (int64 words): see wdgh_plan_build
constexpr int kPlanNHeavy = 0;      // number of split rows
constexpr int kPlanNChunks = 1;     // number of chunks over all split rows
constexpr int kPlanThreshold = 2;   // heavy threshold (entries)
constexpr int kPlanCapacity = 3;    // capacity (chunks) of the three arrays below
// Scheduling words of the persistent row-group kernels: a 64-bit ticket counter (groups handed out so far) and a
// count of finished CTAs.  Zeroed by wdgh_plan_build; the last CTA of a launch zeroes both again, so no memset is
// needed between launches.  One pair per kernel family: a plan must not be used by two launches that run
// concurrently (they would also share the split-row `partial` scratch).
constexpr int kPlanSpmmTicket = 8;   // [8] tickets, [9] finished CTAs of spmm_rowgroup_kernel
constexpr int kPlanLabelTicket = 10; // [10] tickets, [11] finished CTAs of the label edge pass
// arrays after the header, each `capacity` long:
//   heavy_row[k]      row id of split row k
//   heavy_chunk0[k]   first chunk id of split row k
//   chunk_owner[c]    split-row index k that owns chunk c
__host__ __device__ inline const int64_t *plan_heavy_row(const int64_t *p) { return p + WDGH_PLAN_HEADER; }
__host__ __device__ inline const int64_t *plan_heavy_chunk0(const int64_t *p, int64_t cap) {
  return p + WDGH_PLAN_HEADER + cap;
}
__host__ __device__ inline const int64_t *plan_chunk_owner(const int64_t *p, int64_t cap) {
  return p + WDGH_PLAN_HEADER + 2 * cap;
}
inline unsigned long long *plan_sched(int64_t *plan, int word) {
  return reinterpret_cast<unsigned long long *>(plan + word);
}

// Ticket dispenser of a persistent row-group kernel.  `sched[0]` counts tickets, `sched[1]` finished CTAs.
// retire(): every CTA calls it once after its last ticket; the CTA that finishes last re-arms both words.
__device__ __forceinline__ void sched_retire(unsigned long long *sched, unsigned total_ctas) {
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned long long done = atomicAdd(&sched[1], 1ull);
    if (done == (unsigned long long)total_ctas - 1ull) {
      atomicExch(&sched[0], 0ull);
      atomicExch(&sched[1], 0ull);
    }
  }
}

}  // namespace wdgh
