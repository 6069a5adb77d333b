// Microbenchmark (not product code): ceiling of random 512-byte row gathers on this GPU, the access
// pattern of the d=128 SpMM.  Variants: register path (U independent LDG.128 per lane) and cp.async
// staging into shared memory (in-flight bytes not limited by registers).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/gather_bw tools/gather_bw.cu
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <stdint.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1);} } while (0)

__device__ __forceinline__ float4 ldg_na(const float4 *p) {
  float4 r;
  asm("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
  return r;
}

__global__ void fill_idx(int *idx, long m, int n, unsigned seed) {
  long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  long stride = (long)gridDim.x * blockDim.x;
  for (; i < m; i += stride) {
    unsigned long long z = (i + 1) * 0x9E3779B97F4A7C15ull + seed;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull; z = (z ^ (z >> 27)) * 0x94D049BB133111EBull; z ^= z >> 31;
    idx[i] = (int)(z % (unsigned long long)n);
  }
}

// register path: each warp walks its share of idx in batches of 32, U gathers in flight
template <int U>
__global__ void gather_reg(const float4 *__restrict__ x, const int *__restrict__ idx, long m, float4 *out) {
  const int lane = threadIdx.x & 31;
  const long warp = ((long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long nwarps = ((long)gridDim.x * blockDim.x) >> 5;
  float4 acc = make_float4(0, 0, 0, 0);
  for (long base = warp * 32; base < m; base += nwarps * 32) {
    int j = idx[base + lane];
    for (int k = 0; k < 32; k += U) {
      float4 v[U];
#pragma unroll
      for (int u = 0; u < U; ++u) v[u] = ldg_na(x + (long)__shfl_sync(0xffffffffu, j, k + u) * 32 + lane);
#pragma unroll
      for (int u = 0; u < U; ++u) { acc.x += v[u].x; acc.y += v[u].y; acc.z += v[u].z; acc.w += v[u].w; }
    }
  }
  if (acc.x == 12345.678f) out[warp * 32 + lane] = acc;
}

// register path + a second dependent random gather per entry (the D^-1/2 scale of the SpMM):
// MODE 1: 4-byte scale from an n*4-byte array (200 MB at n=50M: misses L2);  MODE 2: 1-byte code from an
// n-byte array (L2-resident) + 256-entry table
template <int U, int MODE>
__global__ void gather_reg_scale(const float4 *__restrict__ x, const int *__restrict__ idx, long m,
                                 const float *__restrict__ sc4, const unsigned char *__restrict__ sc1,
                                 const float *__restrict__ table, float4 *out) {
  const int lane = threadIdx.x & 31;
  const long warp = ((long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long nwarps = ((long)gridDim.x * blockDim.x) >> 5;
  float4 acc = make_float4(0, 0, 0, 0);
  for (long base = warp * 32; base < m; base += nwarps * 32) {
    int j = idx[base + lane];
    float w = (MODE == 1) ? __ldg(sc4 + j) : __ldg(table + __ldg(sc1 + j));
    for (int k = 0; k < 32; k += U) {
      float4 v[U];
#pragma unroll
      for (int u = 0; u < U; ++u) v[u] = ldg_na(x + (long)__shfl_sync(0xffffffffu, j, k + u) * 32 + lane);
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const float wu = __shfl_sync(0xffffffffu, w, k + u);
        acc.x += wu * v[u].x; acc.y += wu * v[u].y; acc.z += wu * v[u].z; acc.w += wu * v[u].w;
      }
    }
  }
  if (acc.x == 12345.678f) out[warp * 32 + lane] = acc;
}

// cp.async path: per warp STAGES x B slots of 512 B; stage s+1 is issued before stage s is consumed
template <int B, int STAGES>
__global__ void gather_cpasync(const float4 *__restrict__ x, const int *__restrict__ idx, long m, float4 *out) {
  extern __shared__ float4 smem[];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  float4 *ring = smem + (size_t)wid * STAGES * B * 32;
  const long warp = ((long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long nwarps = ((long)gridDim.x * blockDim.x) >> 5;
  float4 acc = make_float4(0, 0, 0, 0);
  auto issue = [&](long base, int stage) {
    int j = (lane < B && base + lane < m) ? idx[base + lane] : 0;
#pragma unroll
    for (int k = 0; k < B; ++k) {
      const float4 *src = x + (long)__shfl_sync(0xffffffffu, j, k) * 32 + lane;
      unsigned dst = (unsigned)__cvta_generic_to_shared(ring + (stage * B + k) * 32 + lane);
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src));
    }
    asm volatile("cp.async.commit_group;");
  };
  long base = warp * B;
  const long step = nwarps * B;
  int st = 0;
#pragma unroll
  for (int s = 0; s < STAGES - 1; ++s) { issue(base + s * step, s); }
  for (; base < m; base += step) {
    issue(base + (STAGES - 1) * step, (st + STAGES - 1) % STAGES);
    asm volatile("cp.async.wait_group %0;" ::"n"(STAGES - 1));
#pragma unroll
    for (int k = 0; k < B; ++k) {
      float4 v = ring[(st * B + k) * 32 + lane];
      acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
    st = (st + 1) % STAGES;
  }
  asm volatile("cp.async.wait_group 0;");
  if (acc.x == 12345.678f) out[warp * 32 + lane] = acc;
}

// gathers + one 512-byte output row per K gathers (the SpMM's read/write mix at full memory-level parallelism:
// batches of U are cut from the entry stream of 32 consecutive output rows, whatever K is).
// STORE: 0 = st.global, 1 = st.global.cs (streaming), 2 = no store (the K-row sums are discarded)
template <int U, int STORE>
__global__ void gather_write(const float4 *__restrict__ x, const int *__restrict__ idx, long rows, int K, float4 *y) {
  const int lane = threadIdx.x & 31;
  const long warp = ((long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long nwarps = ((long)gridDim.x * blockDim.x) >> 5;
  const long groups = rows / 32;
  float4 acc = make_float4(0, 0, 0, 0);
  float4 sink = make_float4(0, 0, 0, 0);
  for (long g = warp; g < groups; g += nwarps) {
    const long e0 = g * 32 * K, total = 32L * K;
    long row = g * 32;
    int in_row = 0;
    for (long t0 = 0; t0 < total; t0 += 32) {
      const int j = idx[e0 + t0 + lane];
      for (int k = 0; k < 32; k += U) {
        float4 v[U];
#pragma unroll
        for (int u = 0; u < U; ++u) v[u] = ldg_na(x + (long)__shfl_sync(0xffffffffu, j, k + u) * 32 + lane);
#pragma unroll
        for (int u = 0; u < U; ++u) {
          acc.x += v[u].x; acc.y += v[u].y; acc.z += v[u].z; acc.w += v[u].w;
          if (++in_row == K) {
            float4 *dst = y + row * 32 + lane;
            if (STORE == 0) *dst = acc;
            else if (STORE == 1) asm volatile("st.global.cs.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(dst), "f"(acc.x), "f"(acc.y), "f"(acc.z), "f"(acc.w) : "memory");
            else { sink.x += acc.x; }
            acc = make_float4(0, 0, 0, 0);
            in_row = 0;
            ++row;
          }
        }
      }
    }
  }
  if (sink.x == 12345.678f) y[lane] = sink;
}

template <typename F>
static float time_ms(F f, int reps) {
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  f(); cudaDeviceSynchronize();
  cudaEventRecord(a);
  for (int i = 0; i < reps; ++i) f();
  cudaEventRecord(b); cudaEventSynchronize(b);
  float ms; cudaEventElapsedTime(&ms, a, b);
  return ms / reps;
}

int main(int argc, char **argv) {
  const int n = argc > 1 ? atoi(argv[1]) : 16000000;   // rows of 512 B
  const long m = argc > 2 ? atol(argv[2]) : 256000000; // gathers
  const bool rw_only = argc > 3;                       // any third argument: only the read/write-mix cases
  float4 *x, *out; int *idx;
  CK(cudaMalloc(&x, (size_t)n * 512)); CK(cudaMemset(x, 0, (size_t)n * 512));
  CK(cudaMalloc(&out, 1 << 24)); CK(cudaMalloc(&idx, m * 4 + 4096));
  fill_idx<<<4096, 256>>>(idx, m + 1024, n, 7u); CK(cudaDeviceSynchronize());
  int sms = 0; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  const double bytes = (double)m * 516;
  printf("rows=%d (%.1f GB) gathers=%ld sms=%d\n", n, n * 512.0 / 1e9, m, sms);
#define RUN_REG(U, BLOCK, PER_SM) { float ms = time_ms([&] { gather_reg<U><<<sms * PER_SM, BLOCK>>>(x, idx, m, out); }, 3); \
    CK(cudaGetLastError()); printf("reg  U=%2d block=%3d ctas/sm=%2d : %7.2f ms %7.1f GB/s\n", U, BLOCK, PER_SM, ms, bytes / ms / 1e6); }
  {
    // read/write mix: K gathers per 512 B output row (K = 20: the 1-D shard, K = 5: a row slice of the 2-D partition)
    float4 *y;
    const int Ks[3] = {20, 5, 2};
    CK(cudaMalloc(&y, (size_t)(m / 2) * 512 + 4096));
    for (int K : Ks) {
      const long rows = (m / K) / 32 * 32;
      const double rd = (double)rows * K * 516, wr = (double)rows * 512;
      float ms = time_ms([&] { gather_write<8, 2><<<sms * 32, 32>>>(x, idx, rows, K, y); }, 3);
      printf("K=%2d gathers/row, no store      : %7.2f ms %7.1f GB/s\n", K, ms, rd / ms / 1e6);
      ms = time_ms([&] { gather_write<8, 0><<<sms * 32, 32>>>(x, idx, rows, K, y); }, 3);
      printf("K=%2d gathers/row, st.global     : %7.2f ms %7.1f GB/s (reads + writes)\n", K, ms, (rd + wr) / ms / 1e6);
      ms = time_ms([&] { gather_write<8, 1><<<sms * 32, 32>>>(x, idx, rows, K, y); }, 3);
      printf("K=%2d gathers/row, st.global.cs  : %7.2f ms %7.1f GB/s (reads + writes)\n", K, ms, (rd + wr) / ms / 1e6);
    }
    CK(cudaFree(y));
    if (rw_only) return 0;
  }
  RUN_REG(8, 32, 32) RUN_REG(8, 128, 8) RUN_REG(16, 32, 20) RUN_REG(16, 32, 32) RUN_REG(4, 32, 32) RUN_REG(32, 32, 16)
  {
    float *sc4; unsigned char *sc1; float *table;
    CK(cudaMalloc(&sc4, (size_t)n * 4)); CK(cudaMemset(sc4, 0, (size_t)n * 4));
    CK(cudaMalloc(&sc1, (size_t)n)); CK(cudaMemset(sc1, 1, (size_t)n));
    CK(cudaMalloc(&table, 1024)); CK(cudaMemset(table, 0, 1024));
    float ms = time_ms([&] { gather_reg_scale<8, 1><<<sms * 32, 32>>>(x, idx, m, sc4, sc1, table, out); }, 3);
    printf("reg+scale f32 gather (%.0f MB array): %7.2f ms %7.1f GB/s (row bytes only)\n", n * 4.0 / 1e6, ms, bytes / ms / 1e6);
    ms = time_ms([&] { gather_reg_scale<8, 2><<<sms * 32, 32>>>(x, idx, m, sc4, sc1, table, out); }, 3);
    printf("reg+scale u8 code + table (%.0f MB array): %7.2f ms %7.1f GB/s (row bytes only)\n", n * 1.0 / 1e6, ms, bytes / ms / 1e6);
  }
#define RUN_CPA(B, ST, WARPS, PER_SM) { size_t sh = (size_t)WARPS * ST * B * 512; \
    CK(cudaFuncSetAttribute(gather_cpasync<B, ST>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sh)); \
    float ms = time_ms([&] { gather_cpasync<B, ST><<<sms * PER_SM, WARPS * 32, sh>>>(x, idx, m, out); }, 3); \
    CK(cudaGetLastError()); printf("cpas B=%2d stages=%d warps/cta=%2d ctas/sm=%d smem/sm=%3zu KB : %7.2f ms %7.1f GB/s\n", B, ST, WARPS, PER_SM, sh * PER_SM / 1024, ms, bytes / ms / 1e6); }
  RUN_CPA(16, 2, 8, 1) RUN_CPA(16, 2, 12, 1) RUN_CPA(16, 3, 8, 1) RUN_CPA(32, 2, 6, 1) RUN_CPA(8, 2, 24, 1) RUN_CPA(8, 3, 16, 1)
  RUN_CPA(16, 2, 4, 3) RUN_CPA(8, 4, 12, 1) RUN_CPA(32, 3, 4, 1) RUN_CPA(16, 4, 6, 1)
  return 0;
}
