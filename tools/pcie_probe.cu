// Microbenchmark (not product code): what the PCIe link of this box gives the e2e entry (wdgh_pipeline_host).
//   * contiguous H2D / D2H copies, alone and concurrently (is the link full duplex in practice?)
//   * pitched (2-D) copies of a column block of a row-major [n][128] float matrix: width 64 .. 256 bytes out of a
//     512-byte pitch, copy engines (cudaMemcpy2DAsync) vs a kernel that reads / writes the pinned host buffer
//     directly (zero copy)
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/pcie_probe tools/pcie_probe.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1);} } while (0)

// copy the column block [c0, c0 + w4) (in float4 units) of every row: src / dst pitches in float4 units
__global__ void __launch_bounds__(256) colblock_copy(const float4 *__restrict__ src, int64_t sp, float4 *__restrict__ dst,
                                                     int64_t dp, int64_t n, int w4) {
  const int64_t total = n * w4;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / w4;
    const int c = (int)(i - r * w4);
    dst[r * dp + c] = src[r * sp + c];
  }
}

static float ms_between(cudaEvent_t a, cudaEvent_t b) {
  float ms = 0;
  CK(cudaEventElapsedTime(&ms, a, b));
  return ms;
}

int main(int argc, char **argv) {
  const int64_t n = argc > 1 ? atoll(argv[1]) : (8 << 20);   // rows of 512 bytes
  const size_t bytes = (size_t)n * 512;
  float *hx, *hy, *dx, *dy;
  CK(cudaHostAlloc(&hx, bytes, cudaHostAllocDefault));
  CK(cudaHostAlloc(&hy, bytes, cudaHostAllocDefault));
  CK(cudaMalloc(&dx, bytes));
  CK(cudaMalloc(&dy, bytes));
  for (size_t i = 0; i < bytes / 4; i += 1024) hx[i] = (float)i;
  cudaStream_t s0, s1;
  CK(cudaStreamCreateWithFlags(&s0, cudaStreamNonBlocking));
  CK(cudaStreamCreateWithFlags(&s1, cudaStreamNonBlocking));
  cudaEvent_t a0, b0, a1, b1;
  CK(cudaEventCreate(&a0)); CK(cudaEventCreate(&b0)); CK(cudaEventCreate(&a1)); CK(cudaEventCreate(&b1));
  const double gb = bytes / 1e9;

  // --- contiguous ---
  for (int rep = 0; rep < 2; ++rep) {
    CK(cudaEventRecord(a0, s0));
    CK(cudaMemcpyAsync(dx, hx, bytes, cudaMemcpyHostToDevice, s0));
    CK(cudaEventRecord(b0, s0));
    CK(cudaStreamSynchronize(s0));
    if (rep) printf("contiguous H2D           : %6.1f GB/s\n", gb / ms_between(a0, b0) * 1e3);
    CK(cudaEventRecord(a1, s1));
    CK(cudaMemcpyAsync(hy, dy, bytes, cudaMemcpyDeviceToHost, s1));
    CK(cudaEventRecord(b1, s1));
    CK(cudaStreamSynchronize(s1));
    if (rep) printf("contiguous D2H           : %6.1f GB/s\n", gb / ms_between(a1, b1) * 1e3);
  }
  CK(cudaEventRecord(a0, s0));
  CK(cudaMemcpyAsync(dx, hx, bytes, cudaMemcpyHostToDevice, s0));
  CK(cudaEventRecord(b0, s0));
  CK(cudaEventRecord(a1, s1));
  CK(cudaMemcpyAsync(hy, dy, bytes, cudaMemcpyDeviceToHost, s1));
  CK(cudaEventRecord(b1, s1));
  CK(cudaDeviceSynchronize());
  printf("concurrent H2D + D2H     : %6.1f + %6.1f GB/s\n", gb / ms_between(a0, b0) * 1e3, gb / ms_between(a1, b1) * 1e3);

  // --- pitched column blocks, copy engines ---
  for (int w = 64; w <= 256; w *= 2) {
    const double g2 = (double)n * w / 1e9;
    CK(cudaEventRecord(a0, s0));
    CK(cudaMemcpy2DAsync(dx, 512, hx, 512, w, n, cudaMemcpyHostToDevice, s0));
    CK(cudaEventRecord(b0, s0));
    CK(cudaStreamSynchronize(s0));
    printf("2-D CE H2D  width %3d B   : %6.1f GB/s\n", w, g2 / ms_between(a0, b0) * 1e3);
    CK(cudaEventRecord(a0, s0));
    CK(cudaMemcpy2DAsync(dx, w, hx, 512, w, n, cudaMemcpyHostToDevice, s0));   // packed on the device side
    CK(cudaEventRecord(b0, s0));
    CK(cudaStreamSynchronize(s0));
    printf("2-D CE H2D  width %3d B -> packed : %6.1f GB/s\n", w, g2 / ms_between(a0, b0) * 1e3);
    CK(cudaEventRecord(a1, s1));
    CK(cudaMemcpy2DAsync(hy, 512, dy, 512, w, n, cudaMemcpyDeviceToHost, s1));
    CK(cudaEventRecord(b1, s1));
    CK(cudaStreamSynchronize(s1));
    printf("2-D CE D2H  width %3d B   : %6.1f GB/s\n", w, g2 / ms_between(a1, b1) * 1e3);
    CK(cudaEventRecord(a0, s0));
    CK(cudaMemcpy2DAsync(dx, 512, hx, 512, w, n, cudaMemcpyHostToDevice, s0));
    CK(cudaEventRecord(b0, s0));
    CK(cudaEventRecord(a1, s1));
    CK(cudaMemcpy2DAsync(hy, 512, dy, 512, w, n, cudaMemcpyDeviceToHost, s1));
    CK(cudaEventRecord(b1, s1));
    CK(cudaDeviceSynchronize());
    printf("2-D CE both width %3d B   : %6.1f + %6.1f GB/s\n", w, g2 / ms_between(a0, b0) * 1e3, g2 / ms_between(a1, b1) * 1e3);
  }

  // --- zero copy: kernels touching the pinned host buffers ---
  float4 *hx4 = (float4 *)hx, *hy4 = (float4 *)hy, *dx4 = (float4 *)dx, *dy4 = (float4 *)dy;
  for (int ctas = 148; ctas <= 148 * 8; ctas *= 2) {
    for (int w = 64; w <= 512; w *= 2) {
      const double g2 = (double)n * w / 1e9;
      const int w4 = w / 16;
      CK(cudaEventRecord(a0, s0));
      colblock_copy<<<ctas, 256, 0, s0>>>(hx4, 32, dx4, 32, n, w4);
      CK(cudaEventRecord(b0, s0));
      CK(cudaStreamSynchronize(s0));
      const float t_r = ms_between(a0, b0);
      CK(cudaEventRecord(a1, s1));
      colblock_copy<<<ctas, 256, 0, s1>>>(dy4, 32, hy4, 32, n, w4);
      CK(cudaEventRecord(b1, s1));
      CK(cudaStreamSynchronize(s1));
      const float t_w = ms_between(a1, b1);
      CK(cudaEventRecord(a0, s0));
      colblock_copy<<<ctas, 256, 0, s0>>>(hx4, 32, dx4, 32, n, w4);
      CK(cudaEventRecord(b0, s0));
      CK(cudaEventRecord(a1, s1));
      colblock_copy<<<ctas, 256, 0, s1>>>(dy4, 32, hy4, 32, n, w4);
      CK(cudaEventRecord(b1, s1));
      CK(cudaDeviceSynchronize());
      printf("zero copy %4d CTAs width %3d B : read %6.1f  write %6.1f  both %6.1f + %6.1f GB/s\n", ctas, w,
             g2 / t_r * 1e3, g2 / t_w * 1e3, g2 / ms_between(a0, b0) * 1e3, g2 / ms_between(a1, b1) * 1e3);
    }
  }
  CK(cudaGetLastError());
  return 0;
}
