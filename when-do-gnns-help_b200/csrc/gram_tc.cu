// Gram matrix G = Z Z^T on the 5th-generation tensor cores (tcgen05 + TMEM), fed by TMA.
//
// Replaces the dense contractions of utils/homophily_metrics.py:
//   inner_prod = (A X)(A X)^T        :192, :199-200   (aggregation similarity)
//   G_gram / gram                    :234-235, :246   (GNTK / kernel-regression Gram)
// The reference computes them in float32 (torch.mm).  tcgen05 has no fp32 input kind, so each
// operand is split z = hi + lo (hi = top 19 bits, i.e. exactly a TF32 number; lo = z - hi, exact
// in fp32) and the product is evaluated as  hi hi^T + hi lo^T + lo hi^T  ("3xTF32", error ~2^-19
// relative per product, fp32 accumulation in TMEM).  The three terms are ONE GEMM over a
// concatenated K axis:   A' = [hi | hi | lo],  B' = [hi | lo | hi],  G = A' B'^T,  K' = 3 * dpad.
//
// Kernel shape (one 128x128 output tile per CTA, upper triangle only, mirrored in the epilogue):
//   warp 0 : TMA producer   -- cp.async.bulk.tensor.2d, 128 rows x 32 fp32 (128 B, SWIZZLE_128B) per operand/stage
//   warp 1 : MMA issuer     -- one elected thread, tcgen05.mma.cta_group::1.kind::tf32, M=128 N=128 K=8
//   warps 2-5 : epilogue    -- tcgen05.ld 32x32b.x32 (TMEM -> registers) -> global stores (+ transposed copy)
//   4-stage shared-memory ring, full/empty mbarriers, accumulator handed over with tcgen05.commit.
// The tensor core adds into its fp32 accumulator with truncation, so the error of one accumulator
// grows with (number of k-steps) x (running magnitude).  k-blocks are therefore dealt round-robin to
// kAccums = 4 accumulators (all 512 TMEM columns) that the epilogue adds in registers with proper
// rounding: 4x smaller accumulation error (measured 4.8e-5 -> 1.2e-5 relative at K = 1433).
// Short K (< 16 k-blocks) uses one accumulator (128 TMEM columns) and as many stages as k-blocks, so
// that several CTAs fit an SM and the output-bound small-K case (K = number of classes) overlaps
// the epilogue of one tile with the loads of the next.
#include <cuda.h>

#include "common.cuh"

namespace wdgh {

constexpr int kTileM = 128, kTileN = 128, kBlockK = 32;  // 32 fp32 = one 128-byte swizzle row
constexpr int kMaxStages = 4;
constexpr int kUmmaK = 8;                                 // tf32: 32 bytes of K per instruction
constexpr uint32_t kStageBytesA = kTileM * kBlockK * 4;   // 16 KB
constexpr uint32_t kStageBytesB = kTileN * kBlockK * 4;   // 16 KB
constexpr int kMaxAccums = 4;                             // round-robin accumulators (see below)
constexpr int kGramThreads = 192;                         // 6 warps
static size_t gram_smem_bytes(int stages) { return 1024 /*align slack*/ + (size_t)stages * (kStageBytesA + kStageBytesB) + 256; }

// ---- PTX wrappers ------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred P1;\n\t"
      "WAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
      "@P1 bra DONE;\n\t"
      "bra WAIT_LOOP;\n\t"
      "DONE:\n\t"
      "}" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d(void *dst, const CUtensorMap *map, uint64_t *bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
          smem_u32(dst)),
      "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
// K-major operand tile, 128-byte swizzle: 8-row groups 1024 B apart (SBO), LBO unused (=1), version 1
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t smem_addr) {
  return (uint64_t)((smem_addr & 0x3FFFF) >> 4) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}
// D=F32 (1<<4), A=B=TF32 (2<<7, 2<<10), both K-major, N>>3 at bit 17, M>>4 at bit 24
constexpr uint32_t kInstrDesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(kTileN >> 3) << 17) |
                                ((uint32_t)(kTileM >> 4) << 24);

__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(tmem_d),
      "l"(da), "l"(db), "r"(kInstrDesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t *bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, "
      "%24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

// ---- operand preparation: A' = [hi | hi | lo], B' = [hi | lo | hi], zero padded to dpad ---------
__global__ void gram_split_kernel(const float *__restrict__ z, int64_t m, int64_t d, int64_t ldz, int64_t dpad,
                                  float *__restrict__ a, float *__restrict__ b) {
  const int64_t total = m * dpad;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const int64_t ld = 3 * dpad;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += stride) {
    const int64_t r = t / dpad, k = t - r * dpad;
    float hi = 0.f, lo = 0.f;
    if (k < d) {
      const float v = z[r * ldz + k];
      hi = __uint_as_float(__float_as_uint(v) & 0xFFFFE000u);  // exactly representable in TF32
      lo = v - hi;                                             // exact; the tensor core keeps its top 11 bits
    }
    float *ar = a + r * ld, *br = b + r * ld;
    ar[k] = hi; ar[dpad + k] = hi; ar[2 * dpad + k] = lo;
    br[k] = hi; br[dpad + k] = lo; br[2 * dpad + k] = hi;
  }
}

// ---- the tensor-core kernel ----------------------------------------------------------------------
__global__ void __launch_bounds__(kGramThreads, 2)
gram_tcgen05_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b, int64_t m,
                    int num_k_blocks, int kStages, int kAccums, float *__restrict__ g, int64_t ldg) {
  const int tj = blockIdx.x, ti = blockIdx.y;
  if (tj < ti) return;  // G is symmetric: lower tiles are written by the mirrored store of the upper ones

  extern __shared__ uint8_t smem_raw[];
  uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t *smem_a = smem;                                   // kStages x 16 KB, 1024-byte aligned
  uint8_t *smem_b = smem + kStages * kStageBytesA;
  uint64_t *full_bar = reinterpret_cast<uint64_t *>(smem + kStages * (kStageBytesA + kStageBytesB));  // 9 barriers + slot < 256 B
  uint64_t *empty_bar = full_bar + kMaxStages;
  uint64_t *accum_bar = empty_bar + kMaxStages;
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(accum_bar + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == 1 && lane == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(accum_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  const uint32_t kTmemCols = 128u * (uint32_t)kAccums;  // 128 or 512: a power of two >= 32
  if (warp == 2) {  // one warp allocates the accumulator columns and later frees them
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"(kTmemCols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_acc = *tmem_slot;

  if (warp == 0) {
    // ===== TMA producer =====
    if (lane == 0) {
      for (int kb = 0; kb < num_k_blocks; ++kb) {
        const int s = kb % kStages;
        const uint32_t phase = (kb / kStages) & 1;
        mbar_wait(&empty_bar[s], phase ^ 1);  // slot free (first pass returns immediately)
        mbar_expect_tx(&full_bar[s], kStageBytesA + kStageBytesB);
        tma_load_2d(smem_a + s * kStageBytesA, &map_a, &full_bar[s], kb * kBlockK, ti * kTileM);
        tma_load_2d(smem_b + s * kStageBytesB, &map_b, &full_bar[s], kb * kBlockK, tj * kTileN);
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer (single thread) =====
    if (lane == 0) {
      for (int kb = 0; kb < num_k_blocks; ++kb) {
        const int s = kb % kStages;
        const uint32_t phase = (kb / kStages) & 1;
        mbar_wait(&full_bar[s], phase);  // TMA bytes have landed
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint64_t da = umma_desc_sw128(smem_u32(smem_a + s * kStageBytesA));
        const uint64_t db = umma_desc_sw128(smem_u32(smem_b + s * kStageBytesB));
#pragma unroll
        for (int k = 0; k < kBlockK / kUmmaK; ++k) {
          // advance 32 bytes along K inside the 128-byte swizzle row: +2 in the (addr >> 4) field
          umma_tf32(tmem_acc + (uint32_t)((kb % kAccums) * kTileN), da + (uint64_t)(2 * k), db + (uint64_t)(2 * k),
                    (kb >= kAccums || k > 0) ? 1u : 0u);
        }
        umma_commit(&empty_bar[s]);  // frees the smem slot once these MMAs have read it
      }
      umma_commit(accum_bar);        // accumulator complete
    }
  } else {
    // ===== epilogue: warps 2..5; a warp may only touch TMEM lanes 32*(warp%4) .. +31 =====
    const int q = warp & 3;
    mbar_wait(accum_bar, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const int64_t i = (int64_t)ti * kTileM + q * 32 + lane;  // output row of this thread
#pragma unroll 1
    const int used = num_k_blocks < kAccums ? num_k_blocks : kAccums;  // accumulators that were written
    for (int c0 = 0; c0 < kTileN; c0 += 32) {
      float v[32];
      tmem_ld32(tmem_acc + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, v);
      for (int a = 1; a < used; ++a) {
        float t[32];
        tmem_ld32(tmem_acc + ((uint32_t)(q * 32) << 16) + (uint32_t)(a * kTileN + c0), t);
#pragma unroll
        for (int c = 0; c < 32; ++c) v[c] += t[c];
      }
      const int64_t j0 = (int64_t)tj * kTileN + c0;
      if (i < m) {
        float *dst = g + i * ldg + j0;
        if (j0 + 32 <= m && (reinterpret_cast<uintptr_t>(dst) & 15) == 0) {  // 8 x 16-byte stores: whole sectors
#pragma unroll
          for (int c = 0; c < 32; c += 4)
            *reinterpret_cast<float4 *>(dst + c) = make_float4(v[c], v[c + 1], v[c + 2], v[c + 3]);
        } else {
#pragma unroll
          for (int c = 0; c < 32; ++c)
            if (j0 + c < m) dst[c] = v[c];
        }
      }
      if (ti != tj && i < m) {  // mirrored tile: lanes hold consecutive i -> coalesced
#pragma unroll
        for (int c = 0; c < 32; ++c)
          if (j0 + c < m) g[(j0 + c) * ldg + i] = v[c];
      }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  }
  __syncthreads();
  if (warp == 2) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_acc), "r"(kTmemCols) : "memory");
  }
}

// ---- host side -------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (fn == nullptr) {
    void *p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

// 2-D fp32 tensor [rows][kp] (kp contiguous), box = 32 (K) x 128 (rows), 128-byte swizzle, OOB rows read as 0
static int make_map(CUtensorMap *map, float *base, int64_t rows, int64_t kp) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) return fail(WDGH_ENODEV, "wdgh_gram: cuTensorMapEncodeTiled is not available from the driver");
  cuuint64_t dims[2] = {(cuuint64_t)kp, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)kp * sizeof(float)};
  cuuint32_t box[2] = {(cuuint32_t)kBlockK, (cuuint32_t)kTileM};
  cuuint32_t elem[2] = {1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, base, dims, strides, box, elem, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    snprintf(g_err, sizeof(g_err), "wdgh_gram: cuTensorMapEncodeTiled failed (CUresult %d)", (int)r);
    return WDGH_EINVAL;
  }
  return 0;
}

}  // namespace wdgh

using namespace wdgh;

extern "C" int64_t wdgh_gram_workspace_floats(int64_t m, int64_t d) {
  const int64_t dpad = ceil_div(d, kBlockK) * kBlockK;
  return 2 * m * 3 * dpad;
}

int wdgh_gram_tc_launch(const float *z, int64_t m, int64_t d, int64_t ldz, float *g, int64_t ldg, float *workspace,
                        cudaStream_t st) {
  if (workspace == nullptr) return fail(WDGH_EINVAL, "wdgh_gram: the tensor-core path needs the workspace");
  if (reinterpret_cast<uintptr_t>(workspace) % 16 != 0) return fail(WDGH_EINVAL, "wdgh_gram: workspace must be 16-byte aligned");
  const int64_t dpad = ceil_div(d, kBlockK) * kBlockK;
  const int64_t kp = 3 * dpad;
  float *a = workspace, *b = workspace + m * kp;
  gram_split_kernel<<<persistent_grid(ceil_div(m * dpad, 256), 8), 256, 0, st>>>(z, m, d, ldz, dpad, a, b);
  WDGH_LAUNCHED("gram_split_kernel");
  CUtensorMap map_a, map_b;
  int rc = make_map(&map_a, a, m, kp);
  if (rc) return rc;
  rc = make_map(&map_b, b, m, kp);
  if (rc) return rc;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(gram_tcgen05_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)gram_smem_bytes(kMaxStages));
    if (e != cudaSuccess) return fail_cuda(e, "cudaFuncSetAttribute(gram_tcgen05_kernel)");
    configured = true;
  }
  const int num_k_blocks = (int)(kp / kBlockK);
  const int stages = num_k_blocks < kMaxStages ? num_k_blocks : kMaxStages;
  const int accums = num_k_blocks >= 16 ? kMaxAccums : 1;
  const unsigned nt = (unsigned)ceil_div(m, kTileM);
  gram_tcgen05_kernel<<<dim3(nt, nt), kGramThreads, gram_smem_bytes(stages), st>>>(map_a, map_b, m, num_k_blocks,
                                                                                   stages, accums, g, ldg);
  WDGH_LAUNCHED("gram_tcgen05_kernel");
  return 0;
}
