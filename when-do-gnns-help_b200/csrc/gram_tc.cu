// Gram matrix G = Z Z^T on the 5th-generation tensor cores (tcgen05 + TMEM), fed by TMA.
//
// Replaces the dense contractions of utils/homophily_metrics.py:
//   inner_prod = (A X)(A X)^T        :192, :199-200   (aggregation similarity)
//   G_gram / gram                    :234-235, :246   (GNTK / kernel-regression Gram)
// The reference computes them in float32 (torch.mm).  tcgen05 has no fp32 input kind, so each operand is split
// z = hi + lo with hi = RN_tf32(z) and lo = RN_tf32(z - hi) (residual <= 2^-24 |z|: fp32 precision) and the product
// is evaluated as  hi hi^T + hi lo^T + lo hi^T  ("3xTF32"; the dropped lo lo^T term is ~2^-24 relative).
//
// What bounds the kernel is L2 -> shared-memory operand traffic, not the tensor pipe (round 1: 36% pipe-active with
// 128x128 tiles and a 6x expanded operand workspace, i.e. 128 B of operands per MMA clock against the ~42 B/clk/SM the
// L2 delivers chip-wide).  This version therefore
//   * keeps ONE hi and ONE lo copy of Z (2x workspace instead of 6x) and stages four tiles per k-block --
//     A_hi, A_lo (128 rows) and B_hi, B_lo (256 rows) -- from which the three products are issued;
//   * computes 128 x 256 output tiles (M = 128, N = 256 per instruction): 96 KB of operands per 1536 MMA clocks;
//   * runs persistent CTAs (one per SM) over the upper-triangle tiles only, consecutive tiles sharing the B rows.
// Accumulation: the tensor core adds into its fp32 TMEM accumulator with truncation, so the error of one accumulator
// grows with the number of instructions added into it.  The k-blocks are therefore cut into CHUNKS; each chunk gets a
// fresh TMEM accumulator (two, ping-pong) that the epilogue warps drain with tcgen05.ld and add into fp32 REGISTERS
// with round-to-nearest while the tensor core works on the next chunk.  chunk = 1 k-block (12 instructions) is
// fp32-faithful -- what the KR metric needs, whose pinv(rcond=1e-15) amplifies Gram noise -- chunk = 4 is the fast
// form for the similarity scores.
//
// Warp roles (320 threads): warp 0 = TMA producer, warp 1 = MMA issuer (one elected thread), warps 2..9 = chunk
// drain + epilogue (a warp may only touch TMEM lanes 32*(warp%4)..+31; two warps share a lane quarter, 128 columns
// each).  2-stage shared-memory ring (full/empty mbarriers), 2 TMEM accumulators (acc_full/acc_empty mbarriers).
#include <cuda.h>

#include "common.cuh"

namespace wdgh {

constexpr int kTileM = 128, kTileN = 256, kBlockK = 32;   // 32 fp32 = one 128-byte swizzle row
constexpr int kStages = 2;
constexpr int kUmmaK = 8;                                  // tf32: 32 bytes of K per instruction
constexpr uint32_t kBytesA = kTileM * kBlockK * 4;         // 16 KB per A tile (hi or lo)
constexpr uint32_t kBytesB = kTileN * kBlockK * 4;         // 32 KB per B tile (hi or lo)
constexpr uint32_t kStageBytes = 2 * kBytesA + 2 * kBytesB;  // 96 KB
constexpr int kAccums = 2;                                 // ping-pong TMEM accumulators of kTileN columns each
constexpr int kEpiWarps = 8;
constexpr int kGramThreads = 32 * (2 + kEpiWarps);         // 320
constexpr size_t kGramSmem = 1024 /*align slack*/ + (size_t)kStages * kStageBytes + 256;

// ---- PTX wrappers ------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
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

// ---- operand preparation: hi = RN_tf32(z), lo = RN_tf32(z - hi), zero padded to dpad columns ------
__global__ void gram_split_kernel(const float *__restrict__ z, int64_t m, int64_t d, int64_t ldz, int64_t dpad,
                                  float *__restrict__ hi, float *__restrict__ lo) {
  const int64_t total = m * dpad;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += stride) {
    const int64_t r = t / dpad, k = t - r * dpad;
    float h = 0.f, l = 0.f;
    if (k < d) {
      const float v = z[r * ldz + k];
      uint32_t hb, lb;
      asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(hb) : "f"(v));
      h = __uint_as_float(hb);
      if (isfinite(v)) {
        const float res = v - h;  // exact in fp32
        asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(lb) : "f"(res));
        l = __uint_as_float(lb);
      } else {
        h = v;  // inf / nan travel in the hi part only
      }
    }
    hi[t] = h;
    lo[t] = l;
  }
}

// tile t of the schedule -> (I, J, h): 256 x 256 super tile (I <= J, column-major over the upper triangle) and the
// 128-row half h of it.  Consecutive tiles share J, i.e. the B rows.
__device__ __forceinline__ void tile_coords(int t, int &I, int &J, int &h) {
  const int st = t >> 1;
  h = t & 1;
  int j = (int)((sqrtf(8.f * (float)st + 1.f) - 1.f) * 0.5f);
  while ((j + 1) * (j + 2) / 2 <= st) ++j;
  while (j * (j + 1) / 2 > st) --j;
  J = j;
  I = st - j * (j + 1) / 2;
}

// ---- the tensor-core kernel ----------------------------------------------------------------------
__global__ void __launch_bounds__(kGramThreads, 1)
gram_tcgen05_kernel(const __grid_constant__ CUtensorMap map_hi_a, const __grid_constant__ CUtensorMap map_lo_a,
                    const __grid_constant__ CUtensorMap map_hi_b, const __grid_constant__ CUtensorMap map_lo_b,
                    int64_t m, int num_k_blocks, int chunk, int num_tiles, float *__restrict__ g, int64_t ldg) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t *full_bar = reinterpret_cast<uint64_t *>(smem + kStages * kStageBytes);
  uint64_t *empty_bar = full_bar + kStages;
  uint64_t *acc_full = empty_bar + kStages;
  uint64_t *acc_empty = acc_full + kAccums;
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(acc_empty + kAccums);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == 1 && lane == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int a = 0; a < kAccums; ++a) {
      mbar_init(&acc_full[a], 1);
      mbar_init(&acc_empty[a], kEpiWarps);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  constexpr uint32_t kTmemCols = kAccums * kTileN;  // 512: the whole tensor memory of the SM (one CTA per SM)
  if (warp == 2) {  // one warp allocates the accumulator columns and later frees them
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "n"(kTmemCols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;
  const int n_chunks = (num_k_blocks + chunk - 1) / chunk;

  if (warp == 0) {
    // ===== TMA producer =====
    if (lane == 0) {
      uint32_t it = 0;
      for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
        int I, J, h;
        tile_coords(t, I, J, h);
        const int row_a = I * 256 + h * kTileM, row_b = J * 256;
        if (row_a >= m) continue;
        for (int kb = 0; kb < num_k_blocks; ++kb, ++it) {
          const int s = it % kStages;
          const uint32_t phase = (it / kStages) & 1;
          mbar_wait(&empty_bar[s], phase ^ 1);  // slot free (first pass returns immediately)
          mbar_expect_tx(&full_bar[s], kStageBytes);
          uint8_t *st = smem + s * kStageBytes;
          tma_load_2d(st, &map_hi_a, &full_bar[s], kb * kBlockK, row_a);
          tma_load_2d(st + kBytesA, &map_lo_a, &full_bar[s], kb * kBlockK, row_a);
          tma_load_2d(st + 2 * kBytesA, &map_hi_b, &full_bar[s], kb * kBlockK, row_b);
          tma_load_2d(st + 2 * kBytesA + kBytesB, &map_lo_b, &full_bar[s], kb * kBlockK, row_b);
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer (single thread) =====
    if (lane == 0) {
      uint32_t it = 0, ch = 0;
      for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
        int I, J, h;
        tile_coords(t, I, J, h);
        if (I * 256 + h * kTileM >= m) continue;
        for (int kb = 0; kb < num_k_blocks; ++kb, ++it) {
          const int s = it % kStages;
          const int a = ch % kAccums;
          const bool first = (kb % chunk) == 0;
          if (first) {  // a fresh accumulator: the drain warps must have emptied it
            mbar_wait(&acc_empty[a], ((ch / kAccums) & 1) ^ 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          }
          mbar_wait(&full_bar[s], (it / kStages) & 1);  // TMA bytes have landed
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint32_t base = smem_u32(smem + s * kStageBytes);
          const uint64_t a_hi = umma_desc_sw128(base), a_lo = umma_desc_sw128(base + kBytesA);
          const uint64_t b_hi = umma_desc_sw128(base + 2 * kBytesA), b_lo = umma_desc_sw128(base + 2 * kBytesA + kBytesB);
          const uint32_t acc = tmem_base + (uint32_t)(a * kTileN);
          // small terms first: lo hi^T, hi lo^T, then hi hi^T (advance 32 bytes along K: +2 in the addr>>4 field)
#pragma unroll
          for (int k = 0; k < kBlockK / kUmmaK; ++k)
            umma_tf32(acc, a_lo + (uint64_t)(2 * k), b_hi + (uint64_t)(2 * k), (first && k == 0) ? 0u : 1u);
#pragma unroll
          for (int k = 0; k < kBlockK / kUmmaK; ++k) umma_tf32(acc, a_hi + (uint64_t)(2 * k), b_lo + (uint64_t)(2 * k), 1u);
#pragma unroll
          for (int k = 0; k < kBlockK / kUmmaK; ++k) umma_tf32(acc, a_hi + (uint64_t)(2 * k), b_hi + (uint64_t)(2 * k), 1u);
          umma_commit(&empty_bar[s]);  // frees the smem slot once these MMAs have read it
          if ((kb % chunk) == chunk - 1 || kb == num_k_blocks - 1) {
            umma_commit(&acc_full[a]);  // chunk complete: hand the accumulator to the drain warps
            ++ch;
          }
        }
      }
    }
  } else {
    // ===== chunk drain + epilogue: warps 2..9 =====
    const int q = warp & 3;               // TMEM lane quarter this warp may touch
    const int half = (warp - 2) >> 2;     // which 128 of the 256 accumulator columns
    const uint32_t lane_addr = ((uint32_t)(q * 32) << 16) + (uint32_t)(half * 128);
    uint32_t ch = 0;
    for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
      int I, J, h;
      tile_coords(t, I, J, h);
      const int row_a = I * 256 + h * kTileM;
      if (row_a >= m) continue;
      float acc[128];
#pragma unroll
      for (int c = 0; c < 128; ++c) acc[c] = 0.f;
      for (int cc = 0; cc < n_chunks; ++cc, ++ch) {
        const int a = ch % kAccums;
        mbar_wait(&acc_full[a], (ch / kAccums) & 1);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
        for (int c0 = 0; c0 < 128; c0 += 32) {
          float v[32];
          tmem_ld32(tmem_base + lane_addr + (uint32_t)(a * kTileN + c0), v);
#pragma unroll
          for (int c = 0; c < 32; ++c) acc[c0 + c] += v[c];
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncwarp();
        if (lane == 0) mbar_arrive(&acc_empty[a]);
      }
      // store: thread = output row i, 128 consecutive columns.  Element (i, j) is written directly when its 128-column
      // block is not left of the row's 128-row block, and mirrored to (j, i) when it is strictly right of it.
      const int64_t i = (int64_t)row_a + q * 32 + lane;
      const int64_t j0 = (int64_t)J * 256 + half * 128;
      const int64_t rb = (int64_t)row_a;  // first row of this tile's 128-row block
      const bool direct = j0 >= rb, mirror = j0 >= rb + 128;
      if (i < m && direct) {
        float *dst = g + i * ldg + j0;
        if (j0 + 128 <= m && (reinterpret_cast<uintptr_t>(dst) & 15) == 0) {
#pragma unroll
          for (int c = 0; c < 128; c += 4)
            *reinterpret_cast<float4 *>(dst + c) = make_float4(acc[c], acc[c + 1], acc[c + 2], acc[c + 3]);
        } else {
#pragma unroll
          for (int c = 0; c < 128; ++c)
            if (j0 + c < m) dst[c] = acc[c];
        }
      }
      if (i < m && mirror) {  // lanes hold consecutive i -> coalesced
#pragma unroll
        for (int c = 0; c < 128; ++c)
          if (j0 + c < m) g[(j0 + c) * ldg + i] = acc[c];
      }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  }
  __syncthreads();
  if (warp == 2) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(kTmemCols) : "memory");
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

// 2-D fp32 tensor [rows][kp] (kp contiguous), box = 32 (K) x box_rows, 128-byte swizzle, OOB rows read as 0
static int make_map(CUtensorMap *map, float *base, int64_t rows, int64_t kp, int box_rows) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) return fail(WDGH_ENODEV, "wdgh_gram: cuTensorMapEncodeTiled is not available from the driver");
  cuuint64_t dims[2] = {(cuuint64_t)kp, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)kp * sizeof(float)};
  cuuint32_t box[2] = {(cuuint32_t)kBlockK, (cuuint32_t)box_rows};
  cuuint32_t elem[2] = {1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, base, dims, strides, box, elem, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
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
  return 2 * m * dpad;
}

int wdgh_gram_tc_launch(const float *z, int64_t m, int64_t d, int64_t ldz, float *g, int64_t ldg, float *workspace,
                        int chunk, cudaStream_t st) {
  if (workspace == nullptr) return fail(WDGH_EINVAL, "wdgh_gram: the tensor-core path needs the workspace");
  if (reinterpret_cast<uintptr_t>(workspace) % 16 != 0) return fail(WDGH_EINVAL, "wdgh_gram: workspace must be 16-byte aligned");
  if (m >= (int64_t)1 << 30) return fail(WDGH_EINVAL, "wdgh_gram: too many rows");
  const int64_t dpad = ceil_div(d, kBlockK) * kBlockK;
  float *hi = workspace, *lo = workspace + m * dpad;
  gram_split_kernel<<<persistent_grid(ceil_div(m * dpad, 256), 8), 256, 0, st>>>(z, m, d, ldz, dpad, hi, lo);
  WDGH_LAUNCHED("gram_split_kernel");
  CUtensorMap map_hi_a, map_lo_a, map_hi_b, map_lo_b;
  int rc = make_map(&map_hi_a, hi, m, dpad, kTileM);
  if (!rc) rc = make_map(&map_lo_a, lo, m, dpad, kTileM);
  if (!rc) rc = make_map(&map_hi_b, hi, m, dpad, kTileN);
  if (!rc) rc = make_map(&map_lo_b, lo, m, dpad, kTileN);
  if (rc) return rc;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(gram_tcgen05_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kGramSmem);
    if (e != cudaSuccess) return fail_cuda(e, "cudaFuncSetAttribute(gram_tcgen05_kernel)");
    configured = true;
  }
  const int num_k_blocks = (int)(dpad / kBlockK);
  const int64_t S = ceil_div(m, 256);
  const int num_tiles = (int)(S * (S + 1));  // two 128-row halves per upper-triangle super tile
  const int grid = num_tiles < sm_count() ? num_tiles : sm_count();
  gram_tcgen05_kernel<<<grid, kGramThreads, kGramSmem, st>>>(map_hi_a, map_lo_a, map_hi_b, map_lo_b, m, num_k_blocks,
                                                            chunk < 1 ? 1 : chunk, num_tiles, g, ldg);
  WDGH_LAUNCHED("gram_tcgen05_kernel");
  return 0;
}
