// Internal (non-ABI) entry points shared between translation units.
#pragma once
#include "common.cuh"

namespace wdgh {

constexpr int kHistSmemBins = 4096;  // C*C bins kept in shared memory per CTA

// Folds one warp-wide batch of class-pair keys (key < 0: nothing to count) into the histogram:
// one atomic per distinct key, multiplicity from __match_any_sync + popc.
__device__ __forceinline__ void fold_keys(int key, unsigned *s_hist, unsigned long long *g_hist, bool use_smem) {
  const unsigned peers = __match_any_sync(0xffffffffu, key);
  if (key >= 0 && (threadIdx.x & 31) == (__ffs(peers) - 1)) {
    const unsigned c = __popc(peers);
    if (use_smem) atomicAdd(&s_hist[key], c);
    else atomicAdd(&g_hist[key], (unsigned long long)c);
  }
}

// structure.cu: zero the counters, optionally build the 1-byte label copy (returns it, or nullptr when
// the int32 labels must be used: C > 254 or no scratch)
int structure_prepare(const int32_t *labels, int64_t n_labels, int C, uint8_t *labels_u8_scratch, int64_t *counters,
                      double *node_sum, const uint8_t **labels8_out, cudaStream_t st);
// structure.cu: split-row chunks + per-node reductions (everything after the per-row edge pass)
int structure_finish(const int64_t *rowptr, const int32_t *col, int64_t n, const int32_t *labels,
                     const uint8_t *labels8, int C, const int64_t *plan_i64, const int64_t *plan_host,
                     int64_t *counters, double *node_sum, int32_t *deg_nsl, int32_t *match_nsl, int64_t row_offset,
                     cudaStream_t st);

}  // namespace wdgh
