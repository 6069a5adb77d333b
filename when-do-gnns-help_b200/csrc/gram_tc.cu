// tcgen05 / TMEM Gram kernel -- placeholder translation unit until the tensor-core kernel lands.
#include "common.cuh"
using namespace wdgh;
int wdgh_gram_tc_launch(const float *, int64_t, int64_t, int64_t, float *, int64_t, cudaStream_t) {
  return fail(WDGH_ESTATE, "wdgh_gram: tensor-core path not built in this revision");
}
