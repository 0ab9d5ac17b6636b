// stage1_tc.cu -- placeholder until the tcgen05 shortlist kernel lands.
#include "common.cuh"
int32_t cb_stage1_tc_shortlist(cb_index* ix, const float* dQ, int64_t nrows, float* topv, int32_t* topi,
                               int* nsplit_out, float* guard_out, cudaStream_t st) {
  (void)ix; (void)dQ; (void)nrows; (void)topv; (void)topi; (void)nsplit_out; (void)guard_out; (void)st;
  return CB_ERR_UNSUPPORTED;
}
