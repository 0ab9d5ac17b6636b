// stage34_tc.cu -- placeholder until the tcgen05 kernel lands (replaced in the next commit).
#include "common.cuh"
bool cb_stage34_tc_supported(const cb_index* ix, int T) { (void)ix; (void)T; return false; }
int32_t cb_stage34_tc(cb_index* ix, const float* dQ, int nq, int T, int W, const uint32_t* d_bitmap,
                      const int64_t* d_list_off, int32_t* d_cursors, uint64_t* d_pairs, cudaStream_t st) {
  (void)ix; (void)dQ; (void)nq; (void)T; (void)W; (void)d_bitmap; (void)d_list_off; (void)d_cursors; (void)d_pairs; (void)st;
  cb_set_error("tcgen05 scoring kernel not built");
  return CB_ERR_UNSUPPORTED;
}
