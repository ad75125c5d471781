// PHMLinear tensor-core path (tcgen05 / TMEM / TMA) — placeholder until the kernel lands:
// reports "unsupported" so the dispatch in api.cu uses the fp32 FFMA path.
#include "common.cuh"

int phm_tc_supported(int, int, int, int, int) { return 0; }
size_t phm_tc_fwd_workspace_bytes(int, int, int, int, int) { return 16; }
size_t phm_tc_bwd_workspace_bytes(int, int, int, int, int) { return 16; }
int phm_tc_fwd(const float*, const float*, const float*, const float*, const float*, float*, int, int, int, int, int, int, void*,
               cudaStream_t) {
  phc_set_error("phm_tc_fwd: not built");
  return PHC_ERR_UNSUPPORTED;
}
int phm_tc_bwd(const float*, const float*, const float*, const float*, float*, float*, float*, float*, int, int, int, int, int, void*,
               cudaStream_t) {
  phc_set_error("phm_tc_bwd: not built");
  return PHC_ERR_UNSUPPORTED;
}
