// tcgen05 / TMA engines (placeholder until the tensor-core kernels land in this file).
#include "common.cuh"
namespace vkn {
bool tc_supported(const VknShape &) { return false; }
int pool_tc_chunks(const VknShape &) { return 0; }
int launch_pool_tc(const VknShape &, const void *, const void *, float *, float *, int *, cudaStream_t) {
  VKN_FAIL(VKN_E_UNSUPPORTED, "tcgen05 pooling engine not built");
}
int launch_maskgemm_tc(const VknShape &, const void *, const float *, int, void *, void *, cudaStream_t) {
  VKN_FAIL(VKN_E_UNSUPPORTED, "tcgen05 mask-conv engine not built");
}
}  // namespace vkn
