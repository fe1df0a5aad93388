// tcgen05 (3xTF32) dense contractions -- placeholder until the tensor-core kernels land:
// reports "unsupported" so every contraction runs on the fp32 FFMA kernels of gemm_ffma.cu.
#include "common.cuh"

namespace cgcn {

bool tc_rowpanel_supported(int64_t, int64_t, int, int, const void*, const void*) { return false; }
bool tc_gram_supported(int64_t, int64_t, int, int, const void*, const void*) { return false; }
size_t tc_workspace_bytes() { return 256; }

int gemm_rowpanel_tc(const float*, int64_t, const float*, int, const float*, float*, int64_t, int64_t, int, int,
                     const int32_t*, int, void*, size_t, cudaStream_t) {
  set_error("tcgen05 row-panel GEMM not built");
  return CGCN_ERR_INVALID;
}
int gemm_gram_tc(const float*, int64_t, const float*, int64_t, float*, int64_t, int64_t, int, int, int, void*, size_t,
                 cudaStream_t) {
  set_error("tcgen05 gram GEMM not built");
  return CGCN_ERR_INVALID;
}

}  // namespace cgcn
