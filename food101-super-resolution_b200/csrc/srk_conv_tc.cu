// tcgen05 / TMEM / TMA implicit-GEMM convolution (placeholder until the kernel lands).
#include "srk_common.cuh"

namespace srk {
bool conv_tc_shape_ok(int, int, int, int, int, int) { return false; }
int conv_fprop_tc_launch(const srk_tensor*, const srk_tensor*, const void*, int, int, int, const float*, int,
                         const float*, const srk_tensor*, int, cudaStream_t) {
  SRK_FAIL("tcgen05 conv path not built");
}
bool conv_wgrad_tc_shape_ok(const srk_tensor*, const srk_tensor*, int, int) { return false; }
int64_t conv_wgrad_tc_workspace(const srk_tensor*, const srk_tensor*, int, int) { return 0; }
int conv_wgrad_tc_launch(const srk_tensor*, const srk_tensor*, float*, float*, int, int, void*, cudaStream_t) {
  SRK_FAIL("tcgen05 wgrad path not built");
}
}  // namespace srk
extern "C" int srk_tc_probe(int, float*, int) { srk::set_error("probe not built"); return 1; }
