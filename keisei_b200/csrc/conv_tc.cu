// conv_tc.cu — placeholder until the tcgen05 kernels land.
#include "kb_common.cuh"
#include "kb_kernels.h"
int kbk_conv3x3_tc_supported(int, int, int) { return 0; }
int kbk_conv3x3_tc(const void*, const void*, void*, int, int, int, const ConvEpi&, int, cudaStream_t) {
  kb_set_error("tcgen05 conv not built"); return KB_ERR_UNSUPPORTED;
}
int kbk_conv3x3_wgrad_tc(const void*, const void*, float*, int, int, int, int, int, cudaStream_t) {
  kb_set_error("tcgen05 wgrad not built"); return KB_ERR_UNSUPPORTED;
}
