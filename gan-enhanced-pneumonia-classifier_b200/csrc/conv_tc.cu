// tcgen05 / TMEM / TMA implicit-GEMM convolutions (sm_100a).  Placeholder until the tensor-core kernels
// land: every entry reports "does not qualify" (1) so the dispatcher uses the SIMT path, or fails
// loudly when B200GAN_ALGO_TCGEN05 is forced.
#include "common.cuh"

namespace b200gan {
int tc_conv_fprop(const b200gan_conv*, const b200gan_view*, const void*, const b200gan_view*, cudaStream_t) { return 1; }
int tc_conv_dgrad(const b200gan_conv*, const b200gan_view*, const void*, const b200gan_view*, cudaStream_t) { return 1; }
int tc_conv_wgrad(const b200gan_conv*, const b200gan_view*, const b200gan_view*, float*, cudaStream_t) { return 1; }
}  // namespace b200gan
