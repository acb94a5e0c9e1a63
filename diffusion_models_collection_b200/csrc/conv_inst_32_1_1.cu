// kernel variants of the tcgen05 convolution for the tile configuration BN=32, MT=1, CG=1 (see conv_umma_kernel.cuh)
#include "conv_umma_kernel.cuh"

namespace dmc {
int launch_conv_32_1_1(const ConvPrepared* P, const ConvKParams& kp, cudaStream_t st) {
  return launch_tile_cfg<32, 1, 1>(P, kp, st);
}
}  // namespace dmc
