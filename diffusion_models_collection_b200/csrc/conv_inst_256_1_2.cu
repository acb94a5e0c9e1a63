// kernel variants of the tcgen05 convolution for the tile configuration BN=256, MT=1, CG=2 (see conv_umma_kernel.cuh)
#include "conv_umma_kernel.cuh"

namespace dmc {
int launch_conv_256_1_2(const ConvPrepared* P, const ConvKParams& kp, cudaStream_t st) {
  return launch_tile_cfg<256, 1, 2>(P, kp, st);
}
}  // namespace dmc
