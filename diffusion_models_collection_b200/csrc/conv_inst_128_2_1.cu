// kernel variants of the tcgen05 convolution for the tile configuration BN=128, MT=2, CG=1 (see conv_umma_kernel.cuh)
#include "conv_umma_kernel.cuh"

namespace dmc {
int launch_conv_128_2_1(const ConvPrepared* P, const ConvKParams& kp, cudaStream_t st) {
  return launch_tile_cfg<128, 2, 1>(P, kp, st);
}
}  // namespace dmc
