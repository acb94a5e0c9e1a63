// CUDA-core implementation of the dmc_conv_desc contract: one thread per output element, fp32 accumulate over the
// same bf16 operands and packed weights the tcgen05 kernel reads.  TEST INFRASTRUCTURE (impl = 1): it exists so a
// wrong TMA box / UMMA descriptor can be told apart from a precision effect on the GPU box.  Never on the product path.
#include "common.cuh"
#include "kernels.h"

namespace dmc {

struct ConvRefArgs {
  const __nv_bfloat16* src[3];
  int src_c[3], src_taps[3], nsrc;
  int B, Hin, Win, stride, up_phase;
  const __nv_bfloat16* w;
  int Cout, Ktot;
  const float* bias;
  const float* cond;
  int cond_stride;
  const __nv_bfloat16* residual;
  __nv_bfloat16* out;
  float* out_nchw;
};

__global__ void __launch_bounds__(256) conv_ref_kernel(ConvRefArgs a) {
  const int Hout = a.Hin / a.stride, Wout = a.Win / a.stride;
  const int oscale = a.up_phase >= 0 ? 2 : 1;
  const int ph = a.up_phase >= 0 ? (a.up_phase >> 1) : 0, pw = a.up_phase >= 0 ? (a.up_phase & 1) : 0;
  const int out_H = Hout * oscale, out_W = Wout * oscale;
  const size_t total = static_cast<size_t>(a.B) * Hout * Wout * a.Cout;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(i % a.Cout);
    size_t pixi = i / a.Cout;
    const int ow = static_cast<int>(pixi % Wout);
    const int oh = static_cast<int>((pixi / Wout) % Hout);
    const int n = static_cast<int>(pixi / (static_cast<size_t>(Wout) * Hout));
    const __nv_bfloat16* wrow = a.w + static_cast<size_t>(c) * a.Ktot;
    float acc = 0.f;
    int koff = 0;
    for (int s = 0; s < a.nsrc; ++s) {
      const int C = a.src_c[s], taps = a.src_taps[s];
      for (int t = 0; t < taps; ++t) {
        int dh = 0, dw = 0;
        if (taps == 9) { dh = t / 3 - 1; dw = t % 3 - 1; }
        else if (taps == 4) { dh = ph - 1 + t / 2; dw = pw - 1 + t % 2; }
        const int ih = oh * a.stride + dh, iw = ow * a.stride + dw;
        if (ih >= 0 && ih < a.Hin && iw >= 0 && iw < a.Win) {
          const __nv_bfloat16* xp = a.src[s] + ((static_cast<size_t>(n) * a.Hin + ih) * a.Win + iw) * C;
          for (int k = 0; k < C; ++k) acc = fmaf(__bfloat162float(xp[k]), __bfloat162float(wrow[koff + k]), acc);
        }
        koff += C;
      }
    }
    if (a.bias) acc += a.bias[c];
    if (a.cond) acc += a.cond[static_cast<size_t>(n) * a.cond_stride + c];
    const int yh = oh * oscale + ph, yw = ow * oscale + pw;
    const size_t opix = (static_cast<size_t>(n) * out_H + yh) * out_W + yw;
    if (a.residual) acc += __bfloat162float(a.residual[opix * a.Cout + c]);
    if (a.out) a.out[opix * a.Cout + c] = __float2bfloat16(acc);
    if (a.out_nchw) a.out_nchw[((static_cast<size_t>(n) * a.Cout + c) * out_H + yh) * out_W + yw] = acc;
  }
}

int launch_conv_ref(const dmc_conv_desc& d, cudaStream_t st) {
  DMC_REQUIRE(d.nsrc >= 1 && d.nsrc <= 3 && d.weight && (d.out_bf16 || d.out_f32_nchw), "conv_ref: bad arguments");
  DMC_REQUIRE(d.stats == nullptr, "conv_ref: the debug kernel does not produce GroupNorm statistics");
  ConvRefArgs a;
  int k = 0;
  for (int s = 0; s < 3; ++s) {
    a.src[s] = reinterpret_cast<const __nv_bfloat16*>(s < d.nsrc ? d.src[s] : nullptr);
    a.src_c[s] = s < d.nsrc ? d.src_c[s] : 0;
    a.src_taps[s] = s < d.nsrc ? d.src_taps[s] : 0;
    k += a.src_c[s] * a.src_taps[s];
  }
  DMC_REQUIRE(k == d.Ktot, "conv_ref: Ktot=%d does not match the sources (%d)", d.Ktot, k);
  a.nsrc = d.nsrc; a.B = d.B; a.Hin = d.Hin; a.Win = d.Win; a.stride = d.stride; a.up_phase = d.up_phase;
  a.w = reinterpret_cast<const __nv_bfloat16*>(d.weight);
  a.Cout = d.Cout; a.Ktot = d.Ktot; a.bias = d.bias; a.cond = d.cond; a.cond_stride = d.cond_stride;
  a.residual = reinterpret_cast<const __nv_bfloat16*>(d.residual);
  a.out = reinterpret_cast<__nv_bfloat16*>(d.out_bf16);
  a.out_nchw = d.out_f32_nchw;
  size_t total = static_cast<size_t>(d.B) * (d.Hin / d.stride) * (d.Win / d.stride) * d.Cout;
  int blocks = static_cast<int>(std::min<size_t>((total + 255) / 256, static_cast<size_t>(num_sms()) * 16));
  conv_ref_kernel<<<blocks, 256, 0, st>>>(a);
  DMC_CUDA_OK(cudaGetLastError());
  return 0;
}

}  // namespace dmc
