// Internal launch interface between the plan interpreter (plan.cu) and the kernel translation units.
#pragma once

#include <cuda.h>
#include <cuda_runtime.h>

#include <algorithm>

#include "../../include/dmc.h"

namespace dmc {

// sched.cu
int launch_step(bool ddpm, const dmc_step_desc& d, cudaStream_t st);
int launch_advance(int* counter, const int64_t* t_table, int64_t* t_out, int n, cudaStream_t st);
int launch_q_sample(const float* x0, const float* noise, const int64_t* t, const float* sa, const float* s1, float* out,
                    int B, int n, cudaStream_t st);

// elementwise.cu
int launch_cond(const dmc_cond_desc& d, cudaStream_t st);
int cond_num_launches(const dmc_cond_desc& d);
int launch_stem(const dmc_stem_desc& d, cudaStream_t st);
int launch_gn_stats(const dmc_gn_stats_desc& d, cudaStream_t st);
int launch_gn_apply(const dmc_gn_apply_desc& d, cudaStream_t st);
int launch_upsample(const dmc_upsample_desc& d, cudaStream_t st);
int launch_head_fused(const dmc_head_desc& d, cudaStream_t st);
int launch_head_taps(const dmc_head_taps_desc& d, cudaStream_t st);
int launch_stem_cols(const dmc_stem_cols_desc& d, cudaStream_t st);
int launch_gn_coeff(const dmc_gn_coeff_desc& d, cudaStream_t st);
bool head_fused_supported(const dmc_head_desc& d);

// dit_ops.cu
int launch_dit_cond(const dmc_dit_cond_desc& d, cudaStream_t st);
int dit_cond_num_launches(const dmc_dit_cond_desc& d);
int launch_patch_embed(const dmc_patch_embed_desc& d, cudaStream_t st);
int launch_ln_modulate(const dmc_ln_mod_desc& d, cudaStream_t st);

// attention.cu : CUDA-core flash kernel (any L; debug / shapes the tensor-core kernel does not cover)
int launch_attention(const dmc_attn_desc& d, cudaStream_t st);
// attention_umma.cu : tcgen05 kernel
struct AttnPrepared;
bool attention_umma_supported(const dmc_attn_desc& d);
int attention_prepare(const dmc_attn_desc& d, AttnPrepared** out);
void attention_release(AttnPrepared* p);
int launch_attention_umma(const AttnPrepared* p, cudaStream_t st);

// conv_umma.cu : the tcgen05 implicit-GEMM convolution.  `prepared` holds the TMA descriptors and tile geometry
struct ConvPrepared;
int conv_prepare(const dmc_conv_desc& d, ConvPrepared** out);
void conv_release(ConvPrepared* p);
int launch_conv(const dmc_conv_desc& d, const ConvPrepared* p, cudaStream_t st);
bool conv_gn_supported(int B, int Hout, int Wout, int Cout, int max_gsz);
bool conv_affine_supported(int B, int H, int W, int Cin, int Cout);
// train_ops.cu : non-GEMM backward kernels
uint32_t dropout_threshold(float p);
int launch_gn_backward(const dmc_gn_bwd_desc& d, cudaStream_t st);
int launch_attention_backward(const dmc_attn_bwd_desc& d, cudaStream_t st);
// dit_train_ops.cu
int launch_dit_gate_ln_mod(const dmc_dit_glm_desc& d, cudaStream_t st);
int launch_dit_gate_ln_mod_backward(const dmc_dit_glm_bwd_desc& d, cudaStream_t st);
int launch_gelu_forward(const void* u, void* m, int64_t n, float drop_p, uint32_t seed, cudaStream_t st);
int launch_gelu_backward(const void* u, const void* dm, void* du, int64_t n, float drop_p, uint32_t seed, cudaStream_t st);
int launch_channel_sum(const void* src, float* out, int B, int HW, int C, int per_image, int accumulate, float* scratch, cudaStream_t st);
int launch_dilate2x(const void* src, void* dst, int B, int h, int w, int C, cudaStream_t st);
int launch_attention_backward_mma(const dmc_attn_bwd_desc& d, cudaStream_t st);
int launch_pack_weights(const dmc_pack_item* items_dev, int n, cudaStream_t st);
long long gn_backward_scratch_floats(const dmc_gn_bwd_desc& d);
int launch_opt_grad_norm(const dmc_opt_item* items, const dmc_opt_chunk* chunks, int n_chunks, float* partial, float* norm, cudaStream_t st);
int launch_opt_adamw(const dmc_opt_item* items, const dmc_opt_chunk* chunks, int n_chunks, const dmc_adamw_desc& h, const float* norm, cudaStream_t st);
int launch_add_bf16(void* dst, const void* src, size_t n, int accumulate, cudaStream_t st);
int launch_block_sum2x2(const void* dhigh, void* dlow, int B, int H, int W, int C, int accumulate, cudaStream_t st);
int launch_nchw_to_nhwc_pad(const float* src, void* dst, int B, int Cs, int HW, int Cd, cudaStream_t st);
int launch_conv_dgrad_strided(const void* dy, const float* w, void* dx, int B, int Hin, int Win, int Cin, int Cout, int stride,
                              int accumulate, cudaStream_t st);
// conv_wgrad.cu : weight gradient on tcgen05 (training)
int conv_wgrad_splits(const dmc_wgrad_desc& d);
int launch_conv_wgrad(const dmc_wgrad_desc& d, cudaStream_t st);
// conv_ref.cu : CUDA-core debug implementation of the same contract (tests only)
int launch_conv_ref(const dmc_conv_desc& d, cudaStream_t st);

// driver entry point for TMA descriptors, resolved by dmc_init()
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn encode_tiled_fn();

}  // namespace dmc
