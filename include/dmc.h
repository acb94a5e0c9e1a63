/*
 * dmc.h -- C ABI of libdmc_b200.so: the B200 (sm_100a) denoising hot path of
 * sunyzhi55/Diffusion_Models_Collection.
 *
 * The reference has no FFI / plugin interface of its own (it is pure PyTorch); this
 * header is the drop-in boundary a maintainer binds with ctypes.  Each entry
 * names the reference code it replaces (paths relative to the reference repo root).
 *
 * Conventions
 *   - extern "C", plain pointers and sizes; no torch / C++ types.  Every function returns 0 on success
 *     (or a non-negative index where stated) and a negative value on error; dmc_last_error() returns
 *     a thread-local message.  No C++ exception crosses the boundary.
 *   - All data pointers are DEVICE pointers owned by the caller (torch tensors); the library never
 *     allocates device memory.  `stream` is a cudaStream_t passed as void*; hot calls are
 *     allocation-free, sync-free and CUDA-graph-capturable.
 *   - Activations inside the model are bf16 NHWC; the model boundary (x, eps) is fp32 NCHW exactly
 *     like the reference's `model(x, t, y)`.
 */
#ifndef DMC_H_
#define DMC_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DMC_ABI_VERSION 1

#if defined(__GNUC__)
#define DMC_API __attribute__((visibility("default")))
#else
#define DMC_API
#endif

DMC_API const char* dmc_last_error(void);
DMC_API int dmc_abi_version(void);
/* Resolves the driver entry points (cuTensorMapEncodeTiled) and caches device properties.  Returns the
 * SM count of the current device, or <0 when no sm_100 device is usable (the product path then
 * fails loudly -- there is no CPU fallback). */
DMC_API int dmc_init(void);

/* ------------------------------------------------------------------------------------------------
 * Fused scheduler steps  (stateless; replace ~80-190 ATen calls + 1 host sync per step)
 * ---------------------------------------------------------------------------------------------- */

/* Per-step scalar coefficients, computed by the host with the reference's own fp32 expressions. */
typedef struct {
  float sqrt_one_minus_a; /* sqrt(1 - a_t)                       diffusion/ddim.py:186,315 */
  float sqrt_a;           /* sqrt(a_t)                            (true division, not rsqrt) */
  float sqrt_a_next;      /* sqrt(a_next), a_next = 1 if t_next<0 diffusion/ddim.py:176-179,203 */
  float dir_coef;         /* sqrt(clamp(1 - a_next - sigma^2, 0)) diffusion/ddim.py:201 */
  float sigma;            /* eta * sqrt(clamp(..., 0))            diffusion/ddim.py:193-199 */
} dmc_ddim_coef;

typedef struct {
  float sqrt_recip_a;   /* sqrt(1/a_t)          diffusion/ddpm.py:170-172 */
  float sqrt_recipm1_a; /* sqrt(1/a_t - 1)      diffusion/ddpm.py:173-175 */
  float coef1;          /* posterior_mean_coef1 diffusion/ddpm.py:184 */
  float coef2;          /* posterior_mean_coef2 diffusion/ddpm.py:185 */
  float noise_scale;    /* (t != 0) * exp(0.5 * posterior_log_variance_clipped[t])  diffusion/ddpm.py:218-220 */
} dmc_ddpm_coef;

/* Guidance / x0 post-processing shared by both samplers. */
typedef struct {
  float cfg_scale;   /* eps = eps_u + cfg_scale * (eps_c - eps_u) when eps_u != NULL   ddim.py:302 ddpm.py:292 */
  int32_t clip_mode; /* 0 none, 1 clamp(x0,-1,1) (ddim.py:189-190), 2 dynamic threshold (ddim.py:320-325) */
  int32_t q_lo;      /* dynamic threshold: indices into the ascending sort of |x0| ...                       */
  int32_t q_hi;      /* ... and the lerp weight, exactly as torch.quantile derives them (fp32 rank)          */
  float q_weight;
} dmc_guidance;

/* x_out = DDIM update of x.  Replaces DDIM.p_sample (diffusion/ddim.py:154-208) and the CFG / dynamic
 * threshold block of DDIM.sample_with_cfg (:300-339).  x, eps_c, eps_u, noise, x_out: fp32 [B, n_per_sample]
 * (eps_u and noise may be NULL; noise is required iff coef->sigma != 0).  coef is a DEVICE pointer
 * (one table row per step, so a captured graph can walk it). */
DMC_API int dmc_ddim_step(const float* x, const float* eps_c, const float* eps_u, const float* noise, float* x_out,
                  int32_t B, int32_t n_per_sample, const dmc_ddim_coef* coef_dev, const dmc_guidance* g,
                  void* stream);

/* x_out = DDPM posterior sample.  Replaces DDPM.p_mean_variance + p_sample (diffusion/ddpm.py:151-220)
 * and the CFG block of DDPM.sample_with_cfg (:289-324).  noise is always read (the reference draws it
 * at every step, including t == 0 where noise_scale is 0). */
DMC_API int dmc_ddpm_step(const float* x, const float* eps_c, const float* eps_u, const float* noise, float* x_out,
                  int32_t B, int32_t n_per_sample, const dmc_ddpm_coef* coef_dev, const dmc_guidance* g,
                  void* stream);

/* The same two steps with the coefficient row chosen on the DEVICE: row = coef_table_dev + step_index_dev[0].  Lets one
 * captured CUDA graph of "forward + step" be replayed for every step of a sampling loop (x_out may alias x). */
DMC_API int dmc_ddim_step_at(const float* x, const float* eps_c, const float* eps_u, const float* noise, float* x_out,
                     int32_t B, int32_t n_per_sample, const dmc_ddim_coef* coef_table_dev,
                     const int32_t* step_index_dev, const dmc_guidance* g, void* stream);
DMC_API int dmc_ddpm_step_at(const float* x, const float* eps_c, const float* eps_u, const float* noise, float* x_out,
                     int32_t B, int32_t n_per_sample, const dmc_ddpm_coef* coef_table_dev,
                     const int32_t* step_index_dev, const dmc_guidance* g, void* stream);
/* Loop counter of a replayed step graph: cur = counter[0]; counter[1] = cur; counter[0] = cur + 1;
 * t_out[0..n) = t_table[cur]  (the timestep tensor the denoiser reads; replaces torch.full, ddim.py:287-297). */
DMC_API int dmc_advance(int32_t* counter_dev, const int64_t* t_table_dev, int64_t* t_out_dev, int32_t n, void* stream);

/* x_t = sqrt_acp[t_n] * x0 + sqrt_1m_acp[t_n] * noise with per-sample t (training).  Replaces q_sample
 * (diffusion/ddpm.py:84-104, ddim.py:87-107). */
DMC_API int dmc_q_sample(const float* x0, const float* noise, const int64_t* t, const float* sqrt_acp,
                 const float* sqrt_1m_acp, float* x_t, int32_t B, int32_t n_per_sample, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Training step building blocks.  models/unet_train.py assembles them into the UNet training step:
 * forward plan with kept activations, then per layer the entry points below, in reverse order of the forward ops.
 * ---------------------------------------------------------------------------------------------- */

/* Weight gradient of a 3x3 (padding 1) or 1x1 convolution, the backward of nn.Conv2d w.r.t. its weight
 * (models/unet.py:37,54,58,81,82,106,116 under autograd):
 *   dw[co, ci, r, s] (+)= sum_{n,h,w} dy[n, h, w, co] * x[n, h*stride + r - pad, w*stride + s - pad, ci]
 * tcgen05 GEMM over K = output pixels with both operands MN-major straight from the NHWC tensors; the pixels are cut into
 * `splits` slices whose partial sums are added in index order (deterministic).  Two kernels behind one entry point: stride 1,
 * rows of 8..64 pixels, >= 64 pixels per image and Cin % 128 == 0 take the row-slab kernel (128 co x 128 ci x the three
 * vertical taps of one column shift per CTA, one TMA box of BH + 2 rows for the three taps); everything else (4x4 maps,
 * stride 2, Cin % 128 == 64) the 128 co x 64 ci x 5-tap kernel.
 * (The input gradient needs no new kernel: it is dmc_plan_add_conv over dy with the weights packed transposed and
 * tap-flipped.) */
typedef struct {
  const void* x;    /* bf16 NHWC [B, Hin, Win, Cin], the convolution's input */
  const void* dy;   /* bf16 NHWC [B, Hin/stride, Win/stride, Cout], gradient of its output */
  int32_t B, Hin, Win, Cin, Cout; /* Cin % 64 == 0, Cout % 128 == 0 */
  int32_t stride;   /* 1 or 2 */
  int32_t taps;     /* 9: 3x3 with padding 1;  1: 1x1 */
  int32_t splits;   /* must equal dmc_conv_wgrad_splits() */
  float* partial;   /* scratch, fp32 [splits, Cout, taps, Cin] */
  float* dw;        /* fp32 [Cout, Cin, kh, kw] (the reference's parameter layout) */
  int32_t accumulate; /* 1: dw += ..., 0: dw = ... */
  /* optional window: dw is the parameter [dw_cout, dw_cin_total, kh, kw] and receives rows co < dw_cout, input channels
   * ci < dw_cin at columns dw_ci0 + ci (zero-padded operands: stem Cin 3 -> 64, head Cout 3 -> 128; one source of a fused 1x1
   * shortcut over concatenated inputs).  All zero = the whole [Cout, Cin, kh, kw] tensor. */
  int32_t dw_cout, dw_cin, dw_cin_total, dw_ci0;
} dmc_wgrad_desc;
DMC_API int dmc_conv_wgrad_splits(const dmc_wgrad_desc* d);
DMC_API int dmc_conv_wgrad(const dmc_wgrad_desc* d, void* stream);

/* GroupNorm(groups) (+ SiLU) (+ dropout) backward over the concatenation of up to two sources: the backward of the
 * dmc_gn_apply pass (models/unet.py:35-36,51-53,80).  Recomputes the normalised value from src and the forward partial
 * sums; writes (or accumulates into) the gradients of the sources and writes dgamma / dbeta. */
typedef struct {
  int32_t nsrc;
  const void* src[2];      /* bf16 [B, HW, c_i]: the un-normalised inputs of the forward pass */
  int32_t src_c[2];
  const float* stats[2];   /* their forward partial sums */
  int32_t stats_slots[2];
  const void* dout;        /* bf16 [B, HW, C]: gradient of the pass's output */
  void* dsrc[2];           /* bf16 [B, HW, c_i]: gradients of the sources */
  int32_t accumulate[2];   /* 1: dsrc += ..., 0: dsrc = ... */
  int32_t B, HW, groups;
  const float* gamma;
  const float* beta;
  float eps;
  int32_t silu;
  float drop_p;            /* dropout probability applied after SiLU in the forward pass (0: none) */
  uint32_t seed;           /* the forward pass's dropout seed */
  const uint32_t* seed_dev; /* the forward pass's seed_dev (see dmc_gn_apply_desc) */
  float* dgamma;           /* fp32 [C] */
  float* dbeta;            /* fp32 [C] */
  float* scratch;          /* fp32, dmc_gn_backward_scratch() floats */
} dmc_gn_bwd_desc;
/* number of fp32 scratch elements dmc_gn_backward needs for this geometry (B, HW, channels, groups) */
DMC_API int64_t dmc_gn_backward_scratch(const dmc_gn_bwd_desc* d);
DMC_API int dmc_gn_backward(const dmc_gn_bwd_desc* d, void* stream);

/* Attention backward (models/unet.py:88-96 under autograd): dqkv from qkv, the forward output and its gradient.
 * Head dim 64, L <= 256; warp-level tensor-core kernel (mma.sync bf16, fp32 accumulate), one CTA per (image, head). */
typedef struct {
  const void* qkv;  /* bf16 [B, L, 3C] */
  const void* out;  /* bf16 [B, L, C] forward output */
  const void* dout; /* bf16 [B, L, C] */
  void* dqkv;       /* bf16 [B, L, 3C] */
  int32_t B, L, heads, C;
} dmc_attn_bwd_desc;
DMC_API int dmc_attention_backward(const dmc_attn_bwd_desc* d, void* stream);

/* ---- DiT training step: the memory-bound glue between the GEMMs (models/dit_train.py; reference models/dit.py:111-132 under
 * autograd).  Forward:  x_out = x_in + gate[n] * y   (skipped when y == NULL: the first LayerNorm of the model)
 *                       h     = LayerNorm(x_out, eps) * (1 + scale[n]) + shift[n]  -> bf16 (the next GEMM's operand)
 * shift / scale: row n at ptr + n * mod_stride, gate: row n at ptr + n * gate_stride (chunks of adaLN_modulation outputs). */
typedef struct {
  const float* x_in;   /* fp32 [B, L, C] token stream */
  const void* y;       /* bf16 [B, L, C] branch output (attention out_proj / mlp fc2), or NULL */
  const float* gate;   /* per-image gate rows, NULL iff y is NULL */
  float* x_out;        /* fp32 [B, L, C], NULL iff y is NULL (x_out == x_in then) */
  void* h;             /* bf16 [B, L, C] */
  const float* shift;
  const float* scale;
  int32_t mod_stride;  /* row stride (floats) of shift / scale */
  int32_t gate_stride; /* row stride (floats) of gate (it comes from the PREVIOUS adaLN table when this LayerNorm opens a block) */
  int32_t B, L, C;     /* C % 128 == 0, 128 <= C <= 1024 */
  float eps;
  float drop_p;        /* nn.Dropout on y before the gated add (models/dit.py:100, training mode); 0: none */
  uint32_t seed;       /* counter-based mask: element i is kept when its 16 hash bits of (seed, i) are >= round(p * 65536) */
} dmc_dit_glm_desc;
DMC_API int dmc_dit_gate_ln_mod(const dmc_dit_glm_desc* d, void* stream);
/* Its backward: from dh (gradient of h, the GEMM's input gradient) and dx_out (gradient of the stream from later layers, or NULL):
 *   dx_in = dx_out + LayerNorm'(dh * (1 + scale)),  dy = dx_in * gate,
 *   dgate[n] = sum_l dx_in * y,  dshift[n] = sum_l dh,  dscale[n] = sum_l dh * xhat      (fp32 [B, C] each, contiguous)
 * One CTA per image; the per-image sums are added in a fixed order (deterministic).  dx_in may alias dx_out. */
typedef struct {
  const float* x;      /* fp32 [B, L, C]: x_out of the forward call (x_in when it had no y) */
  const void* dh;      /* bf16 [B, L, C] */
  const float* dx_out; /* fp32 [B, L, C] or NULL */
  const void* y;       /* bf16 [B, L, C] or NULL */
  const float* gate;   /* NULL iff y is NULL */
  const float* scale;
  int32_t mod_stride, gate_stride;
  float* dx_in;        /* fp32 [B, L, C] */
  void* dy;            /* bf16 [B, L, C], NULL iff y is NULL */
  float* dgate;        /* fp32 [B, C], NULL iff y is NULL */
  float* dshift;       /* fp32 [B, C] */
  float* dscale;       /* fp32 [B, C] */
  int32_t B, L, C;
  float eps;
  float drop_p;        /* the forward call's dropout (same seed: the mask is regenerated) */
  uint32_t seed;
  float* scratch;      /* optional fp32 [B, DMC_DIT_GLM_BWD_SLICES, 3, C]: with it every image is cut into row slices over several
                          CTAs whose partial sums a second launch adds in slice order; NULL: one CTA per image */
} dmc_dit_glm_bwd_desc;
#define DMC_DIT_GLM_BWD_SLICES 4
DMC_API int dmc_dit_gate_ln_mod_backward(const dmc_dit_glm_bwd_desc* d, void* stream);
/* nn.GELU() (erf form) followed by nn.Dropout(drop_p) on bf16 tensors of n elements (n % 8 == 0): m = drop(gelu(u));
 * du = drop(dm) * gelu'(u) with the same (seed-generated) mask   (models/dit.py:96-97) */
DMC_API int dmc_gelu_forward(const void* u_bf16, void* m_bf16, int64_t n, float drop_p, uint32_t seed, void* stream);
DMC_API int dmc_gelu_backward(const void* u_bf16, const void* dm_bf16, void* du_bf16, int64_t n, float drop_p, uint32_t seed,
                              void* stream);

/* out[n or 0][c] (+)= sum over pixels (and images unless per_image) of src[n, p, c]: bias and conditioning-row gradients.
 * scratch: fp32 [B, C], needed when per_image == 0 (the images are added in index order by a second kernel) */
DMC_API int dmc_channel_sum(const void* src_bf16, float* out, int32_t B, int32_t HW, int32_t C, int32_t per_image,
                    int32_t accumulate, float* scratch, void* stream);
/* dst[n, 2i, 2j, :] = src[n, i, j, :], zeros elsewhere (bf16 NHWC, src [B, h, w, C] -> dst [B, 2h, 2w, C]): spreads the output
 * gradient of a stride-2 convolution (Downsample, models/unet.py:106-109) onto the input grid, so that its input gradient is
 * the stride-1 tensor-core convolution with transposed, tap-flipped weights */
DMC_API int dmc_dilate2x(const void* src_bf16, void* dst_bf16, int32_t B, int32_t h, int32_t w, int32_t C, void* stream);
/* dlow[n, i, j, :] (+)= the 2x2 block sum of dhigh (backward of the nearest 2x upsample, models/unet.py:119) */
DMC_API int dmc_block_sum2x2(const void* dhigh_bf16, void* dlow_bf16, int32_t B, int32_t H, int32_t W, int32_t C,
                     int32_t accumulate, void* stream);
/* Re-pack of convolution weights after an optimizer step, all layers in ONE launch.  Item i converts the fp32 parameter
 * src [cout, cin_total, kh, kw] (taps = kh * kw; input channels ci0 .. ci0 + cin) into a bf16 GEMM operand:
 *   mode 0 (forward conv, dmc_conv_desc.weight):     dst[co * ld + col0 + tap * cin + ci]
 *   mode 1 (input-gradient conv, transposed + flipped): dst[ci * ld + col0 + (taps - 1 - tap) * cpad + co]
 * `items_dev` is an array in DEVICE memory. */
typedef struct {
  const float* src;
  void* dst;
  int32_t cout, cin_total, ci0, cin, taps, mode, ld, col0, cpad;
  int32_t pad_;
} dmc_pack_item;
DMC_API int dmc_pack_weights(const dmc_pack_item* items_dev, int32_t n_items, void* stream);
/* dst (+)= src over n bf16 elements, n % 8 == 0 (gradient of an identity residual branch, models/unet.py:72,99) */
DMC_API int dmc_add_bf16(void* dst_bf16, const void* src_bf16, int64_t n, int32_t accumulate, void* stream);
/* fp32 NCHW [B, Csrc, H*W] -> bf16 NHWC [B, H*W, Cdst >= Csrc] with zero-padded channels */
DMC_API int dmc_nchw_f32_to_nhwc_bf16(const float* src, void* dst, int32_t B, int32_t Csrc, int32_t HW, int32_t Cdst,
                              void* stream);
/* input gradient of a strided 3x3 / padding-1 convolution (Downsample layers), CUDA cores:
 * dx bf16 [B, Hin, Win, Cin] (+)= ..., dy bf16 [B, Hin/stride, Win/stride, Cout], w fp32 [Cout, Cin, 3, 3] */
DMC_API int dmc_conv_dgrad_strided(const void* dy, const float* w, void* dx, int32_t B, int32_t Hin, int32_t Win,
                           int32_t Cin, int32_t Cout, int32_t stride, int32_t accumulate, void* stream);

/* Multi-tensor optimizer step (utils/trainer.py:256-262: clip_grad_norm_(1.0), AdamW.step(), EMA update) for ALL parameters in two
 * launches.  `items_dev` (one per tensor) and `chunks_dev` (one per 16384-element chunk of a tensor: the work list) live in
 * DEVICE memory. */
typedef struct {
  float* p;        /* parameter */
  const float* g;  /* its gradient */
  float* m;        /* exp_avg */
  float* v;        /* exp_avg_sq */
  float* ema;      /* EMA copy of the parameter, or NULL */
  int64_t n;       /* elements */
} dmc_opt_item;
typedef struct {
  int32_t item;
  int32_t pad_;
  int64_t start;   /* first element of the chunk */
} dmc_opt_chunk;
typedef struct {
  float lr, beta1, beta2, eps, weight_decay;
  float bias_correction1, bias_correction2; /* 1 - beta^step, computed by the host */
  float max_norm;   /* > 0: gradients are scaled by min(1, max_norm / (norm + 1e-6)) as clip_grad_norm_ does; <= 0: no clipping */
  float ema_decay;  /* > 0: ema = decay * ema + (1 - decay) * new parameter for items with an ema pointer */
} dmc_adamw_desc;
/* norm_dev[0] = l2 norm over all gradients; partial_dev: fp32 [n_chunks] scratch (chunk sums added in index order) */
DMC_API int dmc_opt_grad_norm(const dmc_opt_item* items_dev, const dmc_opt_chunk* chunks_dev, int32_t n_chunks, float* partial_dev,
                      float* norm_dev, void* stream);
/* AdamW with decoupled weight decay (torch.optim.AdamW semantics) on the clipped gradient; norm_dev may be NULL when max_norm <= 0 */
DMC_API int dmc_opt_adamw_step(const dmc_opt_item* items_dev, const dmc_opt_chunk* chunks_dev, int32_t n_chunks,
                       const dmc_adamw_desc* h, const float* norm_dev, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Denoiser forward as a "plan": an ordered list of kernel launches with all pointers, shapes and TMA
 * descriptors resolved once per (model, batch size, workspace).  One dmc_plan_run() == one
 * model(x, t, y) call of the reference (models/unet.py:243-292, models/dit.py:263-295).
 * ---------------------------------------------------------------------------------------------- */
typedef struct dmc_plan dmc_plan;

DMC_API int dmc_plan_create(dmc_plan** out);
DMC_API int dmc_plan_destroy(dmc_plan* p);
DMC_API int dmc_plan_run(dmc_plan* p, void* stream);
/* runs only op `op_index` (profiling / debugging aid: lets ncu capture one layer with realistic inputs in place) */
DMC_API int dmc_plan_run_op(dmc_plan* p, int32_t op_index, void* stream);
/* number of kernel launches (and memsets) one dmc_plan_run() performs */
DMC_API int dmc_plan_num_launches(const dmc_plan* p);
/* algorithmic tensor-core FLOPs (2*M*N*K summed over GEMM-shaped ops, real channels only) per run */
DMC_API double dmc_plan_gemm_flops(const dmc_plan* p);
/* Re-point one external binding of op `op_index` (returned by dmc_plan_add_*):
 *   which = 0: primary input  (stem / patch_embed: x, cond / dit_cond: t)     which = 1: secondary input (cond: y, may be NULL)
 *   which = 2: primary output (conv: out_f32_nchw, head: out)                                                     */
DMC_API int dmc_plan_rebind(dmc_plan* p, int32_t op_index, int32_t which, const void* ptr);
/* new dropout seed of GroupNorm pass `op_index` (training: a fresh mask every step, models/unet.py:53) */
DMC_API int dmc_plan_set_seed(dmc_plan* p, int32_t op_index, uint32_t seed);
/* Per-op device timing: runs every op `iters` times between CUDA events on `stream`, writes the average
 * milliseconds per op into ms_out[num ops].  Debug / profiling aid used by bench.py's roofline leg. */
DMC_API int dmc_plan_time_ops(dmc_plan* p, void* stream, int32_t iters, float* ms_out, int32_t n_out);
DMC_API int dmc_plan_num_ops(const dmc_plan* p);
/* kind of op i: 0 memset, 1 cond, 2 stem, 3 gn_stats, 4 gn_apply, 5 conv, 6 attention, 7 upsample, 8 ddim, 9 ddpm,
 * 10 dit_cond, 11 patch_embed, 12 ln_modulate, 13 head */
DMC_API int dmc_plan_op_kind(const dmc_plan* p, int32_t i);
DMC_API double dmc_plan_op_flops(const dmc_plan* p, int32_t i);
DMC_API double dmc_plan_op_bytes(const dmc_plan* p, int32_t i);

DMC_API int dmc_plan_add_memset(dmc_plan* p, void* ptr, size_t bytes);

/* Conditioning table.  Replaces TimeEmbedding + time_embed (models/unet.py:18-25,167-172), the label
 * lookup (:256-260) and the 22 time_mlp / label_proj projections (:40-48,65-68):
 *   cond[n, :] = bt_all + Wt_all . SiLU(time_embed(t_n))  +  ytab[clamp(y_n, 0, num_classes), :]
 * freqs: the [half] sinusoid frequencies computed on the host with the reference's expression. */
typedef struct {
  const int64_t* t;     /* [B] */
  const int64_t* y;     /* [B] or NULL */
  int32_t B;
  int32_t uniform_t;    /* 1: all t_n equal (sampling) -> the time part is computed once */
  int32_t num_classes;  /* label clamp upper bound; ignored when ytab == NULL */
  int32_t half;         /* model_channels / 2 */
  int32_t temb;         /* model_channels * 4 */
  int32_t ncols;        /* sum of Cout over all residual blocks */
  const float* freqs;   /* [half] */
  const float* w1;      /* [temb, 2*half]  time_embed.1.weight */
  const float* b1;      /* [temb] */
  const float* w2;      /* [temb, temb]    time_embed.3.weight */
  const float* b2;      /* [temb] */
  const float* wt_all;  /* [ncols, temb]   all time_mlp.1.weight stacked */
  const float* bt_all;  /* [ncols]         time_mlp.1.bias + conv1 bias */
  const float* ytab;    /* [num_classes+1, ncols] label_proj(SiLU(label_embed)) per class, or NULL */
  float* scratch;       /* [2 * R * temb + R * ncols], R = uniform_t ? 1 : B */
  float* cond;          /* [B, ncols] */
} dmc_cond_desc;
DMC_API int dmc_plan_add_cond(dmc_plan* p, const dmc_cond_desc* d);

/* Input convolution 3x3, Cin <= 4, straight from the fp32 NCHW model input to bf16 NHWC
 * (models/unet.py:188,263).  x has x_batch images; image n reads x[n % x_batch] (CFG runs the cond and
 * uncond halves of one 2B batch over the same x). */
typedef struct {
  const float* x;      /* [x_batch, Cin, H, W] fp32 */
  int32_t x_batch, B, Cin, H, W, Cout;
  const float* weight; /* [Cout, Cin, 3, 3] fp32 */
  const float* bias;   /* [Cout] */
  void* out;           /* bf16 [B, H, W, Cout] */
  void* out_lo;        /* optional: bf16 rounding remainder (value - bf16(value)), see "split-bf16 mode" below */
} dmc_stem_desc;
DMC_API int dmc_plan_add_stem(dmc_plan* p, const dmc_stem_desc* d);

/* Input convolution, tensor-core form: this op only gathers the 3x3 neighbourhood of every pixel of the fp32 NCHW input into
 * one 64-channel bf16 NHWC row -- columns [0, 9*Cin): the taps, tap-major / channel-minor, as bf16(v); columns [9*Cin, 18*Cin):
 * the rounding remainders bf16(v - bf16(v)), so the pair carries 16 mantissa bits of the input; the rest zero -- and the
 * convolution itself is then ONE 1x1 tcgen05 GEMM (dmc_plan_add_conv, K = 64) against [W | W | 0] with bias, GroupNorm partial
 * sums and everything else the conv epilogue offers (models/unet.py:188,263).  Needs 18 * Cin <= 64. */
typedef struct {
  const float* x;      /* [x_batch, Cin, H, W] fp32; image n reads x[n % x_batch] */
  int32_t x_batch, B, Cin, H, W;
  void* out;           /* bf16 [B, H, W, 64] */
} dmc_stem_cols_desc;
DMC_API int dmc_plan_add_stem_cols(dmc_plan* p, const dmc_stem_cols_desc* d);

/* Split-bf16 mode ("bf16x3", the fp32-accuracy mode of the UNet): every activation is the SUM of two bf16 tensors, hi =
 * bf16(v) and lo = bf16(v - hi) (16 mantissa bits together), and every convolution is three bf16 tensor-core products
 * hi*W_hi + lo*W_hi + hi*W_lo accumulated in fp32 -- expressed with the ordinary multi-source convolution below as the
 * K-concatenation of the sources (hi, lo, hi) against the weight columns [W_hi | W_hi | W_lo].  The `*_lo` members of the
 * descriptors carry the low parts; they are NULL in the bf16 mode.
 *
 * GroupNorm statistics as deterministic partial sums.  A statistics buffer is fp32 [B, slots, C/8, 2]: (sum, sumsq)
 * of one image over one 8-channel block and one slice ("slot") of its pixels.  Every (image, slot, block) entry is
 * written exactly once with a plain store (no atomics, no pre-zeroing), and the consumer adds the slots in index
 * order -- so results are bit-reproducible and independent of what else is in the batch.
 * This stand-alone kernel writes slots = ceil(HW / 128) (one per 128-pixel slab). */
typedef struct {
  const void* src; /* bf16 [B, HW, C] */
  int32_t B, HW, C;
  float* stats;    /* [B, ceil(HW/128), C/8, 2] */
  const void* src_lo; /* optional low part: the tensor value is src + src_lo */
} dmc_gn_stats_desc;
DMC_API int dmc_plan_add_gn_stats(dmc_plan* p, const dmc_gn_stats_desc* d);

/* GroupNorm(groups, C) (+ SiLU) over the channel concatenation of up to two sources, written as ONE
 * bf16 NHWC tensor (models/unet.py:35-36,51-52,80,238-239 and the torch.cat of :284). */
typedef struct {
  int32_t nsrc;
  const void* src[2];     /* bf16 [B, HW, c_i] */
  int32_t src_c[2];
  const float* stats[2];  /* [B, stats_slots[i], c_i/8, 2] */
  int32_t stats_slots[2];
  int32_t B, HW, groups;
  const float* gamma;     /* [C] */
  const float* beta;      /* [C] */
  float eps;
  int32_t silu;
  void* out;              /* bf16 [B, HW, C] */
  const void* src_lo[2];  /* optional low parts of the sources (value = src + src_lo) */
  void* out_lo;           /* optional low part of the output */
  float drop_p;           /* training: dropout after SiLU (models/unet.py:53), counter-based mask from `seed`; 0 = none */
  uint32_t seed;
  const uint32_t* seed_dev; /* optional: a per-step seed in DEVICE memory, added to `seed` when the kernel runs -- a captured
                               CUDA graph of the step then draws a new mask on every replay */
} dmc_gn_apply_desc;
DMC_API int dmc_plan_add_gn_apply(dmc_plan* p, const dmc_gn_apply_desc* d);

/* Convolution as an implicit GEMM on tcgen05 tensor cores: M = B*Hout*Wout pixels, N = Cout,
 * K = sum_i taps_i * c_i.  Replaces nn.Conv2d 3x3 / 1x1 (models/unet.py:37,54,58,81,82,106,116,240) with
 * bias, conditioning add (:65-68), residual add (:72,:99) and the 1x1 shortcut (:58, as extra K columns)
 * fused.  Every source is bf16 NHWC [B, Hin, Win, c_i] with c_i % 64 == 0.
 *   taps_i = 9: 3x3 window, padding 1 (TMA zero fill), stride `stride`
 *   taps_i = 1: centre tap only (1x1 conv / fused shortcut), sampled at stride `stride`
 *   taps_i = 4: one phase (up_phase = 2*ph+pw) of "nearest 2x upsample then 3x3" (:118-120) computed on the
 *               low-res tensor as a 2x2 window; the output pixel (i, j) is scattered to (2i+ph, 2j+pw).
 * weight: bf16 [Cout_pad, Ktot] row-major, K ordered source by source, tap-major, channel-minor. */
typedef struct {
  int32_t nsrc;
  const void* src[3];
  int32_t src_c[3];
  int32_t src_taps[3];
  int32_t B, Hin, Win;
  int32_t stride;   /* 1 or 2 */
  int32_t up_phase; /* -1, or 0..3 */
  const void* weight;
  int32_t Cout, Cout_pad, Ktot;
  const float* bias;      /* [Cout] or NULL */
  const float* cond;      /* per-image addend row pointer base (already offset to this layer's column) or NULL */
  int32_t cond_stride;    /* floats between images */
  const void* residual;   /* bf16, same shape as out_bf16, or NULL */
  void* out_bf16;         /* bf16 NHWC [B, Hout*, Wout*, Cout] or NULL */
  float* out_f32_nchw;    /* fp32 [B, Cout, Hout, Wout] or NULL (model output head) */
  float* stats;           /* optional [B, stats_slots, Cout/8, 2] partial sums of the OUTPUT (see dmc_gn_stats_desc) */
  int32_t stats_slots;    /* must equal max(1, P/32) * (up_phase >= 0 ? 4 : 1), P = iteration pixels per image
                             (Hin/stride * Win/stride): one slot per 32-pixel epilogue warp, per phase */
  int32_t impl;           /* 0: tcgen05/TMA kernel (product path)   1: CUDA-core debug kernel (tests only, no stats) */
  /* --- transformer (DiT) epilogues: v = acc + bias (+ cond); v = act(v); v *= gate[n, c]; v += residual; store --- */
  int32_t act;                /* 0 none, 1 GELU (erf form, nn.GELU() default; models/dit.py:97) */
  const float* gate;          /* per-image per-channel multiplier (adaLN gate, models/dit.py:124,130), row n at
                                 gate + n * gate_stride; NULL = none */
  int32_t gate_stride;        /* floats between images */
  const float* residual_f32;  /* fp32 NHWC residual stream [B, Hout, Wout, Cout] or NULL (exclusive with `residual`) */
  float* out_f32_nhwc;        /* fp32 NHWC output (may alias residual_f32: each element is read then written by the same
                                 thread) or NULL */
  void* out_lo;               /* optional low part of out_bf16 (split-bf16 mode; per-thread stores, no TMA epilogue) */
  const void* residual_lo;    /* optional low part of `residual` */
  int32_t unpatch_p;          /* > 0 with out_f32_nchw: columns are (pi, qi, c) of a p x p patch and pixel (i, j) of image n
                                 scatters to out[n, c, i*p + pi, j*p + qi] -- DiT.unpatchify (models/dit.py:249-261);
                                 Cout = p * p * channels */
  /* --- GroupNorm(+SiLU) of the OUTPUT fused into this convolution's epilogue (models/unet.py:35-36,51-52,80,238-239: the
   * normalisation the NEXT layer applies to this tensor).  The epilogue keeps the finished fp32 tile in tensor memory, writes
   * the GroupNorm partial sums of the output (`stats`, required), waits until every CTA holding a piece of the same image has
   * done so (shared-memory barrier inside a CTA, a self-resetting counter in `gn_counters` across CTAs), reduces them to
   * mean / rstd in slot order (bit-reproducible, batch-invariant) and writes up to two normalised versions
   *     gn_out[v][n, pixel, gn_coff[v] + c] = act_v((x - mean_g) * rstd_g * gn_gamma[v][c] + gn_beta[v][c])
   * straight from the fp32 accumulator.  A version may be a channel slice of a wider tensor (the torch.cat of
   * models/unet.py:284 is then free): gn_pitch[v] is that tensor's channel count, gn_gsize[v] the channels per statistics group
   * (C_total / 8 of the consumer's GroupNorm; groups must not straddle this tensor).  out_bf16 may be NULL when nothing reads
   * the raw output.  Supported when dmc_conv_gn_supported() says so. --- */
  int32_t gn_nver;            /* 0 (no fused GroupNorm), 1 or 2 */
  void* gn_out[2];            /* bf16 NHWC [B, Hout, Wout, gn_pitch[v]] */
  int32_t gn_pitch[2];
  int32_t gn_coff[2];         /* first channel of this tensor inside gn_out[v] */
  const float* gn_gamma[2];   /* [Cout]: the consumer GroupNorm's weight / bias rows of THIS tensor's channels */
  const float* gn_beta[2];
  int32_t gn_gsize[2];        /* 16, 32 or 64 */
  int32_t gn_silu[2];
  float gn_eps;
  int32_t* gn_counters;       /* int32 [2 * B * (Cout / 32)] zeroed once by the caller (only touched when an image spans CTAs) */
  /* --- GroupNorm WITHOUT activation applied to the INPUT of a 1x1 convolution (the `norm` -> `qkv` pair of AttentionBlock,
   * models/unet.py:80-81,86-87): src[0] is the raw tensor and a_affine[n, c] = (scale, shift) from dmc_plan_add_gn_coeff; two
   * otherwise idle warps rewrite every A tile in shared memory as bf16(x * scale + shift) before the MMAs read it -- bit-identical
   * to the stand-alone GroupNorm pass, which disappears.  One source, one tap, resident weights (dmc_conv_affine_supported). --- */
  const float* a_affine;      /* fp32 [B, src_c[0], 2] or NULL */
} dmc_conv_desc;
DMC_API int dmc_plan_add_conv(dmc_plan* p, const dmc_conv_desc* d);
/* 1 when a 1x1 convolution of this geometry runs with resident weights, i.e. can take a_affine */
DMC_API int dmc_conv_affine_supported(int32_t B, int32_t H, int32_t W, int32_t Cin, int32_t Cout);
/* per-image per-channel (scale, shift) of GroupNorm(groups, C) from the partial sums of a tensor: out[n, c] = (rstd_g * gamma_c,
 * beta_c - mean_g * rstd_g * gamma_c), same summation order as dmc_plan_add_gn_apply */
typedef struct {
  const float* stats;  /* [B, stats_slots, C/8, 2] */
  int32_t stats_slots, B, HW, C, groups;
  const float* gamma;
  const float* beta;
  float eps;
  float* out;          /* fp32 [B, C, 2] */
} dmc_gn_coeff_desc;
DMC_API int dmc_plan_add_gn_coeff(dmc_plan* p, const dmc_gn_coeff_desc* d);
/* 1 when some tile configuration of the tcgen05 kernel can fuse the GroupNorm of a [B, Hout, Wout, Cout] output whose largest
 * statistics group has max_gsize channels (pure geometry: no pointers, no device), else 0 (use dmc_plan_add_gn_apply) */
DMC_API int dmc_conv_gn_supported(int32_t B, int32_t Hout, int32_t Wout, int32_t Cout, int32_t max_gsize);

/* Output head, fused: GroupNorm(groups, C) + SiLU + 3x3 convolution C -> Cout (<= 8) + bias, written as the fp32 NCHW
 * model output (models/unet.py:237-241,287-292).  One read of the activation; replaces a gn_apply + a padded-N conv. */
typedef struct {
  const void* src;      /* bf16 [B, H, W, C] (raw, un-normalised) */
  const float* stats;   /* [B, stats_slots, C/8, 2] partial sums of src */
  int32_t stats_slots;
  int32_t B, H, W, C, Cout, groups;
  const float* gamma;   /* [C] */
  const float* beta;    /* [C] */
  float eps;
  const float* weight;  /* fp32 [Cout, C, 3, 3] (the reference's layout) */
  const float* bias;    /* [Cout] */
  float* out;           /* fp32 [B, Cout, H, W] */
  void* wfrag;          /* caller-owned scratch, 9 * (C/16) * 256 bytes: the weights re-packed as tensor-core fragments */
} dmc_head_desc;
DMC_API int dmc_plan_add_head(dmc_plan* p, const dmc_head_desc* d);
/* 1 when the fused head kernel covers this shape (C % 16 == 0, C <= 256, W % 16 == 0, W <= 64, Cout <= 8) */
DMC_API int dmc_head_supported(const dmc_head_desc* d);

/* Multi-head self-attention core softmax(Q K^T / sqrt(hd)) V over L tokens (models/unet.py:88-96;
 * models/dit.py:94,123).  qkv: bf16 [B, L, 3*C], channel order [q|k|v][head][hd]; out: bf16 [B, L, C]. */
typedef struct {
  const void* qkv;
  void* out;
  int32_t B, L, heads, C;
  int32_t impl; /* 0: tcgen05/TMA kernel (head dim 64, L in {16, 32, 64, 128, 256}; other shapes run the CUDA-core
                   flash kernel)   1: force the CUDA-core flash kernel (tests / debugging) */
  const void* qkv_lo; /* optional low parts (split-bf16 mode): fp32 CUDA-core kernel on qkv + qkv_lo, writes out + out_lo */
  void* out_lo;
} dmc_attn_desc;
DMC_API int dmc_plan_add_attention(dmc_plan* p, const dmc_attn_desc* d);

/* ---- DiT (models/dit.py) -------------------------------------------------------------------------------------- */

/* Conditioning of every adaLN layer at once.  Replaces TimestepEmbedder (models/dit.py:42-55), the label lookup
 * (:278-283) and the depth+1 adaLN_modulation linears (:106-109,115-116,142-147):
 *   c_n   = W2 . SiLU(W1 . [cos(t_n f) | sin(t_n f)] + b1) + b2  (+ emb[clamp(y_n, 0, num_classes)])
 *   mod_n = b_all + W_all . SiLU(c_n)              W_all: all adaLN_modulation.1 weights stacked, [ncols, hidden]
 * With uniform_t (sampling: every t_n equal) only num_classes + 1 distinct rows exist; they are computed once and
 * gathered per image. */
typedef struct {
  const int64_t* t;    /* [B] */
  const int64_t* y;    /* [B] or NULL */
  int32_t B, uniform_t, num_classes;
  int32_t freq_dim;    /* 256 */
  int32_t hidden, ncols;
  const float* freqs;  /* [freq_dim / 2], the reference's expression evaluated on the host */
  const float* w1;     /* [hidden, freq_dim]  t_embedder.mlp.0 */
  const float* b1;
  const float* w2;     /* [hidden, hidden]    t_embedder.mlp.2 */
  const float* b2;
  const float* emb;    /* [num_classes + 1, hidden] y_embedder.embedding_table.weight, or NULL */
  const float* w_all;  /* [ncols, hidden] */
  const float* b_all;  /* [ncols] */
  float* scratch;      /* [R * (2 * hidden + ncols)], R = uniform_t ? (emb ? num_classes + 1 : 1) : B */
  float* mod;          /* [B, ncols] */
} dmc_dit_cond_desc;
DMC_API int dmc_plan_add_dit_cond(dmc_plan* p, const dmc_dit_cond_desc* d);

/* PatchEmbed + positional embedding (models/dit.py:23-27,265): non-overlapping p x p conv straight from the fp32 NCHW
 * input to the fp32 token stream tok[n, i*Wt + j, :] = bias + W . patch(i, j) + pos[i*Wt + j, :]. Image n reads
 * x[n % x_batch] (CFG: both halves of a 2B batch share x). */
typedef struct {
  const float* x;      /* [x_batch, Cin, H, W] */
  int32_t x_batch, B, Cin, H, W, patch, hidden;
  const float* weight; /* TRANSPOSED x_embedder.proj.weight: [Cin * patch * patch, hidden], k = (ci * patch + pi) * patch + qi */
  const float* bias;   /* [hidden] */
  const float* pos;    /* [(H/patch) * (W/patch), hidden] */
  float* out;          /* fp32 [B, (H/patch) * (W/patch), hidden] */
} dmc_patch_embed_desc;
DMC_API int dmc_plan_add_patch_embed(dmc_plan* p, const dmc_patch_embed_desc* d);

/* LayerNorm (no affine, eps) over the channel axis of the fp32 token stream, then adaLN modulation, written as the bf16
 * GEMM operand: out[n, l, :] = LN(x[n, l, :]) * (1 + scale[n, :]) + shift[n, :]   (models/dit.py:119-120,127-128,148-149).
 * shift / scale: row n at ptr + n * mod_stride (columns of the dmc_dit_cond table). */
typedef struct {
  const float* x;      /* fp32 [B, L, C] */
  void* out;           /* bf16 [B, L, C] */
  int32_t B, L, C;
  const float* shift;
  const float* scale;
  int32_t mod_stride;
  float eps;
  void* out_lo;        /* optional low part of the output (split-bf16 mode) */
} dmc_ln_mod_desc;
DMC_API int dmc_plan_add_ln_modulate(dmc_plan* p, const dmc_ln_mod_desc* d);

/* nearest-neighbour 2x upsample of a bf16 NHWC tensor (models/unet.py:119) */
typedef struct {
  const void* src; /* [B, H, W, C] */
  void* out;       /* [B, 2H, 2W, C] */
  int32_t B, H, W, C;
} dmc_upsample_desc;
DMC_API int dmc_plan_add_upsample(dmc_plan* p, const dmc_upsample_desc* d);

/* Output head, second half (models/unet.py:240: conv3x3 C -> Cout, Cout <= 3..10, fp32 NCHW eps).  A 3x3 convolution with so
 * few output channels is bound by re-reading its input 3 - 9 times; instead the head runs as
 *   y[n, p, tap * Cout + co] = sum_c a[n, p, c] * w[co, c, tap]      one 1x1 tcgen05 GEMM: every input pixel is read ONCE
 *   out[n, co, i, j] = bias[co] + sum_tap y[n, (i + dh_tap, j + dw_tap), tap * Cout + co]   (this op; zero outside the image)
 * y is fp32 [B, H, W, ypitch] (ypitch >= 9 * Cout, the GEMM's padded column count), so the nine partial sums are added in
 * fp32 exactly like the accumulator of the direct convolution. */
typedef struct {
  const float* y;    /* fp32 [B, H, W, ypitch] */
  int32_t B, H, W, Cout, ypitch;
  const float* bias; /* [Cout] or NULL */
  float* out;        /* fp32 [B, Cout, H, W] */
} dmc_head_taps_desc;
DMC_API int dmc_plan_add_head_taps(dmc_plan* p, const dmc_head_taps_desc* d);

/* Scheduler steps as plan ops, so that "model forward + update" chains (and whole sampling loops) can be
 * replayed with a single call / captured in one CUDA graph. */
typedef struct {
  const float* x;
  const float* eps_c;
  const float* eps_u;
  const float* noise;
  float* x_out;
  int32_t B, n_per_sample;
  const void* coef_dev; /* dmc_ddim_coef* or dmc_ddpm_coef* (device) */
  dmc_guidance g;
  const int32_t* step_index_dev; /* optional: use row coef_dev[step_index_dev[0]] */
} dmc_step_desc;
DMC_API int dmc_plan_add_ddim_step(dmc_plan* p, const dmc_step_desc* d);
DMC_API int dmc_plan_add_ddpm_step(dmc_plan* p, const dmc_step_desc* d);

#ifdef __cplusplus
}
#endif
#endif /* DMC_H_ */
