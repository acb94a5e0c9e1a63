// C ABI (include/dmc.h): error plumbing, dmc_init, stateless scheduler entry points, and the "plan" interpreter that
// replays one denoiser forward (or a forward + scheduler-step chain) as a fixed list of kernel launches.
#include <stdarg.h>
#include <string.h>

#include <vector>

#include "common.cuh"
#include "kernels.h"

namespace dmc {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

static int g_num_sms = 0;
static EncodeTiledFn g_encode = nullptr;

int num_sms() { return g_num_sms > 0 ? g_num_sms : 148; }
EncodeTiledFn encode_tiled_fn() { return g_encode; }

enum OpKind { OP_MEMSET = 0, OP_COND, OP_STEM, OP_GN_STATS, OP_GN_APPLY, OP_CONV, OP_ATTN, OP_UPSAMPLE, OP_DDIM, OP_DDPM,
              OP_DIT_COND, OP_PATCH_EMBED, OP_LN_MOD, OP_HEAD, OP_HEAD_TAPS, OP_STEM_COLS, OP_GN_COEFF };

struct Op {
  int kind;
  union {
    struct { void* ptr; size_t bytes; } memset_;
    dmc_cond_desc cond;
    dmc_stem_desc stem;
    dmc_gn_stats_desc gn_stats;
    dmc_gn_apply_desc gn_apply;
    dmc_conv_desc conv;
    dmc_attn_desc attn;
    dmc_upsample_desc up;
    dmc_step_desc step;
    dmc_dit_cond_desc dit_cond;
    dmc_patch_embed_desc patch;
    dmc_ln_mod_desc ln_mod;
    dmc_head_desc head;
    dmc_head_taps_desc head_taps;
    dmc_stem_cols_desc stem_cols;
    dmc_gn_coeff_desc gn_coeff;
  };
  ConvPrepared* conv_prep;
  AttnPrepared* attn_prep;
  double flops;  // algorithmic tensor FLOPs (GEMM-shaped ops)
  double bytes;  // algorithmic HBM bytes (read inputs once + write outputs once)
  Op() {
    memset(static_cast<void*>(this), 0, sizeof(*this));
    kind = -1;
  }
};

}  // namespace dmc

struct dmc_plan {
  std::vector<dmc::Op> ops;
};

namespace dmc {

static int run_op(const Op& op, cudaStream_t st) {
  switch (op.kind) {
    case OP_MEMSET: DMC_CUDA_OK(cudaMemsetAsync(op.memset_.ptr, 0, op.memset_.bytes, st)); return 0;
    case OP_COND: return launch_cond(op.cond, st);
    case OP_STEM: return launch_stem(op.stem, st);
    case OP_GN_STATS: return launch_gn_stats(op.gn_stats, st);
    case OP_GN_APPLY: return launch_gn_apply(op.gn_apply, st);
    case OP_CONV: return op.conv.impl == 1 ? launch_conv_ref(op.conv, st) : launch_conv(op.conv, op.conv_prep, st);
    case OP_ATTN: return op.attn_prep ? launch_attention_umma(op.attn_prep, st) : launch_attention(op.attn, st);
    case OP_UPSAMPLE: return launch_upsample(op.up, st);
    case OP_DDIM: return launch_step(false, op.step, st);
    case OP_DDPM: return launch_step(true, op.step, st);
    case OP_DIT_COND: return launch_dit_cond(op.dit_cond, st);
    case OP_PATCH_EMBED: return launch_patch_embed(op.patch, st);
    case OP_LN_MOD: return launch_ln_modulate(op.ln_mod, st);
    case OP_HEAD: return launch_head_fused(op.head, st);
    case OP_HEAD_TAPS: return launch_head_taps(op.head_taps, st);
    case OP_STEM_COLS: return launch_stem_cols(op.stem_cols, st);
    case OP_GN_COEFF: return launch_gn_coeff(op.gn_coeff, st);
  }
  set_error("plan: unknown op kind %d", op.kind);
  return -1;
}

}  // namespace dmc

using namespace dmc;

extern "C" {

const char* dmc_last_error(void) { return g_err; }
int dmc_abi_version(void) { return DMC_ABI_VERSION; }

int dmc_init(void) {
  int dev = 0;
  DMC_CUDA_OK(cudaGetDevice(&dev));
  cudaDeviceProp prop;
  DMC_CUDA_OK(cudaGetDeviceProperties(&prop, dev));
  DMC_REQUIRE(prop.major == 10, "dmc_init: device %d is sm_%d%d; this library contains sm_100a code only", dev, prop.major,
              prop.minor);
  g_num_sms = prop.multiProcessorCount;
  if (g_encode == nullptr) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    DMC_CUDA_OK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
    DMC_REQUIRE(fn != nullptr && qres == cudaDriverEntryPointSuccess, "dmc_init: cuTensorMapEncodeTiled not found in the driver");
    g_encode = reinterpret_cast<EncodeTiledFn>(fn);
  }
  return g_num_sms;
}

int dmc_ddim_step(const float* x, const float* eps_c, const float* eps_u, const float* noise, float* x_out, int32_t B,
                  int32_t n_per_sample, const dmc_ddim_coef* coef_dev, const dmc_guidance* g, void* stream) {
  DMC_REQUIRE(g != nullptr, "dmc_ddim_step: guidance is null");
  dmc_step_desc d{x, eps_c, eps_u, noise, x_out, B, n_per_sample, coef_dev, *g, nullptr};
  return launch_step(false, d, static_cast<cudaStream_t>(stream));
}

int dmc_ddim_step_at(const float* x, const float* eps_c, const float* eps_u, const float* noise, float* x_out, int32_t B,
                     int32_t n_per_sample, const dmc_ddim_coef* coef_table_dev, const int32_t* step_index_dev,
                     const dmc_guidance* g, void* stream) {
  DMC_REQUIRE(g != nullptr && step_index_dev != nullptr, "dmc_ddim_step_at: null guidance / step index");
  dmc_step_desc d{x, eps_c, eps_u, noise, x_out, B, n_per_sample, coef_table_dev, *g, step_index_dev};
  return launch_step(false, d, static_cast<cudaStream_t>(stream));
}

int dmc_ddpm_step_at(const float* x, const float* eps_c, const float* eps_u, const float* noise, float* x_out, int32_t B,
                     int32_t n_per_sample, const dmc_ddpm_coef* coef_table_dev, const int32_t* step_index_dev,
                     const dmc_guidance* g, void* stream) {
  DMC_REQUIRE(g != nullptr && step_index_dev != nullptr, "dmc_ddpm_step_at: null guidance / step index");
  dmc_step_desc d{x, eps_c, eps_u, noise, x_out, B, n_per_sample, coef_table_dev, *g, step_index_dev};
  return launch_step(true, d, static_cast<cudaStream_t>(stream));
}

int dmc_advance(int32_t* counter_dev, const int64_t* t_table_dev, int64_t* t_out_dev, int32_t n, void* stream) {
  return launch_advance(counter_dev, t_table_dev, t_out_dev, n, static_cast<cudaStream_t>(stream));
}

int dmc_ddpm_step(const float* x, const float* eps_c, const float* eps_u, const float* noise, float* x_out, int32_t B,
                  int32_t n_per_sample, const dmc_ddpm_coef* coef_dev, const dmc_guidance* g, void* stream) {
  DMC_REQUIRE(g != nullptr, "dmc_ddpm_step: guidance is null");
  dmc_step_desc d{x, eps_c, eps_u, noise, x_out, B, n_per_sample, coef_dev, *g, nullptr};
  return launch_step(true, d, static_cast<cudaStream_t>(stream));
}

int dmc_q_sample(const float* x0, const float* noise, const int64_t* t, const float* sqrt_acp, const float* sqrt_1m_acp,
                 float* x_t, int32_t B, int32_t n_per_sample, void* stream) {
  return launch_q_sample(x0, noise, t, sqrt_acp, sqrt_1m_acp, x_t, B, n_per_sample, static_cast<cudaStream_t>(stream));
}

int dmc_conv_wgrad_splits(const dmc_wgrad_desc* d) {
  DMC_REQUIRE(d != nullptr && d->stride >= 1 && d->Cout >= 128 && d->Cin >= 64, "dmc_conv_wgrad_splits: bad descriptor");
  return conv_wgrad_splits(*d);
}

int dmc_conv_wgrad(const dmc_wgrad_desc* d, void* stream) {
  DMC_REQUIRE(d != nullptr, "dmc_conv_wgrad: null descriptor");
  return launch_conv_wgrad(*d, static_cast<cudaStream_t>(stream));
}

int dmc_gn_backward(const dmc_gn_bwd_desc* d, void* stream) {
  DMC_REQUIRE(d != nullptr, "dmc_gn_backward: null descriptor");
  return launch_gn_backward(*d, static_cast<cudaStream_t>(stream));
}
int dmc_attention_backward(const dmc_attn_bwd_desc* d, void* stream) {
  DMC_REQUIRE(d != nullptr, "dmc_attention_backward: null descriptor");
  return launch_attention_backward(*d, static_cast<cudaStream_t>(stream));
}
int dmc_channel_sum(const void* src, float* out, int32_t B, int32_t HW, int32_t C, int32_t per_image, int32_t accumulate,
                    float* scratch, void* stream) {
  return launch_channel_sum(src, out, B, HW, C, per_image, accumulate, scratch, static_cast<cudaStream_t>(stream));
}
int dmc_dit_gate_ln_mod(const dmc_dit_glm_desc* d, void* stream) {
  DMC_REQUIRE(d != nullptr, "dmc_dit_gate_ln_mod: null descriptor");
  return launch_dit_gate_ln_mod(*d, static_cast<cudaStream_t>(stream));
}
int dmc_dit_gate_ln_mod_backward(const dmc_dit_glm_bwd_desc* d, void* stream) {
  DMC_REQUIRE(d != nullptr, "dmc_dit_gate_ln_mod_backward: null descriptor");
  return launch_dit_gate_ln_mod_backward(*d, static_cast<cudaStream_t>(stream));
}
int dmc_gelu_forward(const void* u, void* m, int64_t n, float drop_p, uint32_t seed, void* stream) {
  return launch_gelu_forward(u, m, n, drop_p, seed, static_cast<cudaStream_t>(stream));
}
int dmc_gelu_backward(const void* u, const void* dm, void* du, int64_t n, float drop_p, uint32_t seed, void* stream) {
  return launch_gelu_backward(u, dm, du, n, drop_p, seed, static_cast<cudaStream_t>(stream));
}
int dmc_dilate2x(const void* src, void* dst, int32_t B, int32_t h, int32_t w, int32_t C, void* stream) {
  return launch_dilate2x(src, dst, B, h, w, C, static_cast<cudaStream_t>(stream));
}
int dmc_pack_weights(const dmc_pack_item* items_dev, int32_t n_items, void* stream) {
  return launch_pack_weights(items_dev, n_items, static_cast<cudaStream_t>(stream));
}
int64_t dmc_gn_backward_scratch(const dmc_gn_bwd_desc* d) {
  if (!d) return -1;
  return gn_backward_scratch_floats(*d);
}
int dmc_opt_grad_norm(const dmc_opt_item* items_dev, const dmc_opt_chunk* chunks_dev, int32_t n_chunks, float* partial_dev,
                      float* norm_dev, void* stream) {
  return launch_opt_grad_norm(items_dev, chunks_dev, n_chunks, partial_dev, norm_dev, static_cast<cudaStream_t>(stream));
}
int dmc_opt_adamw_step(const dmc_opt_item* items_dev, const dmc_opt_chunk* chunks_dev, int32_t n_chunks, const dmc_adamw_desc* h,
                       const float* norm_dev, void* stream) {
  DMC_REQUIRE(h != nullptr, "dmc_opt_adamw_step: null hyper-parameters");
  return launch_opt_adamw(items_dev, chunks_dev, n_chunks, *h, norm_dev, static_cast<cudaStream_t>(stream));
}
int dmc_add_bf16(void* dst, const void* src, int64_t n, int32_t accumulate, void* stream) {
  DMC_REQUIRE(n > 0, "dmc_add_bf16: n=%lld", static_cast<long long>(n));
  return launch_add_bf16(dst, src, static_cast<size_t>(n), accumulate, static_cast<cudaStream_t>(stream));
}
int dmc_block_sum2x2(const void* dhigh, void* dlow, int32_t B, int32_t H, int32_t W, int32_t C, int32_t accumulate, void* stream) {
  return launch_block_sum2x2(dhigh, dlow, B, H, W, C, accumulate, static_cast<cudaStream_t>(stream));
}
int dmc_nchw_f32_to_nhwc_bf16(const float* src, void* dst, int32_t B, int32_t Csrc, int32_t HW, int32_t Cdst, void* stream) {
  return launch_nchw_to_nhwc_pad(src, dst, B, Csrc, HW, Cdst, static_cast<cudaStream_t>(stream));
}
int dmc_conv_dgrad_strided(const void* dy, const float* w, void* dx, int32_t B, int32_t Hin, int32_t Win, int32_t Cin,
                           int32_t Cout, int32_t stride, int32_t accumulate, void* stream) {
  return launch_conv_dgrad_strided(dy, w, dx, B, Hin, Win, Cin, Cout, stride, accumulate, static_cast<cudaStream_t>(stream));
}

int dmc_plan_create(dmc_plan** out) {
  DMC_REQUIRE(out != nullptr, "dmc_plan_create: null out");
  *out = new (std::nothrow) dmc_plan();
  DMC_REQUIRE(*out != nullptr, "dmc_plan_create: out of memory");
  return 0;
}

int dmc_plan_destroy(dmc_plan* p) {
  if (p == nullptr) return 0;
  for (auto& op : p->ops)
    if (op.conv_prep) conv_release(op.conv_prep);
  for (auto& op : p->ops)
    if (op.attn_prep) attention_release(op.attn_prep);
  delete p;
  return 0;
}

int dmc_plan_run(dmc_plan* p, void* stream) {
  DMC_REQUIRE(p != nullptr, "dmc_plan_run: null plan");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  for (size_t i = 0; i < p->ops.size(); ++i) {
    int r = run_op(p->ops[i], st);
    if (r != 0) return r;
  }
  return 0;
}

int dmc_plan_run_op(dmc_plan* p, int32_t op_index, void* stream) {
  DMC_REQUIRE(p && op_index >= 0 && op_index < static_cast<int>(p->ops.size()), "dmc_plan_run_op: bad op index");
  return run_op(p->ops[op_index], static_cast<cudaStream_t>(stream));
}

int dmc_plan_num_ops(const dmc_plan* p) { return p ? static_cast<int>(p->ops.size()) : -1; }
int dmc_plan_op_kind(const dmc_plan* p, int32_t i) {
  DMC_REQUIRE(p && i >= 0 && i < static_cast<int>(p->ops.size()), "dmc_plan_op_kind: bad index");
  return p->ops[i].kind;
}
double dmc_plan_op_flops(const dmc_plan* p, int32_t i) {
  return (p && i >= 0 && i < static_cast<int>(p->ops.size())) ? p->ops[i].flops : 0.0;
}
double dmc_plan_op_bytes(const dmc_plan* p, int32_t i) {
  return (p && i >= 0 && i < static_cast<int>(p->ops.size())) ? p->ops[i].bytes : 0.0;
}

int dmc_plan_num_launches(const dmc_plan* p) {
  if (!p) return -1;
  int n = 0;
  for (const auto& op : p->ops)
    n += (op.kind == OP_HEAD) ? 2 : (op.kind == OP_COND) ? cond_num_launches(op.cond) : (op.kind == OP_DIT_COND ? dit_cond_num_launches(op.dit_cond) : 1);
  return n;
}

double dmc_plan_gemm_flops(const dmc_plan* p) {
  double f = 0;
  if (p)
    for (const auto& op : p->ops) f += op.flops;
  return f;
}

int dmc_plan_set_seed(dmc_plan* p, int32_t op_index, uint32_t seed) {
  DMC_REQUIRE(p && op_index >= 0 && op_index < static_cast<int>(p->ops.size()), "dmc_plan_set_seed: bad op index");
  Op& op = p->ops[op_index];
  DMC_REQUIRE(op.kind == OP_GN_APPLY, "dmc_plan_set_seed: op %d is not a GroupNorm pass", op_index);
  op.gn_apply.seed = seed;
  return 0;
}

int dmc_plan_rebind(dmc_plan* p, int32_t op_index, int32_t which, const void* ptr) {
  DMC_REQUIRE(p && op_index >= 0 && op_index < static_cast<int>(p->ops.size()), "dmc_plan_rebind: bad op index");
  Op& op = p->ops[op_index];
  if (op.kind == OP_STEM && which == 0) { op.stem.x = static_cast<const float*>(ptr); return 0; }
  if (op.kind == OP_STEM_COLS && which == 0) { op.stem_cols.x = static_cast<const float*>(ptr); return 0; }
  if (op.kind == OP_COND && which == 0) { op.cond.t = static_cast<const int64_t*>(ptr); return 0; }
  if (op.kind == OP_COND && which == 1) { op.cond.y = static_cast<const int64_t*>(ptr); return 0; }
  if (op.kind == OP_HEAD_TAPS && which == 2 && ptr != nullptr) { op.head_taps.out = static_cast<float*>(const_cast<void*>(ptr)); return 0; }
  if (op.kind == OP_HEAD && which == 2 && ptr != nullptr) { op.head.out = static_cast<float*>(const_cast<void*>(ptr)); return 0; }
  if (op.kind == OP_PATCH_EMBED && which == 0) { op.patch.x = static_cast<const float*>(ptr); return 0; }
  if (op.kind == OP_DIT_COND && which == 0) { op.dit_cond.t = static_cast<const int64_t*>(ptr); return 0; }
  if (op.kind == OP_DIT_COND && which == 1) { op.dit_cond.y = static_cast<const int64_t*>(ptr); return 0; }
  if (op.kind == OP_CONV && which == 2 && op.conv.out_f32_nchw != nullptr && ptr != nullptr) {
    op.conv.out_f32_nchw = static_cast<float*>(const_cast<void*>(ptr));
    return 0;
  }
  set_error("dmc_plan_rebind: op %d (kind %d) has no binding %d", op_index, op.kind, which);
  return -1;
}

int dmc_plan_time_ops(dmc_plan* p, void* stream, int32_t iters, float* ms_out, int32_t n_out) {
  DMC_REQUIRE(p && ms_out && iters > 0 && n_out >= static_cast<int>(p->ops.size()), "dmc_plan_time_ops: bad arguments");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  cudaEvent_t e0, e1;
  DMC_CUDA_OK(cudaEventCreate(&e0));
  DMC_CUDA_OK(cudaEventCreate(&e1));
  int rc = 0;
  for (size_t i = 0; i < p->ops.size() && rc == 0; ++i) {
    rc = run_op(p->ops[i], st);  // warm
    if (rc) break;
    cudaEventRecord(e0, st);
    for (int k = 0; k < iters && rc == 0; ++k) rc = run_op(p->ops[i], st);
    cudaEventRecord(e1, st);
    if (cudaEventSynchronize(e1) != cudaSuccess) { set_error("dmc_plan_time_ops: op %zu failed: %s", i, cudaGetErrorString(cudaGetLastError())); rc = -2; break; }
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    ms_out[i] = ms / iters;
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  return rc;
}

static int push(dmc_plan* p, const Op& op) {
  p->ops.push_back(op);
  return static_cast<int>(p->ops.size()) - 1;
}

int dmc_plan_add_memset(dmc_plan* p, void* ptr, size_t bytes) {
  DMC_REQUIRE(p && ptr && bytes > 0, "dmc_plan_add_memset: bad arguments");
  Op op;
  op.kind = OP_MEMSET;
  op.memset_.ptr = ptr;
  op.memset_.bytes = bytes;
  op.bytes = static_cast<double>(bytes);
  return push(p, op);
}

int dmc_plan_add_cond(dmc_plan* p, const dmc_cond_desc* d) {
  DMC_REQUIRE(p && d, "dmc_plan_add_cond: null argument");
  Op op;
  op.kind = OP_COND;
  op.cond = *d;
  op.bytes = 4.0 * (static_cast<double>(d->ncols) * d->temb + 2.0 * d->B * d->ncols);
  return push(p, op);
}

int dmc_plan_add_stem(dmc_plan* p, const dmc_stem_desc* d) {
  DMC_REQUIRE(p && d, "dmc_plan_add_stem: null argument");
  Op op;
  op.kind = OP_STEM;
  op.stem = *d;
  const double pix = static_cast<double>(d->B) * d->H * d->W;
  op.bytes = pix * (4.0 * d->Cin + 2.0 * d->Cout);
  op.flops = 2.0 * pix * d->Cout * d->Cin * 9;
  return push(p, op);
}

int dmc_plan_add_gn_stats(dmc_plan* p, const dmc_gn_stats_desc* d) {
  DMC_REQUIRE(p && d, "dmc_plan_add_gn_stats: null argument");
  Op op;
  op.kind = OP_GN_STATS;
  op.gn_stats = *d;
  op.bytes = 2.0 * d->B * d->HW * d->C;
  return push(p, op);
}

int dmc_plan_add_gn_apply(dmc_plan* p, const dmc_gn_apply_desc* d) {
  DMC_REQUIRE(p && d, "dmc_plan_add_gn_apply: null argument");
  Op op;
  op.kind = OP_GN_APPLY;
  op.gn_apply = *d;
  const int C = d->src_c[0] + (d->nsrc == 2 ? d->src_c[1] : 0);
  op.bytes = 2.0 * 2.0 * d->B * d->HW * C;
  return push(p, op);
}

int dmc_plan_add_conv(dmc_plan* p, const dmc_conv_desc* d) {
  DMC_REQUIRE(p && d, "dmc_plan_add_conv: null argument");
  Op op;
  op.kind = OP_CONV;
  op.conv = *d;
  if (d->impl == 0) {
    int r = conv_prepare(*d, &op.conv_prep);
    if (r != 0) return r;
  } else {
    DMC_REQUIRE(d->impl == 1, "dmc_plan_add_conv: impl=%d", d->impl);
  }
  const double opix = static_cast<double>(d->B) * (d->Hin / d->stride) * (d->Win / d->stride);
  op.flops = 2.0 * opix * d->Cout * d->Ktot;
  double in_bytes = 0;
  for (int s = 0; s < d->nsrc; ++s) in_bytes += 2.0 * d->B * d->Hin * d->Win * d->src_c[s];
  op.bytes = in_bytes + 2.0 * static_cast<double>(d->Cout_pad) * d->Ktot + opix * d->Cout * (d->out_bf16 ? 2.0 : 4.0) +
             (d->residual ? 2.0 * opix * d->Cout : 0.0) + (d->residual_f32 ? 4.0 * opix * d->Cout : 0.0);
  if (d->gn_nver > 0)  // fused GroupNorm: each normalised version is written once; the raw output only if requested
    op.bytes += 2.0 * opix * d->Cout * (d->gn_nver - (d->out_bf16 ? 0 : 1));
  return push(p, op);
}

int dmc_plan_add_attention(dmc_plan* p, const dmc_attn_desc* d) {
  DMC_REQUIRE(p && d, "dmc_plan_add_attention: null argument");
  Op op;
  op.kind = OP_ATTN;
  op.attn = *d;
  DMC_REQUIRE(d->impl == 0 || d->impl == 1, "dmc_plan_add_attention: impl=%d", d->impl);
  if (d->impl == 0 && d->qkv_lo == nullptr && d->out_lo == nullptr && attention_umma_supported(*d)) {
    int r = attention_prepare(*d, &op.attn_prep);
    if (r != 0) return r;
  }
  op.flops = 4.0 * d->B * static_cast<double>(d->L) * d->L * d->C;  // QK^T and PV
  op.bytes = 2.0 * d->B * d->L * 4.0 * d->C;
  return push(p, op);
}

int dmc_plan_add_dit_cond(dmc_plan* p, const dmc_dit_cond_desc* d) {
  DMC_REQUIRE(p && d, "dmc_plan_add_dit_cond: null argument");
  Op op;
  op.kind = OP_DIT_COND;
  op.dit_cond = *d;
  op.bytes = 4.0 * (static_cast<double>(d->ncols) * d->hidden + 2.0 * d->B * d->ncols);
  return push(p, op);
}

int dmc_plan_add_patch_embed(dmc_plan* p, const dmc_patch_embed_desc* d) {
  DMC_REQUIRE(p && d, "dmc_plan_add_patch_embed: null argument");
  DMC_REQUIRE(d->patch > 0, "dmc_plan_add_patch_embed: patch=%d", d->patch);
  Op op;
  op.kind = OP_PATCH_EMBED;
  op.patch = *d;
  const double tokens = static_cast<double>(d->B) * (d->H / d->patch) * (d->W / d->patch);
  op.bytes = 4.0 * (static_cast<double>(d->B) * d->Cin * d->H * d->W + tokens * d->hidden);
  op.flops = 2.0 * tokens * d->hidden * d->Cin * d->patch * d->patch;
  return push(p, op);
}

int dmc_plan_add_ln_modulate(dmc_plan* p, const dmc_ln_mod_desc* d) {
  DMC_REQUIRE(p && d, "dmc_plan_add_ln_modulate: null argument");
  Op op;
  op.kind = OP_LN_MOD;
  op.ln_mod = *d;
  op.bytes = 6.0 * d->B * static_cast<double>(d->L) * d->C;  // fp32 in, bf16 out
  return push(p, op);
}

int dmc_conv_gn_supported(int32_t B, int32_t Hout, int32_t Wout, int32_t Cout, int32_t max_gsize) {
  return conv_gn_supported(B, Hout, Wout, Cout, max_gsize) ? 1 : 0;
}

int dmc_head_supported(const dmc_head_desc* d) { return (d != nullptr && head_fused_supported(*d)) ? 1 : 0; }

int dmc_plan_add_head(dmc_plan* p, const dmc_head_desc* d) {
  DMC_REQUIRE(p && d, "dmc_plan_add_head: null argument");
  DMC_REQUIRE(head_fused_supported(*d), "dmc_plan_add_head: unsupported shape (C=%d W=%d Cout=%d)", d->C, d->W, d->Cout);
  Op op;
  op.kind = OP_HEAD;
  op.head = *d;
  const double pix = static_cast<double>(d->B) * d->H * d->W;
  op.bytes = pix * (2.0 * d->C + 4.0 * d->Cout);
  op.flops = 2.0 * pix * d->Cout * d->C * 9;
  return push(p, op);
}

int dmc_conv_affine_supported(int32_t B, int32_t H, int32_t W, int32_t Cin, int32_t Cout) {
  return conv_affine_supported(B, H, W, Cin, Cout) ? 1 : 0;
}

int dmc_plan_add_gn_coeff(dmc_plan* p, const dmc_gn_coeff_desc* d) {
  DMC_REQUIRE(p && d, "dmc_plan_add_gn_coeff: null argument");
  DMC_REQUIRE(d->stats && d->gamma && d->beta && d->out && d->B > 0 && d->HW > 0 && d->stats_slots > 0 && d->groups > 0 &&
                  d->groups <= 32 && d->C % 8 == 0 && (d->C / 8) % d->groups == 0 && d->C <= 1024,
              "dmc_plan_add_gn_coeff: bad arguments (C=%d groups=%d)", d->C, d->groups);
  Op op;
  op.kind = OP_GN_COEFF;
  op.gn_coeff = *d;
  op.bytes = static_cast<double>(d->B) * (8.0 * d->stats_slots * (d->C / 8) + 8.0 * d->C);
  return push(p, op);
}

int dmc_plan_add_stem_cols(dmc_plan* p, const dmc_stem_cols_desc* d) {
  DMC_REQUIRE(p && d, "dmc_plan_add_stem_cols: null argument");
  DMC_REQUIRE(d->x && d->out && d->B > 0 && d->x_batch > 0 && d->H > 0 && d->W > 0 && d->Cin > 0 && 18 * d->Cin <= 64,
              "dmc_plan_add_stem_cols: bad arguments (Cin=%d)", d->Cin);
  Op op;
  op.kind = OP_STEM_COLS;
  op.stem_cols = *d;
  op.bytes = static_cast<double>(d->B) * d->H * d->W * (4.0 * d->Cin + 128.0);
  return push(p, op);
}

int dmc_plan_add_head_taps(dmc_plan* p, const dmc_head_taps_desc* d) {
  DMC_REQUIRE(p && d, "dmc_plan_add_head_taps: null argument");
  DMC_REQUIRE(d->y && d->out && d->B > 0 && d->H > 0 && d->W > 0 && d->Cout > 0 && d->ypitch >= 9 * d->Cout && d->ypitch <= 64 &&
                  d->ypitch % 4 == 0, "dmc_plan_add_head_taps: bad arguments (Cout=%d ypitch=%d)", d->Cout, d->ypitch);
  Op op;
  op.kind = OP_HEAD_TAPS;
  op.head_taps = *d;
  const double pix = static_cast<double>(d->B) * d->H * d->W;
  op.bytes = pix * 4.0 * (d->ypitch + d->Cout);
  return push(p, op);
}

int dmc_plan_add_upsample(dmc_plan* p, const dmc_upsample_desc* d) {
  DMC_REQUIRE(p && d, "dmc_plan_add_upsample: null argument");
  Op op;
  op.kind = OP_UPSAMPLE;
  op.up = *d;
  op.bytes = 2.0 * 5.0 * d->B * d->H * d->W * d->C;
  return push(p, op);
}

static int add_step(dmc_plan* p, const dmc_step_desc* d, int kind) {
  DMC_REQUIRE(p && d, "dmc_plan_add_*_step: null argument");
  Op op;
  op.kind = kind;
  op.step = *d;
  const double n = static_cast<double>(d->B) * d->n_per_sample;
  op.bytes = 4.0 * n * (3.0 + (d->eps_u ? 1.0 : 0.0) + (d->noise ? 1.0 : 0.0));
  return push(p, op);
}
int dmc_plan_add_ddim_step(dmc_plan* p, const dmc_step_desc* d) { return add_step(p, d, OP_DDIM); }
int dmc_plan_add_ddpm_step(dmc_plan* p, const dmc_step_desc* d) { return add_step(p, d, OP_DDPM); }

}  // extern "C"
