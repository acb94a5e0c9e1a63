// Multi-tensor optimizer step of the training loop (reference: utils/trainer.py:256-262 --
// clip_grad_norm_(1.0), AdamW.step(), EMA update): two launches for ALL parameters instead of torch's per-list foreach kernels.
//   1. global gradient norm: one partial sum of squares per 16 K-element chunk, added in index order (deterministic)
//   2. AdamW (decoupled weight decay, bias-corrected, as torch.optim.AdamW) on the clipped gradient + optional EMA of the new
//      parameter value, one pass over p, g, m, v (, ema)
// Work is cut into fixed chunks listed in a device table, so a 3-element bias and a 1.2 M-element convolution weight load the
// SMs alike.  HBM-bound: 7 (9 with EMA) fp32 streams of the parameter count.
#include "common.cuh"
#include "kernels.h"

namespace dmc {

constexpr int OPT_CHUNK = 16384;

__device__ __forceinline__ float block_sum_256(float s, float* red) {
#pragma unroll
  for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xFFFFFFFFu, s, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  float t = 0.f;
  if (threadIdx.x == 0) {
#pragma unroll
    for (int w = 0; w < 8; ++w) t += red[w];
  }
  return t;  // valid in thread 0
}

__global__ void __launch_bounds__(256) opt_sqnorm_kernel(const dmc_opt_item* __restrict__ items,
                                                         const dmc_opt_chunk* __restrict__ chunks, float* __restrict__ partial) {
  __shared__ float red[8];
  const dmc_opt_chunk c = chunks[blockIdx.x];
  const dmc_opt_item it = items[c.item];
  const int64_t end = it.g != nullptr ? min(c.start + static_cast<int64_t>(OPT_CHUNK), it.n) : c.start;  // no gradient: skipped
  float s = 0.f;
  for (int64_t i = c.start + threadIdx.x; i < end; i += 256) {
    const float x = __ldg(it.g + i);
    s = fmaf(x, x, s);
  }
  const float t = block_sum_256(s, red);
  if (threadIdx.x == 0) partial[blockIdx.x] = t;
}

// norm[0] = sqrt(sum of the chunk partials): 1024 lanes add strided partials in index order, then a fixed tree
__global__ void __launch_bounds__(1024) opt_norm_finish_kernel(const float* __restrict__ partial, int n, float* __restrict__ norm) {
  __shared__ float red[1024];
  float s = 0.f;
  for (int i = threadIdx.x; i < n; i += 1024) s += partial[i];
  red[threadIdx.x] = s;
  __syncthreads();
  for (int o = 512; o; o >>= 1) {
    if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) norm[0] = sqrtf(red[0]);
}

__global__ void __launch_bounds__(256) opt_adamw_kernel(const dmc_opt_item* __restrict__ items,
                                                        const dmc_opt_chunk* __restrict__ chunks, dmc_adamw_desc h,
                                                        const float* __restrict__ norm) {
  const dmc_opt_chunk c = chunks[blockIdx.x];
  const dmc_opt_item it = items[c.item];
  if (it.g == nullptr) return;  // a parameter without a gradient is left alone, as torch.optim.AdamW does
  const int64_t end = min(c.start + static_cast<int64_t>(OPT_CHUNK), it.n);
  // torch.nn.utils.clip_grad_norm_: coef = clamp(max_norm / (total_norm + 1e-6), max = 1)
  const float coef = (h.max_norm > 0.f && norm != nullptr) ? fminf(1.0f, h.max_norm / (__ldg(norm) + 1e-6f)) : 1.0f;
  const float decay = 1.0f - h.lr * h.weight_decay;
  const float step_size = h.lr / h.bias_correction1;
  const float inv_sqrt_bc2 = rsqrtf(h.bias_correction2);
  for (int64_t i = c.start + threadIdx.x; i < end; i += 256) {
    const float g = __ldg(it.g + i) * coef;
    float p = it.p[i] * decay;
    const float m = fmaf(h.beta1, it.m[i], (1.0f - h.beta1) * g);
    const float v = fmaf(h.beta2, it.v[i], (1.0f - h.beta2) * g * g);
    p -= step_size * m / (sqrtf(v) * inv_sqrt_bc2 + h.eps);
    it.p[i] = p;
    it.m[i] = m;
    it.v[i] = v;
    if (it.ema != nullptr && h.ema_decay > 0.f) it.ema[i] = fmaf(h.ema_decay, it.ema[i], (1.0f - h.ema_decay) * p);
  }
}

int launch_opt_grad_norm(const dmc_opt_item* items, const dmc_opt_chunk* chunks, int n_chunks, float* partial, float* norm,
                         cudaStream_t st) {
  DMC_REQUIRE(items && chunks && partial && norm && n_chunks > 0, "opt_grad_norm: bad arguments");
  opt_sqnorm_kernel<<<n_chunks, 256, 0, st>>>(items, chunks, partial);
  DMC_CUDA_OK(cudaGetLastError());
  opt_norm_finish_kernel<<<1, 1024, 0, st>>>(partial, n_chunks, norm);
  DMC_CUDA_OK(cudaGetLastError());
  return 0;
}

int launch_opt_adamw(const dmc_opt_item* items, const dmc_opt_chunk* chunks, int n_chunks, const dmc_adamw_desc& h,
                     const float* norm, cudaStream_t st) {
  DMC_REQUIRE(items && chunks && n_chunks > 0, "opt_adamw: bad arguments");
  DMC_REQUIRE(h.bias_correction1 > 0.f && h.bias_correction2 > 0.f && h.lr >= 0.f, "opt_adamw: bad hyper-parameters");
  opt_adamw_kernel<<<n_chunks, 256, 0, st>>>(items, chunks, h, norm);
  DMC_CUDA_OK(cudaGetLastError());
  return 0;
}

}  // namespace dmc
