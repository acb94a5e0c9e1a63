// Fused scheduler steps: CFG combine + x0 prediction + clamp / dynamic threshold + DDIM / DDPM update in ONE
// 128-bit-vectorised HBM pass.  Replaces /root/reference/diffusion/ddim.py:154-208,300-339 and
// diffusion/ddpm.py:151-220,289-324 (roughly 80-190 ATen launches and one host sync per step).
//
// Every arithmetic step uses explicit round-to-nearest intrinsics in the reference's evaluation order
// (each ATen elementwise op rounds once to fp32; nothing is contracted into an FMA), so results are
// bit-identical to the fp32 oracle for the same inputs and coefficients.
#include "common.cuh"
#include "kernels.h"

namespace dmc {

struct StepArgs {
  const float* x;
  const float* eps_c;
  const float* eps_u;
  const float* noise;
  float* out;
  int B;
  int n;  // elements per sample
  const float* coef;  // device: 5 floats (dmc_ddim_coef / dmc_ddpm_coef) per row
  const int* step_index;  // optional device scalar: row to use
  float cfg_scale;
  int clip_mode;
  int q_lo, q_hi;
  float q_w;
};

__device__ __forceinline__ float guided_eps(float ec, float eu, bool has_u, float s) {
  // eps_u + s * (eps_c - eps_u)
  return has_u ? __fadd_rn(eu, __fmul_rn(s, __fsub_rn(ec, eu))) : ec;
}

template <bool DDPM>
__device__ __forceinline__ float predict_x0(float x, float eps, const float (&c)[5]) {
  if (DDPM) {
    // sqrt(1/acp)[t] * x - sqrt(1/acp - 1)[t] * eps
    return __fsub_rn(__fmul_rn(c[0], x), __fmul_rn(c[1], eps));
  } else {
    // (x - sqrt(1 - a) * eps) / sqrt(a)
    return __fdiv_rn(__fsub_rn(x, __fmul_rn(c[0], eps)), c[1]);
  }
}

template <bool DDPM>
__device__ __forceinline__ float finish(float x, float eps, float x0, float nz, bool has_noise, const float (&c)[5]) {
  if (DDPM) {
    // c1 * x0 + c2 * x  +  noise_scale * noise
    float mean = __fadd_rn(__fmul_rn(c[2], x0), __fmul_rn(c[3], x));
    return __fadd_rn(mean, __fmul_rn(c[4], nz));
  } else {
    // sqrt(a') * x0 + dir_coef * eps  (+ sigma * noise)
    float v = __fadd_rn(__fmul_rn(c[2], x0), __fmul_rn(c[3], eps));
    if (has_noise) v = __fadd_rn(v, __fmul_rn(c[4], nz));
    return v;
  }
}

// ---- elementwise path (clip_mode 0 / 1): grid-stride over float4 -------------------------------------
template <bool DDPM>
__global__ void __launch_bounds__(256) step_elementwise_kernel(StepArgs a) {
  float c[5];
  const float* crow = a.coef + (a.step_index ? 5 * static_cast<size_t>(__ldg(a.step_index)) : 0);
#pragma unroll
  for (int i = 0; i < 5; ++i) c[i] = __ldg(crow + i);
  const bool has_u = a.eps_u != nullptr;
  const bool has_noise = DDPM ? true : (a.noise != nullptr && c[4] != 0.0f);
  const size_t total4 = (static_cast<size_t>(a.B) * a.n) >> 2;
  const float4* x4 = reinterpret_cast<const float4*>(a.x);
  const float4* ec4 = reinterpret_cast<const float4*>(a.eps_c);
  const float4* eu4 = reinterpret_cast<const float4*>(a.eps_u);
  const float4* nz4 = reinterpret_cast<const float4*>(a.noise);
  float4* o4 = reinterpret_cast<float4*>(a.out);
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total4;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    float4 xv = x4[i], ec = ec4[i];
    float4 eu = has_u ? eu4[i] : make_float4(0, 0, 0, 0);
    float4 nz = has_noise ? nz4[i] : make_float4(0, 0, 0, 0);
    float xs[4] = {xv.x, xv.y, xv.z, xv.w}, es[4] = {ec.x, ec.y, ec.z, ec.w}, us[4] = {eu.x, eu.y, eu.z, eu.w},
          ns[4] = {nz.x, nz.y, nz.z, nz.w}, r[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float eps = guided_eps(es[j], us[j], has_u, a.cfg_scale);
      float x0 = predict_x0<DDPM>(xs[j], eps, c);
      if (a.clip_mode == 1) x0 = fminf(fmaxf(x0, -1.0f), 1.0f);
      r[j] = finish<DDPM>(xs[j], eps, x0, ns[j], has_noise, c);
    }
    o4[i] = make_float4(r[0], r[1], r[2], r[3]);
  }
}

// scalar variant for shapes whose element count is not a multiple of 4
template <bool DDPM>
__global__ void __launch_bounds__(256) step_scalar_kernel(StepArgs a) {
  float c[5];
  const float* crow = a.coef + (a.step_index ? 5 * static_cast<size_t>(__ldg(a.step_index)) : 0);
#pragma unroll
  for (int i = 0; i < 5; ++i) c[i] = __ldg(crow + i);
  const bool has_u = a.eps_u != nullptr;
  const bool has_noise = DDPM ? true : (a.noise != nullptr && c[4] != 0.0f);
  const size_t total = static_cast<size_t>(a.B) * a.n;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    float x = a.x[i];
    float eps = guided_eps(a.eps_c[i], has_u ? a.eps_u[i] : 0.f, has_u, a.cfg_scale);
    float x0 = predict_x0<DDPM>(x, eps, c);
    if (a.clip_mode == 1) x0 = fminf(fmaxf(x0, -1.0f), 1.0f);
    a.out[i] = finish<DDPM>(x, eps, x0, has_noise ? a.noise[i] : 0.f, has_noise, c);
  }
}

// ---- dynamic-threshold path (clip_mode 2): one CTA per sample ----------------------------------------
// s = max(quantile(|x0|, p), 1) per sample; the quantile is torch.quantile's: lerp between the
// elements of rank q_lo and q_hi of the ascending sort.  Ranks are found with an MSB-first 8-bit radix
// select on the IEEE bit patterns of |x0| (monotone for non-negative floats) held in shared memory.
constexpr int THR_THREADS = 256;

template <bool DDPM>
__global__ void __launch_bounds__(THR_THREADS) step_threshold_kernel(StepArgs a) {
  extern __shared__ float sm_x0[];  // n floats
  __shared__ unsigned int hist[256];
  __shared__ unsigned int sel_prefix, sel_rank;
  __shared__ unsigned int red_cnt[THR_THREADS / 32];
  __shared__ unsigned int red_min[THR_THREADS / 32];

  float c[5];
  const float* crow = a.coef + (a.step_index ? 5 * static_cast<size_t>(__ldg(a.step_index)) : 0);
#pragma unroll
  for (int i = 0; i < 5; ++i) c[i] = __ldg(crow + i);
  const bool has_u = a.eps_u != nullptr;
  const bool has_noise = DDPM ? true : (a.noise != nullptr && c[4] != 0.0f);
  const int n = a.n;
  const size_t base = static_cast<size_t>(blockIdx.x) * n;
  const int tid = threadIdx.x;

  for (int i = tid; i < n; i += THR_THREADS) {
    float eps = guided_eps(a.eps_c[base + i], has_u ? a.eps_u[base + i] : 0.f, has_u, a.cfg_scale);
    sm_x0[i] = predict_x0<DDPM>(a.x[base + i], eps, c);
  }
  if (tid == 0) {
    sel_prefix = 0u;
    sel_rank = static_cast<unsigned int>(a.q_lo);
  }
  __syncthreads();

  // radix select of the element of ascending rank q_lo
  for (int shift = 24; shift >= 0; shift -= 8) {
    hist[tid] = 0u;  // THR_THREADS == 256 bins
    __syncthreads();
    const unsigned int prefix = sel_prefix;
    const unsigned int mask_hi = (shift == 24) ? 0u : (0xFFFFFFFFu << (shift + 8));
    for (int i = tid; i < n; i += THR_THREADS) {
      unsigned int key = __float_as_uint(fabsf(sm_x0[i]));
      if ((key & mask_hi) == prefix) atomicAdd(&hist[(key >> shift) & 0xFFu], 1u);
    }
    __syncthreads();
    if (tid == 0) {
      unsigned int r = sel_rank, acc = 0u;
      int b = 0;
      for (; b < 256; ++b) {
        unsigned int h = hist[b];
        if (acc + h > r) break;
        acc += h;
      }
      sel_rank = r - acc;
      sel_prefix = prefix | (static_cast<unsigned int>(b) << shift);
    }
    __syncthreads();
  }
  const unsigned int key_lo = sel_prefix;

  // rank q_hi: equals key_lo if enough elements are <= key_lo, else the smallest key above it
  unsigned int cnt_le = 0u, min_gt = 0xFFFFFFFFu;
  for (int i = tid; i < n; i += THR_THREADS) {
    unsigned int key = __float_as_uint(fabsf(sm_x0[i]));
    if (key <= key_lo) ++cnt_le;
    else min_gt = min(min_gt, key);
  }
  cnt_le = __reduce_add_sync(0xFFFFFFFFu, cnt_le);
  min_gt = __reduce_min_sync(0xFFFFFFFFu, min_gt);
  if ((tid & 31) == 0) {
    red_cnt[tid >> 5] = cnt_le;
    red_min[tid >> 5] = min_gt;
  }
  __syncthreads();
  cnt_le = 0u;
  min_gt = 0xFFFFFFFFu;
#pragma unroll
  for (int w = 0; w < THR_THREADS / 32; ++w) {
    cnt_le += red_cnt[w];
    min_gt = min(min_gt, red_min[w]);
  }
  const float v_lo = __uint_as_float(key_lo);
  const float v_hi = (a.q_hi == a.q_lo || cnt_le > static_cast<unsigned int>(a.q_hi)) ? v_lo : __uint_as_float(min_gt);
  // ATen lerp: w < 0.5 ? a + w * (b - a) : b - (b - a) * (1 - w)
  const float diff = __fsub_rn(v_hi, v_lo);
  float s = (a.q_w < 0.5f) ? __fadd_rn(v_lo, __fmul_rn(a.q_w, diff))
                           : __fsub_rn(v_hi, __fmul_rn(diff, __fsub_rn(1.0f, a.q_w)));
  s = fmaxf(s, 1.0f);

  for (int i = tid; i < n; i += THR_THREADS) {
    float x = a.x[base + i];
    float eps = guided_eps(a.eps_c[base + i], has_u ? a.eps_u[base + i] : 0.f, has_u, a.cfg_scale);
    float x0 = __fdiv_rn(fminf(fmaxf(sm_x0[i], -s), s), s);
    a.out[base + i] = finish<DDPM>(x, eps, x0, has_noise ? a.noise[base + i] : 0.f, has_noise, c);
  }
}

template <bool DDPM>
static int launch_step_t(const StepArgs& a, cudaStream_t st) {
  DMC_REQUIRE(a.x && a.eps_c && a.out && a.coef, "sched step: null pointer argument");
  DMC_REQUIRE(a.B > 0 && a.n > 0, "sched step: empty batch (B=%d, n=%d)", a.B, a.n);
  if (DDPM) DMC_REQUIRE(a.noise != nullptr, "ddpm step: noise is required");
  if (a.clip_mode == 2) {
    DMC_REQUIRE(a.q_lo >= 0 && a.q_hi >= a.q_lo && a.q_hi < a.n && a.q_hi - a.q_lo <= 1,
                "sched step: bad quantile ranks (%d, %d) for n=%d", a.q_lo, a.q_hi, a.n);
    size_t smem = static_cast<size_t>(a.n) * sizeof(float);
    DMC_REQUIRE(smem <= 200 * 1024, "sched step: n_per_sample=%d too large for the threshold kernel", a.n);
    if (smem > 40 * 1024)
      DMC_CUDA_OK(cudaFuncSetAttribute(step_threshold_kernel<DDPM>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       static_cast<int>(smem)));
    step_threshold_kernel<DDPM><<<a.B, THR_THREADS, smem, st>>>(a);
  } else {
    size_t total = static_cast<size_t>(a.B) * a.n;
    const bool vec = (total % 4 == 0) && ((reinterpret_cast<uintptr_t>(a.x) | reinterpret_cast<uintptr_t>(a.eps_c) |
                                           reinterpret_cast<uintptr_t>(a.eps_u) | reinterpret_cast<uintptr_t>(a.noise) |
                                           reinterpret_cast<uintptr_t>(a.out)) % 16 == 0);
    size_t work = vec ? total / 4 : total;
    int blocks = static_cast<int>(std::min<size_t>((work + 255) / 256, static_cast<size_t>(num_sms()) * 8));
    if (blocks < 1) blocks = 1;
    if (vec) step_elementwise_kernel<DDPM><<<blocks, 256, 0, st>>>(a);
    else step_scalar_kernel<DDPM><<<blocks, 256, 0, st>>>(a);
  }
  DMC_CUDA_OK(cudaGetLastError());
  return 0;
}

int launch_step(bool ddpm, const dmc_step_desc& d, cudaStream_t st) {
  StepArgs a;
  a.x = d.x; a.eps_c = d.eps_c; a.eps_u = d.eps_u; a.noise = d.noise; a.out = d.x_out;
  a.B = d.B; a.n = d.n_per_sample; a.coef = reinterpret_cast<const float*>(d.coef_dev);
  a.step_index = d.step_index_dev;
  a.cfg_scale = d.g.cfg_scale; a.clip_mode = d.g.clip_mode; a.q_lo = d.g.q_lo; a.q_hi = d.g.q_hi; a.q_w = d.g.q_weight;
  DMC_REQUIRE(a.clip_mode >= 0 && a.clip_mode <= 2, "sched step: clip_mode %d", a.clip_mode);
  return ddpm ? launch_step_t<true>(a, st) : launch_step_t<false>(a, st);
}

__global__ void advance_kernel(int* counter, const int64_t* __restrict__ t_table, int64_t* __restrict__ t_out, int n) {
  const int cur = counter[0];
  const int64_t t = t_table[cur];
  for (int i = threadIdx.x; i < n; i += blockDim.x) t_out[i] = t;
  __syncthreads();
  if (threadIdx.x == 0) {
    counter[1] = cur;
    counter[0] = cur + 1;
  }
}

int launch_advance(int* counter, const int64_t* t_table, int64_t* t_out, int n, cudaStream_t st) {
  DMC_REQUIRE(counter && t_table && t_out && n > 0, "advance: bad arguments");
  advance_kernel<<<1, 256, 0, st>>>(counter, t_table, t_out, n);
  DMC_CUDA_OK(cudaGetLastError());
  return 0;
}

__global__ void __launch_bounds__(256) q_sample_kernel(const float* __restrict__ x0, const float* __restrict__ noise,
                                                       const int64_t* __restrict__ t, const float* __restrict__ sa,
                                                       const float* __restrict__ s1, float* __restrict__ out, int B,
                                                       int n) {
  const size_t total = static_cast<size_t>(B) * n;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    int b = static_cast<int>(i / n);
    int64_t tt = t[b];
    out[i] = __fadd_rn(__fmul_rn(sa[tt], x0[i]), __fmul_rn(s1[tt], noise[i]));
  }
}

int launch_q_sample(const float* x0, const float* noise, const int64_t* t, const float* sa, const float* s1, float* out,
                    int B, int n, cudaStream_t st) {
  DMC_REQUIRE(x0 && noise && t && sa && s1 && out && B > 0 && n > 0, "q_sample: bad arguments");
  size_t total = static_cast<size_t>(B) * n;
  int blocks = static_cast<int>(std::min<size_t>((total + 255) / 256, static_cast<size_t>(num_sms()) * 8));
  q_sample_kernel<<<blocks, 256, 0, st>>>(x0, noise, t, sa, s1, out, B, n);
  DMC_CUDA_OK(cudaGetLastError());
  return 0;
}

}  // namespace dmc
