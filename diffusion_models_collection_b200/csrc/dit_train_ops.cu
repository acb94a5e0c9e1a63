// Memory-bound glue of the DiT training step (models/dit_train.py; reference models/dit.py:111-132 under autograd), fused so that
// the fp32 token stream is read and written once per residual / LayerNorm boundary:
//
//   dit_gate_ln_mod            x_out = x_in + gate[n] * y                      (gated residual, models/dit.py:121,127)
//                              h     = LayerNorm(x_out) * (1 + scale[n]) + shift[n]  -> bf16, the next GEMM's operand (:117-118,:124-125)
//   dit_gate_ln_mod_backward   from dh (the GEMM's input gradient) and dx_out (the stream's gradient from later layers):
//                              dx_in = dx_out + LayerNorm'(dh * (1 + scale)),  dy = dx_in * gate,
//                              dgate[n] = sum_l dx_in * y,  dshift[n] = sum_l dh,  dscale[n] = sum_l dh * xhat
//   gelu_forward / backward    nn.GELU() (erf form) between fc1 and fc2 (:95-99)
//
// One warp per token row (C = 128 * V4 channels, float4 per lane per 128 channels); the backward kernel runs one CTA per image so
// that the three per-image column sums are reduced in a fixed order (warp 0 .. 7 through shared memory): deterministic, no atomics.
#include "common.cuh"
#include "kernels.h"

namespace dmc {

// nn.Dropout on four consecutive elements (first index idx4, a multiple of 4): the counter-based mask of the UNet's GroupNorm pass
// (one 32-bit hash per element pair, 16 bits each, kept when >= thresh16), regenerated identically by the backward kernels
__device__ __forceinline__ void drop4(float2& a, float2& b, uint32_t seed, uint64_t idx4, uint32_t thresh, float scale) {
  const uint32_t h0 = dropout_hash2(seed, idx4), h1 = dropout_hash2(seed, idx4 + 2);
  a.x = (h0 & 0xFFFFu) >= thresh ? a.x * scale : 0.f;
  a.y = (h0 >> 16) >= thresh ? a.y * scale : 0.f;
  b.x = (h1 & 0xFFFFu) >= thresh ? b.x * scale : 0.f;
  b.y = (h1 >> 16) >= thresh ? b.y * scale : 0.f;
}

template <int V4>
__global__ void __launch_bounds__(256) dit_gate_ln_mod_kernel(const float* __restrict__ x_in, const __nv_bfloat16* __restrict__ y,
                                                              const float* __restrict__ gate, float* __restrict__ x_out,
                                                              __nv_bfloat16* __restrict__ h, const float* __restrict__ shift,
                                                              const float* __restrict__ scale, int mod_stride, int gate_stride,
                                                              size_t tokens, int L, float eps, uint32_t drop_thresh, float drop_scale,
                                                              uint32_t seed) {
  constexpr int C = 128 * V4;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const size_t tok = static_cast<size_t>(blockIdx.x) * 8 + warp;
  if (tok >= tokens) return;
  const int n = static_cast<int>(tok / L);
  const float4* xr = reinterpret_cast<const float4*>(x_in + tok * C);
  float4 v[V4];
#pragma unroll
  for (int i = 0; i < V4; ++i) v[i] = xr[lane + 32 * i];
  if (y != nullptr) {
    const uint2* yr = reinterpret_cast<const uint2*>(y + tok * C);
    const float4* g4 = reinterpret_cast<const float4*>(gate + static_cast<size_t>(n) * gate_stride);
    float4* xo = reinterpret_cast<float4*>(x_out + tok * C);
#pragma unroll
    for (int i = 0; i < V4; ++i) {
      const uint2 yy = yr[lane + 32 * i];
      float2 a = unpack_bf16x2(yy.x), b = unpack_bf16x2(yy.y);
      if (drop_thresh != 0u) drop4(a, b, seed, tok * C + 4 * (lane + 32 * i), drop_thresh, drop_scale);
      const float4 g = __ldg(g4 + lane + 32 * i);
      v[i].x = fmaf(g.x, a.x, v[i].x);
      v[i].y = fmaf(g.y, a.y, v[i].y);
      v[i].z = fmaf(g.z, b.x, v[i].z);
      v[i].w = fmaf(g.w, b.y, v[i].w);
      xo[lane + 32 * i] = v[i];
    }
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < V4; ++i) s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
#pragma unroll
  for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xFFFFFFFFu, s, o);
  const float mean = s * (1.0f / C);
  float ss = 0.f;
#pragma unroll
  for (int i = 0; i < V4; ++i) {
    const float a = v[i].x - mean, b = v[i].y - mean, c = v[i].z - mean, e = v[i].w - mean;
    ss += (a * a + b * b) + (c * c + e * e);
  }
#pragma unroll
  for (int o = 16; o; o >>= 1) ss += __shfl_xor_sync(0xFFFFFFFFu, ss, o);
  const float rstd = rsqrtf(ss * (1.0f / C) + eps);
  const float4* sh4 = reinterpret_cast<const float4*>(shift + static_cast<size_t>(n) * mod_stride);
  const float4* sc4 = reinterpret_cast<const float4*>(scale + static_cast<size_t>(n) * mod_stride);
  uint2* o2 = reinterpret_cast<uint2*>(h + tok * C);
#pragma unroll
  for (int i = 0; i < V4; ++i) {
    const float4 sh = __ldg(sh4 + lane + 32 * i), sc = __ldg(sc4 + lane + 32 * i);
    const float y0 = fmaf((v[i].x - mean) * rstd, 1.0f + sc.x, sh.x);
    const float y1 = fmaf((v[i].y - mean) * rstd, 1.0f + sc.y, sh.y);
    const float y2 = fmaf((v[i].z - mean) * rstd, 1.0f + sc.z, sh.z);
    const float y3 = fmaf((v[i].w - mean) * rstd, 1.0f + sc.w, sh.w);
    o2[lane + 32 * i] = make_uint2(pack_bf16x2(y0, y1), pack_bf16x2(y2, y3));
  }
}

// one CTA per image: warp w owns rows w, w + 8, ...; lane owns channels 4 (lane + 32 i) .. + 3
template <int V4>
__global__ void __launch_bounds__(256) dit_gate_ln_mod_bwd_kernel(const float* __restrict__ x, const __nv_bfloat16* __restrict__ dh,
                                                                  const float* __restrict__ dx_out, const __nv_bfloat16* __restrict__ y,
                                                                  const float* __restrict__ gate, const float* __restrict__ scale,
                                                                  int mod_stride, int gate_stride, float* __restrict__ dx_in,
                                                                  __nv_bfloat16* __restrict__ dy, float* __restrict__ dgate,
                                                                  float* __restrict__ dshift, float* __restrict__ dscale, int L,
                                                                  float eps, uint32_t drop_thresh, float drop_scale, uint32_t seed,
                                                                  float* __restrict__ part) {
  constexpr int C = 128 * V4;
  __shared__ float red[3 * C];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n = blockIdx.x;
  // gridDim.y CTAs share one image (row slices): each writes its partial column sums to part[n][slice][3][C], a second
  // launch adds the slices in index order (deterministic); gridDim.y == 1 writes the outputs directly
  const int S = gridDim.y, rows = (L + S - 1) / S;
  const int l_begin = blockIdx.y * rows, l_end = min(L, l_begin + rows);
  const float4* sc4 = reinterpret_cast<const float4*>(scale + static_cast<size_t>(n) * mod_stride);
  const float4* g4 = gate != nullptr ? reinterpret_cast<const float4*>(gate + static_cast<size_t>(n) * gate_stride) : nullptr;
  float4 sc1[V4], gt[V4];
#pragma unroll
  for (int i = 0; i < V4; ++i) {
    const float4 s_ = __ldg(sc4 + lane + 32 * i);
    sc1[i] = make_float4(1.0f + s_.x, 1.0f + s_.y, 1.0f + s_.z, 1.0f + s_.w);
    gt[i] = g4 != nullptr ? __ldg(g4 + lane + 32 * i) : make_float4(0.f, 0.f, 0.f, 0.f);
  }
  float4 a_gate[V4], a_shift[V4], a_scale[V4];
#pragma unroll
  for (int i = 0; i < V4; ++i) a_gate[i] = a_shift[i] = a_scale[i] = make_float4(0.f, 0.f, 0.f, 0.f);

  for (int l = l_begin + warp; l < l_end; l += 8) {
    const size_t tok = static_cast<size_t>(n) * L + l;
    const float4* xr = reinterpret_cast<const float4*>(x + tok * C);
    const uint2* dhr = reinterpret_cast<const uint2*>(dh + tok * C);
    float4 v[V4], g[V4];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < V4; ++i) {
      v[i] = xr[lane + 32 * i];
      s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xFFFFFFFFu, s, o);
    const float mean = s * (1.0f / C);
    float ss = 0.f;
#pragma unroll
    for (int i = 0; i < V4; ++i) {
      v[i].x -= mean; v[i].y -= mean; v[i].z -= mean; v[i].w -= mean;
      ss += (v[i].x * v[i].x + v[i].y * v[i].y) + (v[i].z * v[i].z + v[i].w * v[i].w);
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) ss += __shfl_xor_sync(0xFFFFFFFFu, ss, o);
    const float rstd = rsqrtf(ss * (1.0f / C) + eps);
    float m1 = 0.f, m2 = 0.f;
#pragma unroll
    for (int i = 0; i < V4; ++i) {
      const uint2 d_ = dhr[lane + 32 * i];
      const float2 a = unpack_bf16x2(d_.x), b = unpack_bf16x2(d_.y);
      v[i].x *= rstd; v[i].y *= rstd; v[i].z *= rstd; v[i].w *= rstd;      // xhat
      a_shift[i].x += a.x; a_shift[i].y += a.y; a_shift[i].z += b.x; a_shift[i].w += b.y;
      a_scale[i].x = fmaf(a.x, v[i].x, a_scale[i].x); a_scale[i].y = fmaf(a.y, v[i].y, a_scale[i].y);
      a_scale[i].z = fmaf(b.x, v[i].z, a_scale[i].z); a_scale[i].w = fmaf(b.y, v[i].w, a_scale[i].w);
      g[i] = make_float4(a.x * sc1[i].x, a.y * sc1[i].y, b.x * sc1[i].z, b.y * sc1[i].w);
      m1 += (g[i].x + g[i].y) + (g[i].z + g[i].w);
      m2 += (g[i].x * v[i].x + g[i].y * v[i].y) + (g[i].z * v[i].z + g[i].w * v[i].w);
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) {
      m1 += __shfl_xor_sync(0xFFFFFFFFu, m1, o);
      m2 += __shfl_xor_sync(0xFFFFFFFFu, m2, o);
    }
    m1 *= (1.0f / C);
    m2 *= (1.0f / C);
    const float4* dxo = dx_out != nullptr ? reinterpret_cast<const float4*>(dx_out + tok * C) : nullptr;
    float4* dxi = reinterpret_cast<float4*>(dx_in + tok * C);
#pragma unroll
    for (int i = 0; i < V4; ++i) {
      float4 d_;
      d_.x = rstd * (g[i].x - m1 - v[i].x * m2);
      d_.y = rstd * (g[i].y - m1 - v[i].y * m2);
      d_.z = rstd * (g[i].z - m1 - v[i].z * m2);
      d_.w = rstd * (g[i].w - m1 - v[i].w * m2);
      if (dxo != nullptr) {
        const float4 o_ = dxo[lane + 32 * i];  // (plain load: dx_in may be the same buffer)
        d_.x += o_.x; d_.y += o_.y; d_.z += o_.z; d_.w += o_.w;
      }
      dxi[lane + 32 * i] = d_;
      if (y != nullptr) {
        const uint2 yy = reinterpret_cast<const uint2*>(y + tok * C)[lane + 32 * i];
        float2 a = unpack_bf16x2(yy.x), b = unpack_bf16x2(yy.y);
        float2 ga = make_float2(d_.x * gt[i].x, d_.y * gt[i].y), gb = make_float2(d_.z * gt[i].z, d_.w * gt[i].w);
        if (drop_thresh != 0u) {  // the forward pass used dropout(y): same mask on y (for dgate) and on dy
          const uint64_t idx4 = tok * C + 4 * (lane + 32 * i);
          drop4(a, b, seed, idx4, drop_thresh, drop_scale);
          drop4(ga, gb, seed, idx4, drop_thresh, drop_scale);
        }
        a_gate[i].x = fmaf(d_.x, a.x, a_gate[i].x); a_gate[i].y = fmaf(d_.y, a.y, a_gate[i].y);
        a_gate[i].z = fmaf(d_.z, b.x, a_gate[i].z); a_gate[i].w = fmaf(d_.w, b.y, a_gate[i].w);
        reinterpret_cast<uint2*>(dy + tok * C)[lane + 32 * i] = make_uint2(pack_bf16x2(ga.x, ga.y), pack_bf16x2(gb.x, gb.y));
      }
    }
  }
  // per-image column sums: warps add their partials in index order
  for (int w = 0; w < 8; ++w) {
    if (warp == w) {
#pragma unroll
      for (int i = 0; i < V4; ++i) {
        float4* r0 = reinterpret_cast<float4*>(red) + lane + 32 * i;
        float4* r1 = reinterpret_cast<float4*>(red + C) + lane + 32 * i;
        float4* r2 = reinterpret_cast<float4*>(red + 2 * C) + lane + 32 * i;
        if (w == 0) {
          *r0 = a_gate[i]; *r1 = a_shift[i]; *r2 = a_scale[i];
        } else {
          float4 t = *r0; t.x += a_gate[i].x; t.y += a_gate[i].y; t.z += a_gate[i].z; t.w += a_gate[i].w; *r0 = t;
          t = *r1; t.x += a_shift[i].x; t.y += a_shift[i].y; t.z += a_shift[i].z; t.w += a_shift[i].w; *r1 = t;
          t = *r2; t.x += a_scale[i].x; t.y += a_scale[i].y; t.z += a_scale[i].z; t.w += a_scale[i].w; *r2 = t;
        }
      }
    }
    __syncthreads();
  }
  if (S > 1) {
    float* dst = part + (static_cast<size_t>(n) * S + blockIdx.y) * 3 * C;
    for (int c = threadIdx.x; c < 3 * C; c += 256) dst[c] = red[c];
    return;
  }
  for (int c = threadIdx.x; c < C; c += 256) {
    if (dgate != nullptr) dgate[static_cast<size_t>(n) * C + c] = red[c];
    dshift[static_cast<size_t>(n) * C + c] = red[C + c];
    dscale[static_cast<size_t>(n) * C + c] = red[2 * C + c];
  }
}

// adds the row-slice partials of dit_gate_ln_mod_bwd_kernel in slice order: part[n][s][3][C] -> dgate / dshift / dscale [n][C]
__global__ void __launch_bounds__(256) dit_glm_bwd_reduce_kernel(const float* __restrict__ part, int S, int C, float* __restrict__ dgate,
                                                                 float* __restrict__ dshift, float* __restrict__ dscale, int total) {
  const int i = blockIdx.x * 256 + threadIdx.x;  // over B * C
  if (i >= total) return;
  const int n = i / C, c = i - n * C;
  const float* p = part + static_cast<size_t>(n) * S * 3 * C + c;
  float a = 0.f, b = 0.f, d = 0.f;
  for (int s = 0; s < S; ++s) {
    a += p[(s * 3 + 0) * C];
    b += p[(s * 3 + 1) * C];
    d += p[(s * 3 + 2) * C];
  }
  if (dgate != nullptr) dgate[i] = a;
  dshift[i] = b;
  dscale[i] = d;
}

int launch_dit_gate_ln_mod(const dmc_dit_glm_desc& d, cudaStream_t st) {
  DMC_REQUIRE(d.x_in && d.h && d.shift && d.scale, "dit_gate_ln_mod: null pointer argument");
  DMC_REQUIRE((d.y == nullptr) == (d.gate == nullptr) && (d.y == nullptr) == (d.x_out == nullptr),
              "dit_gate_ln_mod: y, gate and x_out go together");
  DMC_REQUIRE(d.B > 0 && d.L > 0 && d.C % 128 == 0 && d.C >= 128 && d.C <= 1024 && d.mod_stride % 4 == 0 && d.gate_stride % 4 == 0,
              "dit_gate_ln_mod: C=%d must be a multiple of 128 in [128, 1024]", d.C);
  const size_t tokens = static_cast<size_t>(d.B) * d.L;
  const int blocks = static_cast<int>((tokens + 7) / 8);
  const __nv_bfloat16* y = reinterpret_cast<const __nv_bfloat16*>(d.y);
  __nv_bfloat16* h = reinterpret_cast<__nv_bfloat16*>(d.h);
  DMC_REQUIRE(d.drop_p >= 0.f && d.drop_p < 1.f, "dit_gate_ln_mod: drop_p=%f", d.drop_p);
  const uint32_t dth = dropout_threshold(d.drop_p);
  const float dsc = d.drop_p > 0.f ? 1.0f / (1.0f - d.drop_p) : 1.0f;
#define GLM(V) dit_gate_ln_mod_kernel<V><<<blocks, 256, 0, st>>>(d.x_in, y, d.gate, d.x_out, h, d.shift, d.scale, d.mod_stride, d.gate_stride, tokens, d.L, d.eps, dth, dsc, d.seed)
  switch (d.C / 128) {
    case 1: GLM(1); break;
    case 2: GLM(2); break;
    case 3: GLM(3); break;
    case 4: GLM(4); break;
    case 5: GLM(5); break;
    case 6: GLM(6); break;
    case 7: GLM(7); break;
    default: GLM(8); break;
  }
#undef GLM
  DMC_CUDA_OK(cudaGetLastError());
  return 0;
}

int launch_dit_gate_ln_mod_backward(const dmc_dit_glm_bwd_desc& d, cudaStream_t st) {
  DMC_REQUIRE(d.x && d.dh && d.scale && d.dx_in && d.dshift && d.dscale, "dit_gate_ln_mod_backward: null pointer argument");
  DMC_REQUIRE((d.y == nullptr) == (d.gate == nullptr) && (d.y == nullptr) == (d.dy == nullptr) && (d.y == nullptr) == (d.dgate == nullptr),
              "dit_gate_ln_mod_backward: y, gate, dy and dgate go together");
  DMC_REQUIRE(d.B > 0 && d.L > 0 && d.C % 128 == 0 && d.C >= 128 && d.C <= 1024 && d.mod_stride % 4 == 0 && d.gate_stride % 4 == 0,
              "dit_gate_ln_mod_backward: C=%d must be a multiple of 128 in [128, 1024]", d.C);
  const __nv_bfloat16* dh = reinterpret_cast<const __nv_bfloat16*>(d.dh);
  const __nv_bfloat16* y = reinterpret_cast<const __nv_bfloat16*>(d.y);
  __nv_bfloat16* dy = reinterpret_cast<__nv_bfloat16*>(d.dy);
  DMC_REQUIRE(d.drop_p >= 0.f && d.drop_p < 1.f, "dit_gate_ln_mod_backward: drop_p=%f", d.drop_p);
  const uint32_t dth = dropout_threshold(d.drop_p);
  const float dsc = d.drop_p > 0.f ? 1.0f / (1.0f - d.drop_p) : 1.0f;
  // row slices per image: enough CTAs to fill the chip (one CTA per image leaves 8 warps per SM at batch 128)
  const int S = (d.scratch != nullptr && d.L >= 64) ? DMC_DIT_GLM_BWD_SLICES : 1;
  const dim3 grid(d.B, S);
#define GLB(V) dit_gate_ln_mod_bwd_kernel<V><<<grid, 256, 0, st>>>(d.x, dh, d.dx_out, y, d.gate, d.scale, d.mod_stride, d.gate_stride, d.dx_in, dy, d.dgate, d.dshift, d.dscale, d.L, d.eps, dth, dsc, d.seed, d.scratch)
  switch (d.C / 128) {
    case 1: GLB(1); break;
    case 2: GLB(2); break;
    case 3: GLB(3); break;
    case 4: GLB(4); break;
    case 5: GLB(5); break;
    case 6: GLB(6); break;
    case 7: GLB(7); break;
    default: GLB(8); break;
  }
#undef GLB
  DMC_CUDA_OK(cudaGetLastError());
  if (S > 1) {
    const int total = d.B * d.C;
    dit_glm_bwd_reduce_kernel<<<(total + 255) / 256, 256, 0, st>>>(d.scratch, S, d.C, d.dgate, d.dshift, d.dscale, total);
    DMC_CUDA_OK(cudaGetLastError());
  }
  return 0;
}

// nn.GELU() (erf form), bf16 in / out, fp32 arithmetic; 8 elements per thread
__global__ void __launch_bounds__(256) gelu_fwd_kernel(const uint4* __restrict__ u, uint4* __restrict__ m, size_t n8,
                                                       uint32_t drop_thresh, float drop_scale, uint32_t seed) {
  const size_t i = static_cast<size_t>(blockIdx.x) * 256 + threadIdx.x;
  if (i >= n8) return;
  const uint4 a = u[i];
  const uint32_t w[4] = {a.x, a.y, a.z, a.w};
  uint32_t o[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const float2 f = unpack_bf16x2(w[k]);
    float g0 = 0.5f * f.x * (1.0f + erff(f.x * 0.70710678118654752f)), g1 = 0.5f * f.y * (1.0f + erff(f.y * 0.70710678118654752f));
    if (drop_thresh != 0u) {  // nn.Dropout after the activation (models/dit.py:97)
      const uint32_t hsh = dropout_hash2(seed, i * 8 + 2 * k);
      g0 = (hsh & 0xFFFFu) >= drop_thresh ? g0 * drop_scale : 0.f;
      g1 = (hsh >> 16) >= drop_thresh ? g1 * drop_scale : 0.f;
    }
    o[k] = pack_bf16x2(g0, g1);
  }
  m[i] = make_uint4(o[0], o[1], o[2], o[3]);
}

__global__ void __launch_bounds__(256) gelu_bwd_kernel(const uint4* __restrict__ u, const uint4* __restrict__ dm, uint4* __restrict__ du,
                                                       size_t n8, uint32_t drop_thresh, float drop_scale, uint32_t seed) {
  const size_t i = static_cast<size_t>(blockIdx.x) * 256 + threadIdx.x;
  if (i >= n8) return;
  const uint4 a = u[i], b = dm[i];
  const uint32_t w[4] = {a.x, a.y, a.z, a.w}, g[4] = {b.x, b.y, b.z, b.w};
  uint32_t o[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const float2 f = unpack_bf16x2(w[k]), d_ = unpack_bf16x2(g[k]);
    const float d0 = 0.5f * (1.0f + erff(f.x * 0.70710678118654752f)) + f.x * 0.3989422804014327f * __expf(-0.5f * f.x * f.x);
    const float d1 = 0.5f * (1.0f + erff(f.y * 0.70710678118654752f)) + f.y * 0.3989422804014327f * __expf(-0.5f * f.y * f.y);
    float r0 = d_.x * d0, r1 = d_.y * d1;
    if (drop_thresh != 0u) {
      const uint32_t hsh = dropout_hash2(seed, i * 8 + 2 * k);
      r0 = (hsh & 0xFFFFu) >= drop_thresh ? r0 * drop_scale : 0.f;
      r1 = (hsh >> 16) >= drop_thresh ? r1 * drop_scale : 0.f;
    }
    o[k] = pack_bf16x2(r0, r1);
  }
  du[i] = make_uint4(o[0], o[1], o[2], o[3]);
}

int launch_gelu_forward(const void* u, void* m, int64_t n, float drop_p, uint32_t seed, cudaStream_t st) {
  DMC_REQUIRE(u && m && n > 0 && n % 8 == 0, "gelu_forward: n=%lld must be a positive multiple of 8", static_cast<long long>(n));
  DMC_REQUIRE(drop_p >= 0.f && drop_p < 1.f, "gelu_forward: drop_p=%f", drop_p);
  const size_t n8 = static_cast<size_t>(n / 8);
  gelu_fwd_kernel<<<static_cast<unsigned>((n8 + 255) / 256), 256, 0, st>>>(reinterpret_cast<const uint4*>(u), reinterpret_cast<uint4*>(m), n8,
                                                                             dropout_threshold(drop_p),
                                                                             drop_p > 0.f ? 1.0f / (1.0f - drop_p) : 1.0f, seed);
  DMC_CUDA_OK(cudaGetLastError());
  return 0;
}

int launch_gelu_backward(const void* u, const void* dm, void* du, int64_t n, float drop_p, uint32_t seed, cudaStream_t st) {
  DMC_REQUIRE(u && dm && du && n > 0 && n % 8 == 0, "gelu_backward: n=%lld must be a positive multiple of 8", static_cast<long long>(n));
  DMC_REQUIRE(drop_p >= 0.f && drop_p < 1.f, "gelu_backward: drop_p=%f", drop_p);
  const size_t n8 = static_cast<size_t>(n / 8);
  gelu_bwd_kernel<<<static_cast<unsigned>((n8 + 255) / 256), 256, 0, st>>>(reinterpret_cast<const uint4*>(u), reinterpret_cast<const uint4*>(dm),
                                                                             reinterpret_cast<uint4*>(du), n8, dropout_threshold(drop_p),
                                                                             drop_p > 0.f ? 1.0f / (1.0f - drop_p) : 1.0f, seed);
  DMC_CUDA_OK(cudaGetLastError());
  return 0;
}

}  // namespace dmc
