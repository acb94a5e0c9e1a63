// Multi-head self-attention core softmax(Q K^T / sqrt(hd)) V, head dim 64, flash-style (online softmax, the L x L
// score matrix is never written to memory).  Replaces /root/reference/models/unet.py:88-96 (2 bmm + softmax + 2
// permute copies, L x L fp32 scores materialised) and the attention core of nn.MultiheadAttention in
// models/dit.py:94,123.
//
// v1 (this file): CUDA-core kernel, one thread per query row, K/V tiles broadcast from shared memory.  Attention
// matmuls are 2.8 % of the UNet's FLOPs; the tcgen05 version is the next step for this kernel.
#include "common.cuh"
#include "kernels.h"

namespace dmc {

constexpr int ATT_TQ = 128;  // queries (threads) per CTA
constexpr int ATT_TK = 32;   // keys per shared-memory tile

template <int ATT_HD>
__global__ void __launch_bounds__(ATT_TQ) attention_kernel(const __nv_bfloat16* __restrict__ qkv,
                                                           __nv_bfloat16* __restrict__ out, int L, int C, float scale_log2e) {
  __shared__ uint4 sK[ATT_TK * ATT_HD / 8];
  __shared__ uint4 sV[ATT_TK * ATT_HD / 8];
  const int n = blockIdx.z, h = blockIdx.y;
  const int qi = blockIdx.x * ATT_TQ + threadIdx.x;
  const bool active = qi < L;
  const size_t row_stride = static_cast<size_t>(3) * C;
  const __nv_bfloat16* base = qkv + static_cast<size_t>(n) * L * row_stride;

  float q[ATT_HD], acc[ATT_HD];
#pragma unroll
  for (int d = 0; d < ATT_HD; ++d) acc[d] = 0.f;
  if (active) {
    const uint4* qp = reinterpret_cast<const uint4*>(base + static_cast<size_t>(qi) * row_stride + h * ATT_HD);
#pragma unroll
    for (int v = 0; v < ATT_HD / 8; ++v) {
      uint4 u = __ldg(qp + v);
      uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        float2 f = unpack_bf16x2(w[k]);
        q[v * 8 + 2 * k] = f.x * scale_log2e;
        q[v * 8 + 2 * k + 1] = f.y * scale_log2e;
      }
    }
  } else {
#pragma unroll
    for (int d = 0; d < ATT_HD; ++d) q[d] = 0.f;
  }
  float m = -INFINITY, l = 0.f;

  for (int k0 = 0; k0 < L; k0 += ATT_TK) {
    __syncthreads();
    for (int v = threadIdx.x; v < ATT_TK * (ATT_HD / 8); v += ATT_TQ) {
      const int kr = v / (ATT_HD / 8), kv = v % (ATT_HD / 8);
      uint4 kk = make_uint4(0, 0, 0, 0), vv = make_uint4(0, 0, 0, 0);
      if (k0 + kr < L) {
        const __nv_bfloat16* rp = base + static_cast<size_t>(k0 + kr) * row_stride + h * ATT_HD;
        kk = __ldg(reinterpret_cast<const uint4*>(rp + C) + kv);
        vv = __ldg(reinterpret_cast<const uint4*>(rp + 2 * C) + kv);
      }
      sK[v] = kk;
      sV[v] = vv;
    }
    __syncthreads();
    float s[ATT_TK];
    float tmax = -INFINITY;
#pragma unroll
    for (int j = 0; j < ATT_TK; ++j) {
      float a = 0.f;
#pragma unroll
      for (int v = 0; v < ATT_HD / 8; ++v) {
        uint4 u = sK[j * (ATT_HD / 8) + v];
        uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          float2 f = unpack_bf16x2(w[k]);
          a = fmaf(q[v * 8 + 2 * k], f.x, a);
          a = fmaf(q[v * 8 + 2 * k + 1], f.y, a);
        }
      }
      s[j] = (k0 + j < L) ? a : -INFINITY;
      tmax = fmaxf(tmax, s[j]);
    }
    const float m_new = fmaxf(m, tmax);
    const float corr = exp2f(m - m_new);  // m = -inf on the first tile -> 0
    l *= corr;
#pragma unroll
    for (int d = 0; d < ATT_HD; ++d) acc[d] *= corr;
#pragma unroll
    for (int j = 0; j < ATT_TK; ++j) {
      const float pj = exp2f(s[j] - m_new);
      l += pj;
#pragma unroll
      for (int v = 0; v < ATT_HD / 8; ++v) {
        uint4 u = sV[j * (ATT_HD / 8) + v];
        uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          float2 f = unpack_bf16x2(w[k]);
          acc[v * 8 + 2 * k] = fmaf(pj, f.x, acc[v * 8 + 2 * k]);
          acc[v * 8 + 2 * k + 1] = fmaf(pj, f.y, acc[v * 8 + 2 * k + 1]);
        }
      }
    }
    m = m_new;
  }
  if (active) {
    const float inv = 1.0f / l;
    uint4* op = reinterpret_cast<uint4*>(out + (static_cast<size_t>(n) * L + qi) * C + h * ATT_HD);
#pragma unroll
    for (int v = 0; v < ATT_HD / 8; ++v) {
      uint4 u;
      u.x = pack_bf16x2(acc[v * 8] * inv, acc[v * 8 + 1] * inv);
      u.y = pack_bf16x2(acc[v * 8 + 2] * inv, acc[v * 8 + 3] * inv);
      u.z = pack_bf16x2(acc[v * 8 + 4] * inv, acc[v * 8 + 5] * inv);
      u.w = pack_bf16x2(acc[v * 8 + 6] * inv, acc[v * 8 + 7] * inv);
      op[v] = u;
    }
  }
}

// Split-bf16 mode: q, k, v are (hi, lo) bf16 pairs, all arithmetic in fp32 on the CUDA cores, the output is a pair again.
// One thread per query row, K / V tiles as fp32 in shared memory.
template <int ATT_HD>
__global__ void __launch_bounds__(ATT_TQ) attention_split_kernel(const __nv_bfloat16* __restrict__ qkv,
                                                                 const __nv_bfloat16* __restrict__ qkv_lo,
                                                                 __nv_bfloat16* __restrict__ out,
                                                                 __nv_bfloat16* __restrict__ out_lo, int L, int C,
                                                                 float scale_log2e) {
  __shared__ float sK[ATT_TK * ATT_HD];
  __shared__ float sV[ATT_TK * ATT_HD];
  const int n = blockIdx.z, h = blockIdx.y;
  const int qi = blockIdx.x * ATT_TQ + threadIdx.x;
  const bool active = qi < L;
  const size_t row_stride = static_cast<size_t>(3) * C;
  const size_t img = static_cast<size_t>(n) * L * row_stride;
  float q[ATT_HD], acc[ATT_HD];
#pragma unroll
  for (int d = 0; d < ATT_HD; ++d) {
    acc[d] = 0.f;
    q[d] = 0.f;
  }
  if (active) {
    const size_t o = img + static_cast<size_t>(qi) * row_stride + h * ATT_HD;
#pragma unroll
    for (int d = 0; d < ATT_HD; ++d) q[d] = (__bfloat162float(qkv[o + d]) + __bfloat162float(qkv_lo[o + d])) * scale_log2e;
  }
  float m = -INFINITY, l = 0.f;
  for (int k0 = 0; k0 < L; k0 += ATT_TK) {
    __syncthreads();
    for (int e = threadIdx.x; e < ATT_TK * ATT_HD; e += ATT_TQ) {
      const int kr = e / ATT_HD, d = e % ATT_HD;
      float kk = 0.f, vv = 0.f;
      if (k0 + kr < L) {
        const size_t o = img + static_cast<size_t>(k0 + kr) * row_stride + h * ATT_HD + d;
        kk = __bfloat162float(qkv[o + C]) + __bfloat162float(qkv_lo[o + C]);
        vv = __bfloat162float(qkv[o + 2 * C]) + __bfloat162float(qkv_lo[o + 2 * C]);
      }
      sK[e] = kk;
      sV[e] = vv;
    }
    __syncthreads();
    float s[ATT_TK];
    float tmax = -INFINITY;
#pragma unroll 4
    for (int j = 0; j < ATT_TK; ++j) {
      float a = 0.f;
#pragma unroll
      for (int d = 0; d < ATT_HD; ++d) a = fmaf(q[d], sK[j * ATT_HD + d], a);
      s[j] = (k0 + j < L) ? a : -INFINITY;
      tmax = fmaxf(tmax, s[j]);
    }
    const float m_new = fmaxf(m, tmax);
    const float corr = exp2f(m - m_new);
    l *= corr;
#pragma unroll
    for (int d = 0; d < ATT_HD; ++d) acc[d] *= corr;
#pragma unroll 4
    for (int j = 0; j < ATT_TK; ++j) {
      const float pj = exp2f(s[j] - m_new);
      l += pj;
#pragma unroll
      for (int d = 0; d < ATT_HD; ++d) acc[d] = fmaf(pj, sV[j * ATT_HD + d], acc[d]);
    }
    m = m_new;
  }
  if (active) {
    const float inv = 1.0f / l;
    const size_t o = (static_cast<size_t>(n) * L + qi) * C + h * ATT_HD;
#pragma unroll
    for (int d = 0; d < ATT_HD; ++d) {
      const float v = acc[d] * inv;
      const __nv_bfloat16 hi = __float2bfloat16(v);
      out[o + d] = hi;
      out_lo[o + d] = __float2bfloat16(v - __bfloat162float(hi));
    }
  }
}

int launch_attention(const dmc_attn_desc& d, cudaStream_t st) {
  DMC_REQUIRE(d.qkv && d.out && d.B > 0 && d.L > 0 && d.heads > 0, "attention: bad arguments");
  const int hd = d.C / d.heads;
  DMC_REQUIRE(d.C == d.heads * hd && (hd == 32 || hd == 64), "attention: head dim must be 32 or 64 (C=%d, heads=%d)", d.C,
              d.heads);
  dim3 grid((d.L + ATT_TQ - 1) / ATT_TQ, d.heads, d.B);
  const float scale_log2e = 1.4426950408889634f / sqrtf(static_cast<float>(hd));
  const __nv_bfloat16* qkv = reinterpret_cast<const __nv_bfloat16*>(d.qkv);
  __nv_bfloat16* out = reinterpret_cast<__nv_bfloat16*>(d.out);
  if (d.qkv_lo != nullptr || d.out_lo != nullptr) {
    DMC_REQUIRE(d.qkv_lo && d.out_lo, "attention: split-bf16 mode needs both qkv_lo and out_lo");
    const __nv_bfloat16* ql = reinterpret_cast<const __nv_bfloat16*>(d.qkv_lo);
    __nv_bfloat16* ol = reinterpret_cast<__nv_bfloat16*>(d.out_lo);
    if (hd == 64) attention_split_kernel<64><<<grid, ATT_TQ, 0, st>>>(qkv, ql, out, ol, d.L, d.C, scale_log2e);
    else attention_split_kernel<32><<<grid, ATT_TQ, 0, st>>>(qkv, ql, out, ol, d.L, d.C, scale_log2e);
    DMC_CUDA_OK(cudaGetLastError());
    return 0;
  }
  if (hd == 64) attention_kernel<64><<<grid, ATT_TQ, 0, st>>>(qkv, out, d.L, d.C, scale_log2e);
  else attention_kernel<32><<<grid, ATT_TQ, 0, st>>>(qkv, out, d.L, d.C, scale_log2e);
  DMC_CUDA_OK(cudaGetLastError());
  return 0;
}

}  // namespace dmc
