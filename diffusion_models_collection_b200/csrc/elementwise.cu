// Memory-bound kernels of the UNet forward: conditioning table, input convolution (fp32 NCHW -> bf16 NHWC),
// GroupNorm statistics / apply(+SiLU, + channel concat), nearest 2x upsample.  All 128-bit vectorised.
#include <stdlib.h>

#include "common.cuh"
#include "kernels.h"

namespace dmc {

// =============================================================================================
// Conditioning: sinusoid -> Linear -> SiLU -> Linear -> SiLU -> all time_mlp projections (+ label table)
// /root/reference/models/unet.py:18-25, 167-172, 40-48, 65-68, 256-260
// =============================================================================================
__global__ void __launch_bounds__(256) cond_hidden_kernel(const int64_t* __restrict__ t, const float* __restrict__ freqs,
                                                          const float* __restrict__ w1, const float* __restrict__ b1,
                                                          float* __restrict__ h1, int half, int temb) {
  extern __shared__ float emb[];  // 2*half
  const int r = blockIdx.x;
  const float tf = static_cast<float>(t[r]);  // int64 * fp32 promotes to fp32 (unet.py:23)
  for (int i = threadIdx.x; i < half; i += blockDim.x) {
    float arg = __fmul_rn(tf, freqs[i]);
    emb[i] = sinf(arg);
    emb[half + i] = cosf(arg);
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  const int K = 2 * half;
  for (int j = blockIdx.y * nw + warp; j < temb; j += gridDim.y * nw) {
    float acc = 0.f;
    for (int k = lane; k < K; k += 32) acc = fmaf(w1[static_cast<size_t>(j) * K + k], emb[k], acc);
#pragma unroll
    for (int o = 16; o; o >>= 1) acc += __shfl_xor_sync(0xFFFFFFFFu, acc, o);
    if (lane == 0) {
      float v = acc + b1[j];
      h1[static_cast<size_t>(r) * temb + j] = v / (1.0f + expf(-v));  // SiLU feeding time_embed.3
    }
  }
}

// out[r, j] = act(b[j] + sum_k w[j, k] * in[r, k]);  act = SiLU when silu_out (the shared SiLU(t_emb) of :40-42)
__global__ void __launch_bounds__(256) cond_linear_kernel(const float* __restrict__ in, const float* __restrict__ w,
                                                          const float* __restrict__ b, float* __restrict__ out, int K,
                                                          int N, int silu_out) {
  const int r = blockIdx.x;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  const float* row_in = in + static_cast<size_t>(r) * K;
  for (int j = blockIdx.y * nw + warp; j < N; j += gridDim.y * nw) {
    float acc = 0.f;
    for (int k = lane; k < K; k += 32) acc = fmaf(w[static_cast<size_t>(j) * K + k], row_in[k], acc);
#pragma unroll
    for (int o = 16; o; o >>= 1) acc += __shfl_xor_sync(0xFFFFFFFFu, acc, o);
    if (lane == 0) {
      float v = acc + b[j];
      out[static_cast<size_t>(r) * N + j] = silu_out ? v / (1.0f + expf(-v)) : v;
    }
  }
}

// One warp per output column j keeps its weight row in registers and walks all R rows of SiLU(t_emb).
template <int KPL>  // K / 32
__global__ void __launch_bounds__(256) cond_project_kernel(const float* __restrict__ st, const float* __restrict__ w,
                                                           const float* __restrict__ b, float* __restrict__ out, int R,
                                                           int N) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int j = blockIdx.x * (blockDim.x >> 5) + warp;
  if (j >= N) return;
  constexpr int K = KPL * 32;
  float wr[KPL];
#pragma unroll
  for (int i = 0; i < KPL; ++i) wr[i] = w[static_cast<size_t>(j) * K + lane + 32 * i];
  const float bj = b[j];
  for (int r = 0; r < R; ++r) {
    const float* s = st + static_cast<size_t>(r) * K;
    float acc = 0.f;
#pragma unroll
    for (int i = 0; i < KPL; ++i) acc = fmaf(wr[i], s[lane + 32 * i], acc);
#pragma unroll
    for (int o = 16; o; o >>= 1) acc += __shfl_xor_sync(0xFFFFFFFFu, acc, o);
    if (lane == 0) out[static_cast<size_t>(r) * N + j] = acc + bj;
  }
}

__global__ void __launch_bounds__(256) cond_expand_kernel(const float* __restrict__ cond_t, const float* __restrict__ ytab,
                                                          const int64_t* __restrict__ y, float* __restrict__ cond, int B,
                                                          int ncols, int uniform_t, int num_classes) {
  const int n4 = ncols >> 2;
  const size_t total = static_cast<size_t>(B) * n4;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int n = static_cast<int>(i / n4), c = static_cast<int>(i % n4);
    float4 v = reinterpret_cast<const float4*>(cond_t + static_cast<size_t>(uniform_t ? 0 : n) * ncols)[c];
    if (ytab != nullptr && y != nullptr) {
      long long lab = y[n];
      lab = lab < 0 ? 0 : (lab > num_classes ? num_classes : lab);  // torch.clamp(y, 0, num_classes), unet.py:257
      float4 u = reinterpret_cast<const float4*>(ytab + static_cast<size_t>(lab) * ncols)[c];
      v.x += u.x; v.y += u.y; v.z += u.z; v.w += u.w;
    }
    reinterpret_cast<float4*>(cond + static_cast<size_t>(n) * ncols)[c] = v;
  }
}

int cond_num_launches(const dmc_cond_desc&) { return 4; }

int launch_cond(const dmc_cond_desc& d, cudaStream_t st) {
  DMC_REQUIRE(d.t && d.freqs && d.w1 && d.b1 && d.w2 && d.b2 && d.wt_all && d.bt_all && d.scratch && d.cond,
              "cond: null pointer argument");
  DMC_REQUIRE(d.B > 0 && d.ncols % 4 == 0 && d.temb > 0 && d.half > 0, "cond: unsupported shape (B=%d ncols=%d temb=%d)",
              d.B, d.ncols, d.temb);
  const int R = d.uniform_t ? 1 : d.B;
  float* h1 = d.scratch;
  float* sil = d.scratch + static_cast<size_t>(R) * d.temb;
  float* cond_t = d.scratch + 2 * static_cast<size_t>(R) * d.temb;
  cond_hidden_kernel<<<dim3(R, 8), 256, 2 * d.half * sizeof(float), st>>>(d.t, d.freqs, d.w1, d.b1, h1, d.half, d.temb);
  cond_linear_kernel<<<dim3(R, 8), 256, 0, st>>>(h1, d.w2, d.b2, sil, d.temb, d.temb, 1);
  const int pb = (d.ncols + 7) / 8;
  switch (d.temb) {  // weight row in registers (temb / 32 floats per lane), all R rows streamed past it
    case 128: cond_project_kernel<4><<<pb, 256, 0, st>>>(sil, d.wt_all, d.bt_all, cond_t, R, d.ncols); break;
    case 256: cond_project_kernel<8><<<pb, 256, 0, st>>>(sil, d.wt_all, d.bt_all, cond_t, R, d.ncols); break;
    case 512: cond_project_kernel<16><<<pb, 256, 0, st>>>(sil, d.wt_all, d.bt_all, cond_t, R, d.ncols); break;
    case 1024: cond_project_kernel<32><<<pb, 256, 0, st>>>(sil, d.wt_all, d.bt_all, cond_t, R, d.ncols); break;
    default: cond_linear_kernel<<<dim3(R, 8), 256, 0, st>>>(sil, d.wt_all, d.bt_all, cond_t, d.temb, d.ncols, 0); break;
  }
  size_t total = static_cast<size_t>(d.B) * (d.ncols / 4);
  int blocks = static_cast<int>(std::min<size_t>((total + 255) / 256, static_cast<size_t>(num_sms()) * 8));
  cond_expand_kernel<<<blocks, 256, 0, st>>>(cond_t, d.ytab, d.y, d.cond, d.B, d.ncols, d.uniform_t, d.num_classes);
  DMC_CUDA_OK(cudaGetLastError());
  return 0;
}

// =============================================================================================
// Stem: 3x3 conv with Cin <= 4 from fp32 NCHW straight to bf16 NHWC (models/unet.py:188, 263)
// =============================================================================================
// Memory-bound (AI ~ 26 FLOP/B): fp32 CUDA-core FMAs.  lane = pixel (coalesced fp32 NCHW reads along W), every thread
// keeps the 9*Cin taps of TWO pixels in registers and walks the output channels 16 at a time; the weights are
// warp-uniform broadcast reads from shared memory (8 FMAs per 128-bit LDS), bf16 NHWC stores are full 32-byte sectors.
template <int CIN>
__global__ void __launch_bounds__(256) stem_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                   const float* __restrict__ bias, __nv_bfloat16* __restrict__ out,
                                                   __nv_bfloat16* __restrict__ out_lo, int x_batch, int B, int H, int W,
                                                   int Cout) {
  constexpr int K = CIN * 9;
  extern __shared__ float sw[];  // [K][Cout] transposed weights, then bias[Cout]
  // coalesced reads of the [Cout][K] parameter, transposed on the way into shared memory (the strided form -- one cache line per
  // lane -- cost a third of a block's life); the grid is persistent so that the 14 KB are staged once per SM, not once per 512 pixels
  for (int i = threadIdx.x; i < K * Cout; i += blockDim.x) {
    const int c = i / K, k = i % K;
    sw[k * Cout + c] = __ldg(w + i);
  }
  float* sb = sw + K * Cout;
  for (int i = threadIdx.x; i < Cout; i += blockDim.x) sb[i] = bias[i];
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const size_t total = static_cast<size_t>(B) * H * W;
  for (size_t blk = blockIdx.x; blk * 512 < total; blk += gridDim.x) {
  const size_t base = (blk * 8 + warp) * 64;
  float xv[2][K];
  size_t pix[2];
#pragma unroll
  for (int u = 0; u < 2; ++u) {
    pix[u] = base + u * 32 + lane;
    const bool ok = pix[u] < total;
    const size_t pp = ok ? pix[u] : 0;
    const int ww = static_cast<int>(pp % W);
    const int hh = static_cast<int>((pp / W) % H);
    const int n = static_cast<int>(pp / (static_cast<size_t>(W) * H));
    const float* xin = x + static_cast<size_t>(n % x_batch) * CIN * H * W;
#pragma unroll
    for (int ci = 0; ci < CIN; ++ci)
#pragma unroll
      for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int t = 0; t < 3; ++t) {
          const int ih = hh + r - 1, iw = ww + t - 1;
          const bool in = ok && ih >= 0 && ih < H && iw >= 0 && iw < W;
          xv[u][(ci * 3 + r) * 3 + t] = in ? __ldg(xin + (static_cast<size_t>(ci) * H + ih) * W + iw) : 0.f;
        }
  }
  for (int c0 = 0; c0 < Cout; c0 += 16) {
    float acc[2][16];
#pragma unroll
    for (int j = 0; j < 16; ++j) acc[0][j] = acc[1][j] = sb[c0 + j];
#pragma unroll
    for (int k = 0; k < K; ++k) {
      const float4* wk = reinterpret_cast<const float4*>(sw + k * Cout + c0);
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const float4 wv = wk[q];
        acc[0][4 * q] = fmaf(xv[0][k], wv.x, acc[0][4 * q]);
        acc[0][4 * q + 1] = fmaf(xv[0][k], wv.y, acc[0][4 * q + 1]);
        acc[0][4 * q + 2] = fmaf(xv[0][k], wv.z, acc[0][4 * q + 2]);
        acc[0][4 * q + 3] = fmaf(xv[0][k], wv.w, acc[0][4 * q + 3]);
        acc[1][4 * q] = fmaf(xv[1][k], wv.x, acc[1][4 * q]);
        acc[1][4 * q + 1] = fmaf(xv[1][k], wv.y, acc[1][4 * q + 1]);
        acc[1][4 * q + 2] = fmaf(xv[1][k], wv.z, acc[1][4 * q + 2]);
        acc[1][4 * q + 3] = fmaf(xv[1][k], wv.w, acc[1][4 * q + 3]);
      }
    }
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      if (pix[u] < total) {
        uint4* o = reinterpret_cast<uint4*>(out + pix[u] * Cout + c0);
        o[0] = make_uint4(pack_bf16x2(acc[u][0], acc[u][1]), pack_bf16x2(acc[u][2], acc[u][3]),
                          pack_bf16x2(acc[u][4], acc[u][5]), pack_bf16x2(acc[u][6], acc[u][7]));
        o[1] = make_uint4(pack_bf16x2(acc[u][8], acc[u][9]), pack_bf16x2(acc[u][10], acc[u][11]),
                          pack_bf16x2(acc[u][12], acc[u][13]), pack_bf16x2(acc[u][14], acc[u][15]));
        if (out_lo != nullptr) {  // split-bf16 mode: the rounding remainder as a second bf16 tensor
          uint32_t lo[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float2 hi = unpack_bf16x2(pack_bf16x2(acc[u][2 * j], acc[u][2 * j + 1]));
            lo[j] = pack_bf16x2(acc[u][2 * j] - hi.x, acc[u][2 * j + 1] - hi.y);
          }
          uint4* ol = reinterpret_cast<uint4*>(out_lo + pix[u] * Cout + c0);
          ol[0] = make_uint4(lo[0], lo[1], lo[2], lo[3]);
          ol[1] = make_uint4(lo[4], lo[5], lo[6], lo[7]);
        }
      }
    }
  }
  }
}

int launch_stem(const dmc_stem_desc& d, cudaStream_t st) {
  DMC_REQUIRE(d.x && d.weight && d.bias && d.out, "stem: null pointer argument");
  DMC_REQUIRE(d.Cin >= 1 && d.Cin <= 4 && d.Cout % 16 == 0 && d.x_batch > 0 && d.B > 0, "stem: unsupported shape");
  const size_t total = static_cast<size_t>(d.B) * d.H * d.W;
  const int blocks = static_cast<int>(std::min<size_t>((total + 511) / 512, static_cast<size_t>(num_sms()) * 2));
  const size_t smem = (static_cast<size_t>(d.Cin) * 9 * d.Cout + d.Cout) * sizeof(float);
  DMC_REQUIRE(smem <= 48 * 1024, "stem: Cout=%d too large", d.Cout);
  __nv_bfloat16* out = reinterpret_cast<__nv_bfloat16*>(d.out);
  __nv_bfloat16* lo = reinterpret_cast<__nv_bfloat16*>(d.out_lo);
  switch (d.Cin) {
    case 1: stem_kernel<1><<<blocks, 256, smem, st>>>(d.x, d.weight, d.bias, out, lo, d.x_batch, d.B, d.H, d.W, d.Cout); break;
    case 2: stem_kernel<2><<<blocks, 256, smem, st>>>(d.x, d.weight, d.bias, out, lo, d.x_batch, d.B, d.H, d.W, d.Cout); break;
    case 3: stem_kernel<3><<<blocks, 256, smem, st>>>(d.x, d.weight, d.bias, out, lo, d.x_batch, d.B, d.H, d.W, d.Cout); break;
    default: stem_kernel<4><<<blocks, 256, smem, st>>>(d.x, d.weight, d.bias, out, lo, d.x_batch, d.B, d.H, d.W, d.Cout); break;
  }
  DMC_CUDA_OK(cudaGetLastError());
  return 0;
}

// =============================================================================================
// GroupNorm statistics: (sum, sumsq) per image, per 128-pixel slab (slot), per 8-channel block
// =============================================================================================
constexpr int GN_SLAB = 128;  // pixels per CTA

__global__ void __launch_bounds__(256) gn_stats_kernel(const uint4* __restrict__ src, const uint4* __restrict__ src_lo,
                                                       float* __restrict__ stats, int HW, int C8 /* C/8 */,
                                                       int rows /* blockDim / C8 */) {
  extern __shared__ float red[];  // [rows][C8][2]
  const int n = blockIdx.y;
  const int p0 = blockIdx.x * GN_SLAB;
  const int p1 = min(p0 + GN_SLAB, HW);
  const int cb = threadIdx.x % C8, r = threadIdx.x / C8;
  float s = 0.f, ss = 0.f;
  if (r < rows) {
    for (int p = p0 + r; p < p1; p += rows) {
      uint4 v = src[(static_cast<size_t>(n) * HW + p) * C8 + cb];
      uint32_t u[4] = {v.x, v.y, v.z, v.w};
      uint4 vl = make_uint4(0, 0, 0, 0);
      if (src_lo != nullptr) vl = src_lo[(static_cast<size_t>(n) * HW + p) * C8 + cb];
      const uint32_t ul[4] = {vl.x, vl.y, vl.z, vl.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float2 f = unpack_bf16x2(u[j]);
        const float2 fl = unpack_bf16x2(ul[j]);
        f.x += fl.x;
        f.y += fl.y;
        s += f.x + f.y;
        ss = fmaf(f.x, f.x, ss);
        ss = fmaf(f.y, f.y, ss);
      }
    }
    red[(r * C8 + cb) * 2] = s;
    red[(r * C8 + cb) * 2 + 1] = ss;
  }
  __syncthreads();
  if (r == 0) {
    for (int k = 1; k < rows; ++k) {
      s += red[(k * C8 + cb) * 2];
      ss += red[(k * C8 + cb) * 2 + 1];
    }
    // slot = this 128-pixel slab: written once, summed in order by the consumer (deterministic)
    float2* dst = reinterpret_cast<float2*>(stats) + (static_cast<size_t>(n) * gridDim.x + blockIdx.x) * C8 + cb;
    *dst = make_float2(s, ss);
  }
}

int launch_gn_stats(const dmc_gn_stats_desc& d, cudaStream_t st) {
  DMC_REQUIRE(d.src && d.stats && d.B > 0 && d.HW > 0, "gn_stats: bad arguments");
  DMC_REQUIRE(d.C % 8 == 0 && d.C / 8 <= 256, "gn_stats: C=%d unsupported", d.C);
  const int C8 = d.C / 8;
  const int rows = std::max(1, 256 / C8);
  const int threads = rows * C8;
  dim3 grid((d.HW + GN_SLAB - 1) / GN_SLAB, d.B);
  gn_stats_kernel<<<grid, threads, static_cast<size_t>(threads) * 2 * sizeof(float), st>>>(
      reinterpret_cast<const uint4*>(d.src), reinterpret_cast<const uint4*>(d.src_lo), d.stats, d.HW, C8, rows);
  DMC_CUDA_OK(cudaGetLastError());
  return 0;
}

// =============================================================================================
// GroupNorm apply (+SiLU) over the concatenation of up to two sources -> one bf16 NHWC tensor
// =============================================================================================
struct GnApplyArgs {
  const uint4* src0;
  const uint4* src1;
  const float* stats0;
  const float* stats1;
  const float* gamma;
  const float* beta;
  uint4* out;
  const uint4* lo0;  // split-bf16 mode: low parts of the sources / of the output (all or none)
  const uint4* lo1;
  uint4* out_lo;
  int HW, C0_8, C1_8, groups;
  int slots0, slots1;
  int slab;               // pixels per CTA
  float eps;
  int silu;
  float drop_scale;       // training: 1 / (1 - p) for kept elements
  uint32_t drop_thresh, seed;
  const uint32_t* seed_dev;  // optional per-step seed in device memory (added to `seed`): lets a captured graph draw new masks
};

// Phase 1 (per CTA, cheap): warp g reduces the partial sums of group g (slots x 8-channel blocks, possibly from both
// sources of a concat) with lane-strided loads and a fixed shuffle tree -> mean / rstd, bit-reproducible.
// Phase 2: every thread owns ONE 8-channel block (its 8 scales / shifts live in registers) and walks the pixels of the
// slab: one 128-bit load, 8 FMAs (+ SiLU), one 128-bit store per pixel, four pixels in flight.
template <bool LO>
__global__ void __launch_bounds__(256) gn_apply_kernel(GnApplyArgs a) {
  __shared__ float s_mean[32], s_rstd[32];
  const int C8 = a.C0_8 + a.C1_8;
  const int n = blockIdx.y;
  const int gs8 = C8 / a.groups;  // 8-channel blocks per group
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int g = warp; g < a.groups; g += 8) {
    const int blo = g * gs8, bhi = blo + gs8;
    float s = 0.f, ss = 0.f;
    {  // blocks of this group that live in source 0
      const int lo = min(blo, a.C0_8), hi = min(bhi, a.C0_8), nb = hi - lo;
      const float2* base = reinterpret_cast<const float2*>(a.stats0) + static_cast<size_t>(n) * a.slots0 * a.C0_8;
      for (int e = lane; e < nb * a.slots0; e += 32) {
        const float2 v = __ldg(base + static_cast<size_t>(e / nb) * a.C0_8 + lo + e % nb);
        s += v.x;
        ss += v.y;
      }
    }
    if (a.C1_8 > 0) {  // ... and in source 1
      const int lo = max(blo, a.C0_8) - a.C0_8, hi = max(bhi, a.C0_8) - a.C0_8, nb = hi - lo;
      const float2* base = reinterpret_cast<const float2*>(a.stats1) + static_cast<size_t>(n) * a.slots1 * a.C1_8;
      for (int e = lane; e < nb * a.slots1; e += 32) {
        const float2 v = __ldg(base + static_cast<size_t>(e / nb) * a.C1_8 + lo + e % nb);
        s += v.x;
        ss += v.y;
      }
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) {
      s += __shfl_xor_sync(0xFFFFFFFFu, s, o);
      ss += __shfl_xor_sync(0xFFFFFFFFu, ss, o);
    }
    if (lane == 0) {
      const float inv_cnt = 1.0f / (static_cast<float>(gs8 * 8) * static_cast<float>(a.HW));
      const float mean = s * inv_cnt;
      const float var = fmaxf(ss * inv_cnt - mean * mean, 0.f);
      s_mean[g] = mean;
      s_rstd[g] = rsqrtf(var + a.eps);
    }
  }
  __syncthreads();
  const int cb = threadIdx.x % C8, r0 = threadIdx.x / C8;
  const int rpi = blockDim.x / C8;  // pixel rows per iteration
  if (r0 >= rpi) return;
  float sc[8], sh[8];
  {
    const int g = cb / gs8;
    const float mean = s_mean[g], rstd = s_rstd[g];
    const float4* g4 = reinterpret_cast<const float4*>(a.gamma + cb * 8);
    const float4* b4 = reinterpret_cast<const float4*>(a.beta + cb * 8);
    const float4 ga = __ldg(g4), gb = __ldg(g4 + 1), ba = __ldg(b4), bb = __ldg(b4 + 1);
    const float gam[8] = {ga.x, ga.y, ga.z, ga.w, gb.x, gb.y, gb.z, gb.w};
    const float bet[8] = {ba.x, ba.y, ba.z, ba.w, bb.x, bb.y, bb.z, bb.w};
#pragma unroll
    const float fold = a.silu ? 0.5f : 1.0f;  // SiLU(y) = h + h tanh(h), h = y / 2: fold the 1/2 into scale and shift
    for (int j = 0; j < 8; ++j) {
      const float s1 = rstd * gam[j];
      sc[j] = fold * s1;
      sh[j] = fold * (bet[j] - mean * s1);
    }
  }
  const bool first = cb < a.C0_8;
  const uint4* src = first ? a.src0 + static_cast<size_t>(n) * a.HW * a.C0_8 + cb
                           : a.src1 + static_cast<size_t>(n) * a.HW * a.C1_8 + (cb - a.C0_8);
  const int sstride = first ? a.C0_8 : a.C1_8;
  uint4* dst = a.out + static_cast<size_t>(n) * a.HW * C8 + cb;
  const uint4* src_lo = nullptr;
  uint4* dst_lo = nullptr;
  if (LO) {
    src_lo = first ? a.lo0 + static_cast<size_t>(n) * a.HW * a.C0_8 + cb
                   : a.lo1 + static_cast<size_t>(n) * a.HW * a.C1_8 + (cb - a.C0_8);
    dst_lo = a.out_lo + static_cast<size_t>(n) * a.HW * C8 + cb;
  }
  const int p0 = blockIdx.x * a.slab;
  const int p1 = min(p0 + a.slab, a.HW);
  for (int pb = p0 + r0; pb < p1; pb += 4 * rpi) {
    uint4 in[4], inl[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int p = pb + u * rpi;
      if (p < p1) {
        in[u] = __ldg(src + static_cast<size_t>(p) * sstride);
        if (LO) inl[u] = __ldg(src_lo + static_cast<size_t>(p) * sstride);
      }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int p = pb + u * rpi;
      if (p < p1) {
        const uint32_t w[4] = {in[u].x, in[u].y, in[u].z, in[u].w};
        const uint32_t wl[4] = {inl[u].x, inl[u].y, inl[u].z, inl[u].w};
        uint32_t o[4], ol[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          float2 f = unpack_bf16x2(w[j]);
          if (LO) {
            const float2 fl = unpack_bf16x2(wl[j]);
            f.x += fl.x;
            f.y += fl.y;
          }
          float y0 = fmaf(f.x, sc[2 * j], sh[2 * j]);
          float y1 = fmaf(f.y, sc[2 * j + 1], sh[2 * j + 1]);
          if (a.silu) {
            if (LO) {  // the accuracy mode keeps the exact-to-fp32 SiLU (ex2 + rcp); y is y/2 here (scale folded)
              y0 = __fdividef(2.0f * y0, 1.0f + __expf(-2.0f * y0));
              y1 = __fdividef(2.0f * y1, 1.0f + __expf(-2.0f * y1));
            } else {
              y0 = silu_from_half(y0);
              y1 = silu_from_half(y1);
            }
          }
          if (a.drop_thresh != 0u) {  // dropout after SiLU (ResidualBlock.conv2, models/unet.py:53), training only
            const uint64_t idx = ((static_cast<uint64_t>(n) * a.HW + p) * C8 + cb) * 8 + 2 * j;
            const uint32_t seed = a.seed + (a.seed_dev ? __ldg(a.seed_dev) : 0u);
            const uint32_t h = dropout_hash2(seed, idx);
            y0 = (h & 0xFFFFu) >= a.drop_thresh ? y0 * a.drop_scale : 0.f;
            y1 = (h >> 16) >= a.drop_thresh ? y1 * a.drop_scale : 0.f;
          }
          o[j] = pack_bf16x2(y0, y1);
          if (LO) {
            const float2 hi = unpack_bf16x2(o[j]);
            ol[j] = pack_bf16x2(y0 - hi.x, y1 - hi.y);
          }
        }
        dst[static_cast<size_t>(p) * C8] = make_uint4(o[0], o[1], o[2], o[3]);
        if (LO) dst_lo[static_cast<size_t>(p) * C8] = make_uint4(ol[0], ol[1], ol[2], ol[3]);
      }
    }
  }
}

// =============================================================================================
// GroupNorm coefficients: out[n, c] = (scale, shift) with scale = rstd_g * gamma_c, shift = beta_c - mean_g * scale.  One CTA per
// image; phase 1 is gn_apply_kernel's (warp g reduces the partial sums of group g: lane-strided loads, fixed shuffle tree), so
// the statistics -- and fma(x, scale, shift) -- are bit-identical to the stand-alone pass.  Consumed by the A-operand affine
// transform of the 1x1 convolution that follows (dmc_conv_desc.a_affine).
// =============================================================================================
__global__ void __launch_bounds__(256) gn_coeff_kernel(const float* __restrict__ stats, const float* __restrict__ gamma,
                                                       const float* __restrict__ beta, float2* __restrict__ out, int slots, int HW,
                                                       int C, int groups, float eps) {
  __shared__ float s_mean[32], s_rstd[32];
  const int n = blockIdx.x, C8 = C >> 3, gs8 = C8 / groups;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int g = warp; g < groups; g += 8) {
    const int blo = g * gs8;
    float s = 0.f, ss = 0.f;
    const float2* base = reinterpret_cast<const float2*>(stats) + static_cast<size_t>(n) * slots * C8;
    for (int e = lane; e < gs8 * slots; e += 32) {
      const float2 v = __ldg(base + static_cast<size_t>(e / gs8) * C8 + blo + e % gs8);
      s += v.x;
      ss += v.y;
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) {
      s += __shfl_xor_sync(0xFFFFFFFFu, s, o);
      ss += __shfl_xor_sync(0xFFFFFFFFu, ss, o);
    }
    if (lane == 0) {
      const float inv_cnt = 1.0f / (static_cast<float>(gs8 * 8) * static_cast<float>(HW));
      const float mean = s * inv_cnt;
      const float var = fmaxf(ss * inv_cnt - mean * mean, 0.f);
      s_mean[g] = mean;
      s_rstd[g] = rsqrtf(var + eps);
    }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const int g = (c >> 3) / gs8;
    const float s1 = s_rstd[g] * __ldg(gamma + c);
    out[static_cast<size_t>(n) * C + c] = make_float2(s1, __ldg(beta + c) - s_mean[g] * s1);
  }
}

int launch_gn_coeff(const dmc_gn_coeff_desc& d, cudaStream_t st) {
  gn_coeff_kernel<<<d.B, 256, 0, st>>>(d.stats, d.gamma, d.beta, reinterpret_cast<float2*>(d.out), d.stats_slots, d.HW, d.C,
                                        d.groups, d.eps);
  DMC_CUDA_OK(cudaGetLastError());
  return 0;
}

// =============================================================================================
// Input convolution, tensor-core form (see dmc_stem_cols_desc): gather the 3x3 neighbourhood of every pixel of the fp32 NCHW
// input into one 64-channel bf16 NHWC row [taps as bf16 | their rounding remainders | 0]; the 1x1 tcgen05 GEMM does the rest.
// thread = (pixel, 16-byte chunk of its row): the 8 lanes of a pixel write its 128-byte row as one full line; the fp32
// reads (up to 2 taps x Cin per lane) come from L1 / L2 -- the whole input is 12 KB per image.
// =============================================================================================
__global__ void __launch_bounds__(256) stem_cols_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ out, int x_batch,
                                                        int B, int Cin, int H, int W) {
  const size_t gid = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  const size_t pix = gid >> 3;
  const int chunk = static_cast<int>(gid & 7);  // columns [8 chunk, 8 chunk + 8)
  if (pix >= static_cast<size_t>(B) * H * W) return;
  const int ww = static_cast<int>(pix % W), hh = static_cast<int>((pix / W) % H);
  const int n = static_cast<int>(pix / (static_cast<size_t>(W) * H));
  const float* xin = x + static_cast<size_t>(n % x_batch) * Cin * H * W;
  const int K = 9 * Cin;
  uint32_t o[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    float v[2];
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      const int col = chunk * 8 + 2 * j + e;
      const int k = col < K ? col : col - K;  // the tap this column carries (hi part, then lo part)
      float val = 0.f;
      if (col < 2 * K) {
        const int tap = k / Cin, ci = k % Cin;
        const int ih = hh + tap / 3 - 1, iw = ww + tap % 3 - 1;
        if (ih >= 0 && ih < H && iw >= 0 && iw < W) {
          const float f = __ldg(xin + (static_cast<size_t>(ci) * H + ih) * W + iw);
          const float hi = __bfloat162float(__float2bfloat16_rn(f));
          val = col < K ? hi : f - hi;
        }
      }
      v[e] = val;
    }
    o[j] = pack_bf16x2(v[0], v[1]);
  }
  reinterpret_cast<uint4*>(out)[gid] = make_uint4(o[0], o[1], o[2], o[3]);
}

int launch_stem_cols(const dmc_stem_cols_desc& d, cudaStream_t st) {
  const size_t threads = static_cast<size_t>(d.B) * d.H * d.W * 8;
  stem_cols_kernel<<<static_cast<unsigned>((threads + 255) / 256), 256, 0, st>>>(
      d.x, reinterpret_cast<__nv_bfloat16*>(d.out), d.x_batch, d.B, d.Cin, d.H, d.W);
  DMC_CUDA_OK(cudaGetLastError());
  return 0;
}

// =============================================================================================
// Output head, second half: out[n, co, i, j] = bias[co] + sum over the 9 taps of y[n, (i + dh, j + dw), tap * Cout + co]
// (see dmc_head_taps_desc).  One CTA = HT_ROWS image rows of one image: the y rows (+1 halo row above and below) are
// staged in shared memory with a pixel pitch of ypitch + 1 words (lanes = consecutive pixels -> conflict-free), then every
// thread owns one output pixel.  y is read from HBM once (the halo rows come from L2).
// =============================================================================================
constexpr int HT_ROWS = 8;

// POW2: W and ypitch / 4 are powers of two (the CIFAR head: W = 32, ypitch = 32) -- the index arithmetic is shifts and masks;
// the generic form spends more time in integer division than in memory traffic
template <bool POW2>
__global__ void __launch_bounds__(256) head_taps_kernel(const float* __restrict__ y, const float* __restrict__ bias,
                                                        float* __restrict__ out, int H, int W, int Cout, int ypitch, int lw,
                                                        int lv) {
  extern __shared__ float s_y[];
  const int n = blockIdx.y, r0 = blockIdx.x * HT_ROWS;
  const int pitch = ypitch + 1;
  const int rows = min(HT_ROWS, H - r0);
  const int vpp = ypitch / 4;  // float4 per pixel
  const int nvec = (rows + 2) * W * vpp;
  auto split = [&](int e, int& v, int& px, int& rr) {
    if (POW2) {
      v = e & (vpp - 1);
      px = (e >> lv) & (W - 1);
      rr = e >> (lv + lw);
    } else {
      v = e % vpp;
      px = (e / vpp) % W;
      rr = e / (vpp * W);
    }
  };
  // stage rows r0 - 1 .. r0 + rows (zero outside the image): float4 loads (five in flight per thread), scalar stores (odd pitch)
  for (int e0 = threadIdx.x; e0 < nvec; e0 += blockDim.x * 5) {
    float4 t[5];
#pragma unroll
    for (int u = 0; u < 5; ++u) {
      const int e = e0 + u * blockDim.x;
      t[u] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (e < nvec) {
        int v, px, rr;
        split(e, v, px, rr);
        const int gy = r0 - 1 + rr;
        if (gy >= 0 && gy < H) t[u] = __ldg(reinterpret_cast<const float4*>(y + ((static_cast<size_t>(n) * H + gy) * W + px) * ypitch) + v);
      }
    }
#pragma unroll
    for (int u = 0; u < 5; ++u) {
      const int e = e0 + u * blockDim.x;
      if (e < nvec) {
        int v, px, rr;
        split(e, v, px, rr);
        float* d = s_y + (rr * W + px) * pitch + 4 * v;
        d[0] = t[u].x; d[1] = t[u].y; d[2] = t[u].z; d[3] = t[u].w;
      }
    }
  }
  __syncthreads();
  for (int e = threadIdx.x; e < rows * W; e += blockDim.x) {
    const int j = POW2 ? (e & (W - 1)) : e % W, i = POW2 ? (e >> lw) : e / W;  // local row i -> staged row i + 1
    for (int co = 0; co < Cout; ++co) {
      float acc = bias ? __ldg(bias + co) : 0.f;
#pragma unroll
      for (int tap = 0; tap < 9; ++tap) {
        const int dh = tap / 3 - 1, dw = tap % 3 - 1;
        const int jj = j + dw;
        if (jj >= 0 && jj < W) acc += s_y[((i + 1 + dh) * W + jj) * pitch + tap * Cout + co];
      }
      out[((static_cast<size_t>(n) * Cout + co) * H + (r0 + i)) * W + j] = acc;
    }
  }
}

int launch_head_taps(const dmc_head_taps_desc& d, cudaStream_t st) {
  const size_t smem = static_cast<size_t>(HT_ROWS + 2) * d.W * (d.ypitch + 1) * sizeof(float);
  DMC_REQUIRE(smem <= 200 * 1024, "head_taps: image too wide (W=%d)", d.W);
  const int vpp = d.ypitch / 4;
  const bool pow2 = (d.W & (d.W - 1)) == 0 && (vpp & (vpp - 1)) == 0;
  int lw = 0, lv = 0;
  while ((1 << lw) < d.W) ++lw;
  while ((1 << lv) < vpp) ++lv;
  static DeviceOnce attr_set;
  int attr_dev = 0;
  if (smem > 48 * 1024 && attr_set.need(&attr_dev)) {
    DMC_CUDA_OK(cudaFuncSetAttribute(head_taps_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    DMC_CUDA_OK(cudaFuncSetAttribute(head_taps_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    attr_set.done(attr_dev);
  }
  dim3 grid((d.H + HT_ROWS - 1) / HT_ROWS, d.B);
  if (pow2) head_taps_kernel<true><<<grid, 256, smem, st>>>(d.y, d.bias, d.out, d.H, d.W, d.Cout, d.ypitch, lw, lv);
  else head_taps_kernel<false><<<grid, 256, smem, st>>>(d.y, d.bias, d.out, d.H, d.W, d.Cout, d.ypitch, lw, lv);
  DMC_CUDA_OK(cudaGetLastError());
  return 0;
}

int launch_gn_apply(const dmc_gn_apply_desc& d, cudaStream_t st) {
  DMC_REQUIRE(d.nsrc == 1 || d.nsrc == 2, "gn_apply: nsrc=%d", d.nsrc);
  DMC_REQUIRE(d.src[0] && d.stats[0] && d.gamma && d.beta && d.out, "gn_apply: null pointer argument");
  const int C0 = d.src_c[0], C1 = d.nsrc == 2 ? d.src_c[1] : 0;
  const int C = C0 + C1;
  DMC_REQUIRE(C0 % 8 == 0 && C1 % 8 == 0 && d.groups > 0 && d.groups <= 32 && (C / 8) % d.groups == 0 && C / 8 <= 256,
              "gn_apply: channels (%d + %d) must split into %d (<= 32) groups of a multiple of 8, C <= 2048", C0, C1,
              d.groups);
  if (d.nsrc == 2) DMC_REQUIRE(d.src[1] && d.stats[1], "gn_apply: second source missing");
  GnApplyArgs a;
  a.src0 = reinterpret_cast<const uint4*>(d.src[0]);
  a.src1 = reinterpret_cast<const uint4*>(d.nsrc == 2 ? d.src[1] : d.src[0]);
  a.stats0 = d.stats[0];
  a.stats1 = d.nsrc == 2 ? d.stats[1] : d.stats[0];
  a.gamma = d.gamma; a.beta = d.beta; a.out = reinterpret_cast<uint4*>(d.out);
  a.HW = d.HW; a.C0_8 = C0 / 8; a.C1_8 = C1 / 8; a.groups = d.groups; a.eps = d.eps; a.silu = d.silu;
  DMC_REQUIRE(d.drop_p >= 0.f && d.drop_p < 1.f, "gn_apply: drop_p=%f", d.drop_p);
  a.drop_thresh = dropout_threshold(d.drop_p);
  a.drop_scale = d.drop_p > 0.f ? 1.0f / (1.0f - d.drop_p) : 1.0f;
  a.seed = d.seed;
  a.seed_dev = d.seed_dev;
  a.slots0 = d.stats_slots[0];
  a.slots1 = d.nsrc == 2 ? d.stats_slots[1] : 0;
  DMC_REQUIRE(a.slots0 > 0 && (d.nsrc == 1 || a.slots1 > 0), "gn_apply: stats_slots must be positive");
  const bool lo = d.out_lo != nullptr;
  if (lo) DMC_REQUIRE(d.src_lo[0] && (d.nsrc == 1 || d.src_lo[1]), "gn_apply: split-bf16 mode needs the low part of every source");
  a.lo0 = reinterpret_cast<const uint4*>(d.src_lo[0]);
  a.lo1 = reinterpret_cast<const uint4*>(d.nsrc == 2 ? d.src_lo[1] : d.src_lo[0]);
  a.out_lo = reinterpret_cast<uint4*>(d.out_lo);
  // pixels per CTA: every CTA first reduces the partial sums to mean / rstd (a prologue of a few microseconds); more pixels per CTA
  // shrink its share: 4.14 (128 everywhere) -> 3.97 (256) -> 3.93 ms (512) per 2048-image forward, run 34 (DMC_GN_APPLY_SLAB
  // overrides, for A/B runs)
  static const int slab_env = [] {
    const char* e = getenv("DMC_GN_APPLY_SLAB");
    const int v = e ? atoi(e) : 0;
    return (v >= 32 && v <= 4096) ? v : 0;
  }();
  a.slab = slab_env ? slab_env : (d.HW >= 1024 ? 512 : (d.HW >= 256 ? 256 : GN_SLAB));
  dim3 grid((d.HW + a.slab - 1) / a.slab, d.B);
  if (lo) gn_apply_kernel<true><<<grid, 256, 0, st>>>(a);
  else gn_apply_kernel<false><<<grid, 256, 0, st>>>(a);
  DMC_CUDA_OK(cudaGetLastError());
  return 0;
}

// =============================================================================================
// nearest 2x upsample, bf16 NHWC (models/unet.py:119)
// =============================================================================================
__global__ void __launch_bounds__(256) upsample_kernel(const uint4* __restrict__ src, uint4* __restrict__ out, int B, int H,
                                                       int W, int C8) {
  const size_t total = static_cast<size_t>(B) * 4 * H * W * C8;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int cb = static_cast<int>(i % C8);
    size_t pix = i / C8;
    const int ow = static_cast<int>(pix % (2 * W));
    const int oh = static_cast<int>((pix / (2 * W)) % (2 * H));
    const size_t n = pix / (static_cast<size_t>(4) * W * H);
    out[i] = src[((n * H + (oh >> 1)) * W + (ow >> 1)) * C8 + cb];
  }
}

int launch_upsample(const dmc_upsample_desc& d, cudaStream_t st) {
  DMC_REQUIRE(d.src && d.out && d.C % 8 == 0 && d.B > 0, "upsample: bad arguments");
  size_t total = static_cast<size_t>(d.B) * 4 * d.H * d.W * (d.C / 8);
  int blocks = static_cast<int>(std::min<size_t>((total + 255) / 256, static_cast<size_t>(num_sms()) * 16));
  upsample_kernel<<<blocks, 256, 0, st>>>(reinterpret_cast<const uint4*>(d.src), reinterpret_cast<uint4*>(d.out), d.B, d.H,
                                          d.W, d.C / 8);
  DMC_CUDA_OK(cudaGetLastError());
  return 0;
}

}  // namespace dmc

namespace dmc {

// =============================================================================================
// Output head, fused: GroupNorm apply + SiLU + 3x3 convolution C -> Cout <= 8 (+bias) -> fp32 NCHW
// (/root/reference/models/unet.py:237-241, 287-292).  With 3 output channels the layer is memory-bound (read the
// activation once: 256 B per pixel): on the tcgen05 path it needs N padded to 32 and a separate GroupNorm pass
// (0.55 + 0.20 ms per 2048 images); here one CTA normalises a (TH+2) x (W+2) halo tile into shared memory once and
// each warp runs mma.sync m16n8k16 (N = 8 covers the 3 channels) over the 9 taps -- the tensor work is negligible,
// the kernel runs at the speed of its single read of the input.
// =============================================================================================
constexpr int HEAD_TH = 8;        // output rows per CTA (one per warp)
constexpr int HEAD_THREADS = 256;

struct HeadArgs {
  const uint4* src;      // bf16 [B, H, W, C]
  const float* stats;    // [B, slots, C/8, 2]
  const float* gamma;
  const float* beta;
  const float* weight;   // fp32 [Cout, C, 3, 3]
  const float* bias;     // [Cout]
  uint2* wfrag;          // [9 * C/16][32]: the weights as mma.m16n8k16 B fragments (written by head_pack_kernel)
  float* out;            // fp32 [B, Cout, H, W]
  int H, W, C, Cout, groups, slots;
  float eps;
};

// weight fragments of mma.m16n8k16 (B col-major 16 x 8): lane holds k = 2 (lane % 4) + {0, 1} (+8), n = lane / 4
__global__ void __launch_bounds__(256) head_pack_kernel(HeadArgs a) {
  const int CB = a.C / 16;
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= 9 * CB * 32) return;
  const int ks = e >> 5, l = e & 31;
  const int tap = ks / CB, cb = ks % CB;
  const int nn = l >> 2, k0 = cb * 16 + (l & 3) * 2;
  float w[4] = {0.f, 0.f, 0.f, 0.f};
  if (nn < a.Cout) {
    const float* wp = a.weight + (static_cast<size_t>(nn) * a.C) * 9 + tap;
    w[0] = __ldg(wp + static_cast<size_t>(k0) * 9);
    w[1] = __ldg(wp + static_cast<size_t>(k0 + 1) * 9);
    w[2] = __ldg(wp + static_cast<size_t>(k0 + 8) * 9);
    w[3] = __ldg(wp + static_cast<size_t>(k0 + 9) * 9);
  }
  a.wfrag[e] = make_uint2(pack_bf16x2(w[0], w[1]), pack_bf16x2(w[2], w[3]));
}

template <int C, int W>  // compile-time geometry: the index arithmetic below is shifts and constant divisions
__global__ void __launch_bounds__(HEAD_THREADS) head_fused_kernel(HeadArgs a) {
  extern __shared__ __align__(16) uint8_t hsm[];
  constexpr int C8 = C / 8, CB = C / 16;
  constexpr int pitch = C * 2 + 16;                // bytes per halo pixel: +16 keeps ldmatrix rows on distinct banks
  uint8_t* tile = hsm;                             // (TH+2) x (W+2) pixels
  uint2* bfrag = reinterpret_cast<uint2*>(tile + (HEAD_TH + 2) * (W + 2) * pitch);  // [9 * CB][32]
  float* s_sc = reinterpret_cast<float*>(bfrag + 9 * CB * 32);                      // [C] scale / 2, shift / 2
  float* s_sh = s_sc + C;
  __shared__ float s_mean[32], s_rstd[32];
  const int n = blockIdx.y, y0 = blockIdx.x * HEAD_TH;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  // phase 0: group statistics of image n from the partial sums (fixed order -> bit-reproducible); weight fragments
  const int gs8 = C8 / a.groups;
  for (int g = warp; g < a.groups; g += HEAD_THREADS / 32) {
    const float2* base = reinterpret_cast<const float2*>(a.stats) + static_cast<size_t>(n) * a.slots * C8;
    float s = 0.f, ss = 0.f;
    for (int e = lane; e < gs8 * a.slots; e += 32) {
      const float2 v = __ldg(base + static_cast<size_t>(e / gs8) * C8 + g * gs8 + e % gs8);
      s += v.x;
      ss += v.y;
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) {
      s += __shfl_xor_sync(0xFFFFFFFFu, s, o);
      ss += __shfl_xor_sync(0xFFFFFFFFu, ss, o);
    }
    if (lane == 0) {
      const float inv_cnt = 1.0f / (static_cast<float>(gs8 * 8) * static_cast<float>(a.H * a.W));
      const float mean = s * inv_cnt;
      s_mean[g] = mean;
      s_rstd[g] = rsqrtf(fmaxf(ss * inv_cnt - mean * mean, 0.f) + a.eps);
    }
  }
  for (int e = threadIdx.x; e < 9 * CB * 32; e += HEAD_THREADS) bfrag[e] = __ldg(a.wfrag + e);
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += HEAD_THREADS) {
    const int g = (c / 8) / gs8;
    const float s1 = s_rstd[g] * __ldg(a.gamma + c);
    s_sc[c] = 0.5f * s1;                                   // SiLU(y) = h + h tanh(h), h = y / 2
    s_sh[c] = 0.5f * (__ldg(a.beta + c) - s_mean[g] * s1);
  }
  __syncthreads();
  // phase 1: halo tile, normalised + SiLU, bf16; pixels outside the image are the convolution's zero padding.
  // Four independent 128-bit loads in flight per thread.
  constexpr int HP = HEAD_TH + 2, WP = W + 2;
  constexpr int total = HP * WP * C8;
  for (int e0 = threadIdx.x; e0 < total; e0 += 4 * HEAD_THREADS) {
    uint4 v[4];
    bool in[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int e = e0 + u * HEAD_THREADS;
      const int cb = e % C8, px = e / C8;
      const int iy = y0 + px / WP - 1, ix = px % WP - 1;
      in[u] = e < total && iy >= 0 && iy < a.H && ix >= 0 && ix < W;
      if (in[u]) v[u] = __ldg(a.src + ((static_cast<size_t>(n) * a.H + iy) * W + ix) * C8 + cb);
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int e = e0 + u * HEAD_THREADS;
      if (e >= total) continue;
      const int cb = e % C8, px = e / C8;
      uint4 o = make_uint4(0, 0, 0, 0);
      if (in[u]) {
        const uint32_t wv[4] = {v[u].x, v[u].y, v[u].z, v[u].w};
        uint32_t r[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float2 f = unpack_bf16x2(wv[j]);
          const int c = cb * 8 + 2 * j;
          r[j] = pack_bf16x2(silu_from_half(fmaf(f.x, s_sc[c], s_sh[c])), silu_from_half(fmaf(f.y, s_sc[c + 1], s_sh[c + 1])));
        }
        o = make_uint4(r[0], r[1], r[2], r[3]);
      }
      *reinterpret_cast<uint4*>(tile + px * pitch + cb * 16) = o;
    }
  }
  __syncthreads();
  // phase 2: warp = output row, W / 16 m16 tiles per row
  const int y = y0 + warp;
  if (y >= a.H) return;
  const uint32_t tile_s = smem_u32(tile);
  for (int mt = 0; mt < W / 16; ++mt) {
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    for (int tap = 0; tap < 9; ++tap) {
      const int dy = tap / 3, dx = tap % 3;
      const uint32_t rowaddr = tile_s + ((warp + dy) * WP + mt * 16 + (lane & 15) + dx) * pitch + (lane >> 4) * 16;
#pragma unroll 4
      for (int cb = 0; cb < CB; ++cb) {
        uint32_t a0, a1, a2, a3;
        asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];"
                     : "=r"(a0), "=r"(a1), "=r"(a2), "=r"(a3)
                     : "r"(rowaddr + cb * 32));
        const uint2 b = bfrag[(tap * CB + cb) * 32 + lane];
        asm volatile(
            "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, "
            "{%0, %1, %2, %3};"
            : "+f"(acc[0]), "+f"(acc[1]), "+f"(acc[2]), "+f"(acc[3])
            : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b.x), "r"(b.y));
      }
    }
    // D fragment: rows lane / 4 (+8), columns 2 (lane % 4) + {0, 1}
    const int c0 = (lane & 3) * 2;
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      const int x = mt * 16 + (lane >> 2) + half * 8;
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const int c = c0 + j;
        if (c < a.Cout)
          a.out[((static_cast<size_t>(n) * a.Cout + c) * a.H + y) * W + x] = acc[half * 2 + j] + __ldg(a.bias + c);
      }
    }
  }
}

static size_t head_smem(int C, int W) {
  return static_cast<size_t>(HEAD_TH + 2) * (W + 2) * (C * 2 + 16) + static_cast<size_t>(9) * (C / 16) * 32 * 8 +
         static_cast<size_t>(2) * C * 4;
}

// instantiated geometries: the UNet heads of the reference's configs (model_channels 64 / 128, 16..64-pixel rows)
bool head_fused_supported(const dmc_head_desc& d) {
  return (d.C == 64 || d.C == 128) && (d.W == 16 || d.W == 32 || d.W == 64) && d.Cout >= 1 && d.Cout <= 8 && d.groups > 0 &&
         d.groups <= 32 && (d.C / 8) % d.groups == 0 && head_smem(d.C, d.W) <= 200 * 1024;
}

template <int C, int W>
static int launch_head_t(const HeadArgs& a, dim3 grid, size_t smem, cudaStream_t st) {
  static DeviceOnce attr_set;
  int attr_set_dev = 0;
  if (attr_set.need(&attr_set_dev)) {
    DMC_CUDA_OK(cudaFuncSetAttribute(head_fused_kernel<C, W>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    attr_set.done(attr_set_dev);
  }
  head_fused_kernel<C, W><<<grid, HEAD_THREADS, smem, st>>>(a);
  return 0;
}

int launch_head_fused(const dmc_head_desc& d, cudaStream_t st) {
  DMC_REQUIRE(d.src && d.stats && d.gamma && d.beta && d.weight && d.bias && d.out && d.wfrag, "head: null pointer argument");
  DMC_REQUIRE(head_fused_supported(d), "head: unsupported shape (C=%d W=%d Cout=%d groups=%d)", d.C, d.W, d.Cout, d.groups);
  DMC_REQUIRE(d.B > 0 && d.H > 0 && d.stats_slots > 0, "head: empty input");
  HeadArgs a;
  a.src = reinterpret_cast<const uint4*>(d.src);
  a.stats = d.stats; a.gamma = d.gamma; a.beta = d.beta; a.weight = d.weight; a.bias = d.bias; a.out = d.out;
  a.wfrag = reinterpret_cast<uint2*>(d.wfrag);
  a.H = d.H; a.W = d.W; a.C = d.C; a.Cout = d.Cout; a.groups = d.groups; a.slots = d.stats_slots; a.eps = d.eps;
  const size_t smem = head_smem(d.C, d.W);
  const int nfrag = 9 * (d.C / 16) * 32;
  head_pack_kernel<<<(nfrag + 255) / 256, 256, 0, st>>>(a);  // weights -> B fragments (the caller may have updated them)
  dim3 grid((d.H + HEAD_TH - 1) / HEAD_TH, d.B);
  int rc = -1;
#define DMC_HEAD(CC, WW) if (d.C == CC && d.W == WW) rc = launch_head_t<CC, WW>(a, grid, smem, st)
  DMC_HEAD(64, 16); DMC_HEAD(64, 32); DMC_HEAD(64, 64); DMC_HEAD(128, 16); DMC_HEAD(128, 32); DMC_HEAD(128, 64);
#undef DMC_HEAD
  if (rc != 0) return rc;
  DMC_CUDA_OK(cudaGetLastError());
  return 0;
}

}  // namespace dmc
