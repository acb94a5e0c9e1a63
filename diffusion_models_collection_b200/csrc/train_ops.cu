// Backward kernels of the UNet training step that are not GEMM-shaped: GroupNorm + SiLU
// (+ dropout) backward, attention backward, bias / conditioning-row gradients, the sum over 2x2 blocks behind the
// nearest upsample, a strided input-gradient convolution for the three Downsample layers, and layout conversions.
// The GEMM-shaped 2/3 of the backward pass run on the tensor cores: input gradients through conv_umma (weights packed
// transposed and tap-flipped), weight gradients through conv_wgrad.cu.
#include "common.cuh"
#include "kernels.h"

namespace dmc {

// 16-bit keep threshold of the pair hash (common.cuh: dropout_hash2); 0 = no dropout
uint32_t dropout_threshold(float p) {
  if (p <= 0.f) return 0u;
  const double t = static_cast<double>(p) * 65536.0 + 0.5;
  return t >= 65535.0 ? 65535u : std::max(1u, static_cast<uint32_t>(t));
}

// =============================================================================================
// GroupNorm (+SiLU, +dropout) backward over the concatenation of up to two sources
//   forward:  xh = (x - mean) rstd;  z = xh gamma + beta;  a = silu(z);  out = a * keep / (1 - p)
//   backward: dz = dout * keep / (1 - p) * silu'(z);  dgamma = sum dz xh;  dbeta = sum dz;  g = dz gamma
//             dx = rstd (g - mean_grp(g) - xh mean_grp(g xh))
// pass A writes per (image, 128-pixel slab) partial sums, pass B applies; both recompute xh / z from x and the forward
// statistics (nothing but x and the partial sums of the forward pass is kept).
// =============================================================================================
constexpr int GB_SLAB = 64;

struct GnBwdArgs {
  const uint4* src0; const uint4* src1;      // bf16 [B, HW, c_i]
  const float* stats0; const float* stats1;  // forward partial sums
  int slots0, slots1;
  const uint4* dout;                         // bf16 [B, HW, C]
  uint4* dsrc0; uint4* dsrc1;                // bf16 [B, HW, c_i] gradients (written or accumulated)
  int acc0, acc1;
  const float* gamma; const float* beta;
  float* pgb;                                // [B, slabs, C, 2]   partial dgamma / dbeta
  float* ps;                                 // [B, slabs, C/8, 2] partial sum(g), sum(g xh) per 8-channel block
  float* gm;                                 // [B, groups, 2] forward (mean, rstd) of every group
  float* gg;                                 // [B, groups, 2] mean_grp(g), mean_grp(g xh)
  int HW, C0_8, C1_8, groups;
  float eps;
  int silu;
  float drop_scale;                          // 1 / (1 - p), or 1
  uint32_t drop_thresh, seed;
  const uint32_t* seed_dev;                  // optional per-step seed in device memory, added to `seed`
};

__device__ __forceinline__ void gn_group_stats(const GnBwdArgs& a, int n, float* s_mean, float* s_rstd) {
  const int C8 = a.C0_8 + a.C1_8;
  const int gs8 = C8 / a.groups;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int g = warp; g < a.groups; g += 8) {
    const int blo = g * gs8, bhi = blo + gs8;
    float s = 0.f, ss = 0.f;
    {
      const int lo = min(blo, a.C0_8), hi = min(bhi, a.C0_8), nb = hi - lo;
      const float2* base = reinterpret_cast<const float2*>(a.stats0) + static_cast<size_t>(n) * a.slots0 * a.C0_8;
      for (int e = lane; e < nb * a.slots0; e += 32) {
        const float2 v = __ldg(base + static_cast<size_t>(e / nb) * a.C0_8 + lo + e % nb);
        s += v.x; ss += v.y;
      }
    }
    if (a.C1_8 > 0) {
      const int lo = max(blo, a.C0_8) - a.C0_8, hi = max(bhi, a.C0_8) - a.C0_8, nb = hi - lo;
      const float2* base = reinterpret_cast<const float2*>(a.stats1) + static_cast<size_t>(n) * a.slots1 * a.C1_8;
      for (int e = lane; e < nb * a.slots1; e += 32) {
        const float2 v = __ldg(base + static_cast<size_t>(e / nb) * a.C1_8 + lo + e % nb);
        s += v.x; ss += v.y;
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
      s_mean[g] = mean;
      s_rstd[g] = rsqrtf(fmaxf(ss * inv_cnt - mean * mean, 0.f) + a.eps);
    }
  }
}

// dz of the 8 channels of one pixel-block from x, dout
template <bool SILU, bool DROP>
__device__ __forceinline__ void gn_dz8(const GnBwdArgs& a, const uint4& xv, const uint4& dv, float mean, float rstd,
                                       const float (&gam)[8], const float (&bet)[8], uint64_t idx0, uint32_t seed, float (&xh)[8],
                                       float (&dz)[8]) {
  const uint32_t xw[4] = {xv.x, xv.y, xv.z, xv.w};
  const uint32_t dw[4] = {dv.x, dv.y, dv.z, dv.w};
  const float nmr = -mean * rstd;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float2 xf = unpack_bf16x2(xw[j]);
    const float2 df = unpack_bf16x2(dw[j]);
    const float xs[2] = {xf.x, xf.y};
    float ds[2] = {df.x, df.y};
    if (DROP) {
      const uint32_t h = dropout_hash2(seed, idx0 + 2 * j);
      ds[0] = (h & 0xFFFFu) >= a.drop_thresh ? ds[0] * a.drop_scale : 0.f;
      ds[1] = (h >> 16) >= a.drop_thresh ? ds[1] * a.drop_scale : 0.f;
    }
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      const int c = 2 * j + k;
      xh[c] = fmaf(xs[k], rstd, nmr);
      float d = ds[k];
      if (SILU) {
        const float z = fmaf(xh[c], gam[c], bet[c]);
        float th;  // sigmoid(z) = 0.5 + 0.5 tanh(z / 2): one MUFU op (the forward pass uses the same form)
        asm("tanh.approx.f32 %0, %1;" : "=f"(th) : "f"(0.5f * z));
        const float sg = fmaf(0.5f, th, 0.5f);
        d *= sg * fmaf(z, 1.0f - sg, 1.0f);
      }
      dz[c] = d;
    }
  }
}

// per (image, group) quantities computed ONCE by two tiny kernels (one CTA per image) instead of by every CTA of the two passes
__global__ void __launch_bounds__(256) gn_group_moments_kernel(GnBwdArgs a) {
  __shared__ float s_mean[32], s_rstd[32];
  const int n = blockIdx.x;
  gn_group_stats(a, n, s_mean, s_rstd);
  __syncthreads();
  if (threadIdx.x < a.groups)
    reinterpret_cast<float2*>(a.gm)[n * a.groups + threadIdx.x] = make_float2(s_mean[threadIdx.x], s_rstd[threadIdx.x]);
}

__global__ void __launch_bounds__(256) gn_group_grads_kernel(GnBwdArgs a, int slabs) {
  // group sums of g and g xh over all slabs (fixed order)
  const int C8 = a.C0_8 + a.C1_8;
  const int n = blockIdx.x;
  const int gs8 = C8 / a.groups;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int g = warp; g < a.groups; g += 8) {
    float s1 = 0.f, s2 = 0.f;
    const float2* base = reinterpret_cast<const float2*>(a.ps) + static_cast<size_t>(n) * slabs * C8;
    for (int e = lane; e < gs8 * slabs; e += 32) {
      const float2 v = base[static_cast<size_t>(e / gs8) * C8 + g * gs8 + e % gs8];
      s1 += v.x; s2 += v.y;
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) {
      s1 += __shfl_xor_sync(0xFFFFFFFFu, s1, o);
      s2 += __shfl_xor_sync(0xFFFFFFFFu, s2, o);
    }
    if (lane == 0) {
      const float inv_m = 1.0f / (static_cast<float>(gs8 * 8) * static_cast<float>(a.HW));
      reinterpret_cast<float2*>(a.gg)[n * a.groups + g] = make_float2(s1 * inv_m, s2 * inv_m);
    }
  }
}

template <bool APPLY, bool SILU, bool DROP>
__global__ void __launch_bounds__(256, 4) gn_bwd_kernel(GnBwdArgs a) {
  __shared__ float s_mean[32], s_rstd[32], s_g1[32], s_g2[32];
  extern __shared__ float red[];  // pass A: [rows][C8][18]
  const int C8 = a.C0_8 + a.C1_8;
  const int n = blockIdx.y;
  const int gs8 = C8 / a.groups;
  if (threadIdx.x < a.groups) {
    const float2 m = __ldg(reinterpret_cast<const float2*>(a.gm) + n * a.groups + threadIdx.x);
    s_mean[threadIdx.x] = m.x;
    s_rstd[threadIdx.x] = m.y;
    if (APPLY) {
      const float2 gq = __ldg(reinterpret_cast<const float2*>(a.gg) + n * a.groups + threadIdx.x);
      s_g1[threadIdx.x] = gq.x;
      s_g2[threadIdx.x] = gq.y;
    }
  }
  __syncthreads();
  const int cb = threadIdx.x % C8, r0 = threadIdx.x / C8;
  const int rpi = blockDim.x / C8;
  const bool active = r0 < rpi;
  const int g = cb / gs8;
  const float mean = s_mean[g], rstd = s_rstd[g];
  float gam[8], bet[8];
  {
    const float4* g4 = reinterpret_cast<const float4*>(a.gamma + cb * 8);
    const float4* b4 = reinterpret_cast<const float4*>(a.beta + cb * 8);
    const float4 ga = __ldg(g4), gb = __ldg(g4 + 1), ba = __ldg(b4), bb = __ldg(b4 + 1);
    gam[0] = ga.x; gam[1] = ga.y; gam[2] = ga.z; gam[3] = ga.w; gam[4] = gb.x; gam[5] = gb.y; gam[6] = gb.z; gam[7] = gb.w;
    bet[0] = ba.x; bet[1] = ba.y; bet[2] = ba.z; bet[3] = ba.w; bet[4] = bb.x; bet[5] = bb.y; bet[6] = bb.z; bet[7] = bb.w;
  }
  const bool first = cb < a.C0_8;
  const uint4* src = first ? a.src0 + static_cast<size_t>(n) * a.HW * a.C0_8 + cb
                           : a.src1 + static_cast<size_t>(n) * a.HW * a.C1_8 + (cb - a.C0_8);
  uint4* dsrc = first ? a.dsrc0 + static_cast<size_t>(n) * a.HW * a.C0_8 + cb
                      : a.dsrc1 + static_cast<size_t>(n) * a.HW * a.C1_8 + (cb - a.C0_8);
  const int sstride = first ? a.C0_8 : a.C1_8;
  const int acc = first ? a.acc0 : a.acc1;
  const uint4* dout = a.dout + static_cast<size_t>(n) * a.HW * C8 + cb;
  const int p0 = blockIdx.x * GB_SLAB;
  const int p1 = min(p0 + GB_SLAB, a.HW);
  const uint32_t seed = DROP ? a.seed + (a.seed_dev ? __ldg(a.seed_dev) : 0u) : 0u;
  float dgam[8], dbet[8], s1 = 0.f, s2 = 0.f;
#pragma unroll
  for (int j = 0; j < 8; ++j) dgam[j] = dbet[j] = 0.f;
  if (active) {
    uint4 xn = make_uint4(0, 0, 0, 0), dn = xn, on = xn;
    if (p0 + r0 < p1) {  // software pipeline: the next pixel's loads are in flight while this one is processed
      xn = __ldg(src + static_cast<size_t>(p0 + r0) * sstride);
      dn = __ldg(dout + static_cast<size_t>(p0 + r0) * C8);
      if (APPLY && acc) on = *(dsrc + static_cast<size_t>(p0 + r0) * sstride);
    }
    for (int p = p0 + r0; p < p1; p += rpi) {
      const uint4 xv = xn, dv = dn, old = on;
      if (p + rpi < p1) {
        xn = __ldg(src + static_cast<size_t>(p + rpi) * sstride);
        dn = __ldg(dout + static_cast<size_t>(p + rpi) * C8);
        if (APPLY && acc) on = *(dsrc + static_cast<size_t>(p + rpi) * sstride);
      }
      float xh[8], dz[8];
      gn_dz8<SILU, DROP>(a, xv, dv, mean, rstd, gam, bet, ((static_cast<uint64_t>(n) * a.HW + p) * C8 + cb) * 8, seed, xh, dz);
      if (!APPLY) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          dgam[j] = fmaf(dz[j], xh[j], dgam[j]);
          dbet[j] += dz[j];
          const float gg = dz[j] * gam[j];
          s1 += gg;
          s2 = fmaf(gg, xh[j], s2);
        }
      } else {
        const float m1 = s_g1[g], m2 = s_g2[g];
        float dx[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) dx[j] = rstd * (dz[j] * gam[j] - m1 - xh[j] * m2);
        uint4* o = dsrc + static_cast<size_t>(p) * sstride;
        if (acc) {
          const uint32_t ow[4] = {old.x, old.y, old.z, old.w};
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float2 f = unpack_bf16x2(ow[j]);
            dx[2 * j] += f.x;
            dx[2 * j + 1] += f.y;
          }
        }
        *o = make_uint4(pack_bf16x2(dx[0], dx[1]), pack_bf16x2(dx[2], dx[3]), pack_bf16x2(dx[4], dx[5]),
                        pack_bf16x2(dx[6], dx[7]));
      }
    }
  }
  if (!APPLY) {
    // combine the rpi threads that share a channel block (fixed order), one plain store per (image, slab, block)
    float* mine = red + (static_cast<size_t>(r0) * C8 + cb) * 18;
    if (active) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        mine[j] = dgam[j];
        mine[8 + j] = dbet[j];
      }
      mine[16] = s1;
      mine[17] = s2;
    }
    __syncthreads();
    if (r0 == 0) {
      for (int k = 1; k < rpi; ++k) {
        const float* o = red + (static_cast<size_t>(k) * C8 + cb) * 18;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          dgam[j] += o[j];
          dbet[j] += o[8 + j];
        }
        s1 += o[16];
        s2 += o[17];
      }
      const size_t slab = static_cast<size_t>(n) * gridDim.x + blockIdx.x;
      float2* pg = reinterpret_cast<float2*>(a.pgb) + (slab * C8 + cb) * 8;
#pragma unroll
      for (int j = 0; j < 8; ++j) pg[j] = make_float2(dgam[j], dbet[j]);
      reinterpret_cast<float2*>(a.ps)[slab * C8 + cb] = make_float2(s1, s2);
    }
  }
}

// dgamma[c], dbeta[c] = sum over (image, slab) of the partials: one CTA per 8 channels, 128 row lanes each adding its rows in
// index order, then the 128 lane sums in lane order (fixed association: bit-reproducible).  Many short chains instead of a few
// long ones: the kernel is pure latency (a few MB).
__global__ void __launch_bounds__(1024) gn_param_reduce_kernel(const float2* __restrict__ pgb, float* __restrict__ dgamma,
                                                               float* __restrict__ dbeta, int rows, int C) {
  __shared__ float2 red[128][8];
  const int cl = threadIdx.x & 7, rl = threadIdx.x >> 3;
  const int c = blockIdx.x * 8 + cl;
  float g = 0.f, b = 0.f;
  if (c < C) {
    int r = rl;
    for (; r + 384 < rows; r += 512) {
      float2 v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) v[u] = pgb[static_cast<size_t>(r + 128 * u) * C + c];
#pragma unroll
      for (int u = 0; u < 4; ++u) { g += v[u].x; b += v[u].y; }
    }
    for (; r < rows; r += 128) {
      const float2 v = pgb[static_cast<size_t>(r) * C + c];
      g += v.x; b += v.y;
    }
  }
  red[rl][cl] = make_float2(g, b);
  __syncthreads();
  if (threadIdx.x < 32) {  // 4 lanes per channel add 32 lane sums each (in order), then lane order across the 4
    const int ch = threadIdx.x & 7, part = threadIdx.x >> 3;
    float gs = 0.f, bs = 0.f;
#pragma unroll 8
    for (int k = 0; k < 32; ++k) { gs += red[part * 32 + k][ch].x; bs += red[part * 32 + k][ch].y; }
    const float g1 = __shfl_sync(0xFFFFFFFFu, gs, ch + 8), g2 = __shfl_sync(0xFFFFFFFFu, gs, ch + 16), g3 = __shfl_sync(0xFFFFFFFFu, gs, ch + 24);
    const float b1 = __shfl_sync(0xFFFFFFFFu, bs, ch + 8), b2 = __shfl_sync(0xFFFFFFFFu, bs, ch + 16), b3 = __shfl_sync(0xFFFFFFFFu, bs, ch + 24);
    if (part == 0 && blockIdx.x * 8 + ch < C) {
      dgamma[blockIdx.x * 8 + ch] = ((gs + g1) + g2) + g3;
      dbeta[blockIdx.x * 8 + ch] = ((bs + b1) + b2) + b3;
    }
  }
}

long long gn_backward_scratch_floats(const dmc_gn_bwd_desc& d) {
  const long long C = d.src_c[0] + (d.nsrc == 2 ? d.src_c[1] : 0);
  const long long slabs = (d.HW + GB_SLAB - 1) / GB_SLAB;
  return static_cast<long long>(d.B) * slabs * (2 * C + C / 4) + 4LL * d.B * d.groups;
}

int launch_gn_backward(const dmc_gn_bwd_desc& d, cudaStream_t st) {
  DMC_REQUIRE(d.nsrc == 1 || d.nsrc == 2, "gn_backward: nsrc=%d", d.nsrc);
  DMC_REQUIRE(d.src[0] && d.stats[0] && d.dout && d.dsrc[0] && d.gamma && d.beta && d.dgamma && d.dbeta && d.scratch,
              "gn_backward: null pointer argument");
  const int C0 = d.src_c[0], C1 = d.nsrc == 2 ? d.src_c[1] : 0, C = C0 + C1;
  DMC_REQUIRE(C0 % 8 == 0 && C1 % 8 == 0 && d.groups > 0 && d.groups <= 32 && (C / 8) % d.groups == 0 && C / 8 <= 256,
              "gn_backward: unsupported channels (%d + %d, %d groups)", C0, C1, d.groups);
  if (d.nsrc == 2) DMC_REQUIRE(d.src[1] && d.stats[1] && d.dsrc[1], "gn_backward: second source missing");
  GnBwdArgs a;
  a.src0 = reinterpret_cast<const uint4*>(d.src[0]);
  a.src1 = reinterpret_cast<const uint4*>(d.nsrc == 2 ? d.src[1] : d.src[0]);
  a.stats0 = d.stats[0]; a.stats1 = d.nsrc == 2 ? d.stats[1] : d.stats[0];
  a.slots0 = d.stats_slots[0]; a.slots1 = d.nsrc == 2 ? d.stats_slots[1] : 0;
  a.dout = reinterpret_cast<const uint4*>(d.dout);
  a.dsrc0 = reinterpret_cast<uint4*>(d.dsrc[0]);
  a.dsrc1 = reinterpret_cast<uint4*>(d.nsrc == 2 ? d.dsrc[1] : d.dsrc[0]);
  a.acc0 = d.accumulate[0]; a.acc1 = d.nsrc == 2 ? d.accumulate[1] : 0;
  a.gamma = d.gamma; a.beta = d.beta;
  const int slabs = (d.HW + GB_SLAB - 1) / GB_SLAB;
  a.pgb = d.scratch;
  a.ps = d.scratch + static_cast<size_t>(d.B) * slabs * C * 2;
  a.gm = a.ps + static_cast<size_t>(d.B) * slabs * (C / 8) * 2;
  a.gg = a.gm + static_cast<size_t>(d.B) * d.groups * 2;
  a.HW = d.HW; a.C0_8 = C0 / 8; a.C1_8 = C1 / 8; a.groups = d.groups; a.eps = d.eps; a.silu = d.silu;
  a.drop_thresh = dropout_threshold(d.drop_p);
  a.drop_scale = d.drop_p > 0.f ? 1.0f / (1.0f - d.drop_p) : 1.0f;
  a.seed = d.seed;
  a.seed_dev = d.seed_dev;
  const int C8 = C / 8, rows = std::max(1, 256 / C8);
  dim3 grid(slabs, d.B);
  const size_t smem = static_cast<size_t>(rows) * C8 * 18 * sizeof(float);
  DMC_REQUIRE(smem <= 48 * 1024, "gn_backward: %zu bytes of shared memory", smem);
  const bool silu = d.silu != 0, drop = a.drop_thresh != 0u;
  gn_group_moments_kernel<<<d.B, 256, 0, st>>>(a);
#define DMC_GN_BWD(APPLY, SM)                                                                             \
  do {                                                                                                    \
    if (silu && drop) gn_bwd_kernel<APPLY, true, true><<<grid, 256, SM, st>>>(a);                         \
    else if (silu) gn_bwd_kernel<APPLY, true, false><<<grid, 256, SM, st>>>(a);                           \
    else if (drop) gn_bwd_kernel<APPLY, false, true><<<grid, 256, SM, st>>>(a);                           \
    else gn_bwd_kernel<APPLY, false, false><<<grid, 256, SM, st>>>(a);                                    \
  } while (0)
  DMC_GN_BWD(false, smem);
  gn_group_grads_kernel<<<d.B, 256, 0, st>>>(a, slabs);
  DMC_GN_BWD(true, 0);
#undef DMC_GN_BWD
  gn_param_reduce_kernel<<<(C + 7) / 8, 1024, 0, st>>>(reinterpret_cast<const float2*>(a.pgb), d.dgamma, d.dbeta,
                                                          d.B * slabs, C);
  DMC_CUDA_OK(cudaGetLastError());
  return 0;
}

// =============================================================================================
// Attention backward (CUDA cores, fp32 math, L <= 256, head dim 64): dqkv from qkv, the forward output o and do.
//   P = softmax(Q K^T s);  D_i = do_i . o_i;  dV = P^T dO;  dS = P (dO V^T - D);  dQ = s dS K;  dK = s dS^T Q
// One CTA per (image, head): pass 1 thread = query row (row max / sum, D, dQ), pass 2 thread = key row (dK, dV).
// =============================================================================================
template <int HD>
__global__ void __launch_bounds__(256) attention_bwd_kernel(const __nv_bfloat16* __restrict__ qkv,
                                                            const __nv_bfloat16* __restrict__ o,
                                                            const __nv_bfloat16* __restrict__ dout,
                                                            __nv_bfloat16* __restrict__ dqkv, int L, int C, float scale) {
  extern __shared__ __nv_bfloat16 sm[];  // Q, K, V, dO: [L][HD] each; then fp32 m[L], l[L], D[L]
  __nv_bfloat16* sQ = sm;
  __nv_bfloat16* sK = sQ + L * HD;
  __nv_bfloat16* sV = sK + L * HD;
  __nv_bfloat16* sdO = sV + L * HD;
  float* s_m = reinterpret_cast<float*>(sdO + L * HD);
  float* s_l = s_m + L;
  float* s_D = s_l + L;
  const int n = blockIdx.y, h = blockIdx.x;
  const size_t rs = static_cast<size_t>(3) * C;
  const __nv_bfloat16* base = qkv + static_cast<size_t>(n) * L * rs + h * HD;
  for (int e = threadIdx.x; e < L * (HD / 8); e += blockDim.x) {
    const int r = e / (HD / 8), v = e % (HD / 8);
    const uint4* rp = reinterpret_cast<const uint4*>(base + static_cast<size_t>(r) * rs);
    reinterpret_cast<uint4*>(sQ + r * HD)[v] = __ldg(rp + v);
    reinterpret_cast<uint4*>(sK + r * HD)[v] = __ldg(reinterpret_cast<const uint4*>(base + static_cast<size_t>(r) * rs + C) + v);
    reinterpret_cast<uint4*>(sV + r * HD)[v] = __ldg(reinterpret_cast<const uint4*>(base + static_cast<size_t>(r) * rs + 2 * C) + v);
    reinterpret_cast<uint4*>(sdO + r * HD)[v] =
        __ldg(reinterpret_cast<const uint4*>(dout + (static_cast<size_t>(n) * L + r) * C + h * HD) + v);
  }
  __syncthreads();
  // ---- pass 1: thread i = query row
  for (int i = threadIdx.x; i < L; i += blockDim.x) {
    float q[HD], d_o[HD], dq[HD];
    float D = 0.f;
    const __nv_bfloat16* op = o + (static_cast<size_t>(n) * L + i) * C + h * HD;
#pragma unroll
    for (int d = 0; d < HD; ++d) {
      q[d] = __bfloat162float(sQ[i * HD + d]) * scale;
      d_o[d] = __bfloat162float(sdO[i * HD + d]);
      D = fmaf(d_o[d], __bfloat162float(op[d]), D);
      dq[d] = 0.f;
    }
    float m = -INFINITY;
    for (int j = 0; j < L; ++j) {
      float s = 0.f;
#pragma unroll
      for (int d = 0; d < HD; ++d) s = fmaf(q[d], __bfloat162float(sK[j * HD + d]), s);
      m = fmaxf(m, s);
    }
    float l = 0.f;
    for (int j = 0; j < L; ++j) {
      float s = 0.f;
#pragma unroll
      for (int d = 0; d < HD; ++d) s = fmaf(q[d], __bfloat162float(sK[j * HD + d]), s);
      l += __expf(s - m);
    }
    const float inv_l = 1.0f / l;
    for (int j = 0; j < L; ++j) {
      float s = 0.f, dp = 0.f;
#pragma unroll
      for (int d = 0; d < HD; ++d) {
        s = fmaf(q[d], __bfloat162float(sK[j * HD + d]), s);
        dp = fmaf(d_o[d], __bfloat162float(sV[j * HD + d]), dp);
      }
      const float ds = __expf(s - m) * inv_l * (dp - D) * scale;
#pragma unroll
      for (int d = 0; d < HD; ++d) dq[d] = fmaf(ds, __bfloat162float(sK[j * HD + d]), dq[d]);
    }
    s_m[i] = m; s_l[i] = inv_l; s_D[i] = D;
    __nv_bfloat16* dqp = dqkv + (static_cast<size_t>(n) * L + i) * rs + h * HD;
#pragma unroll
    for (int d = 0; d < HD; ++d) dqp[d] = __float2bfloat16(dq[d]);
  }
  __syncthreads();
  // ---- pass 2: thread j = key row
  for (int j = threadIdx.x; j < L; j += blockDim.x) {
    float k[HD], v[HD], dk[HD], dv[HD];
#pragma unroll
    for (int d = 0; d < HD; ++d) {
      k[d] = __bfloat162float(sK[j * HD + d]);
      v[d] = __bfloat162float(sV[j * HD + d]);
      dk[d] = dv[d] = 0.f;
    }
    for (int i = 0; i < L; ++i) {
      float s = 0.f, dp = 0.f;
#pragma unroll
      for (int d = 0; d < HD; ++d) {
        s = fmaf(__bfloat162float(sQ[i * HD + d]), k[d], s);
        dp = fmaf(__bfloat162float(sdO[i * HD + d]), v[d], dp);
      }
      const float pij = __expf(s * scale - s_m[i]) * s_l[i];
      const float ds = pij * (dp - s_D[i]) * scale;
#pragma unroll
      for (int d = 0; d < HD; ++d) {
        dv[d] = fmaf(pij, __bfloat162float(sdO[i * HD + d]), dv[d]);
        dk[d] = fmaf(ds, __bfloat162float(sQ[i * HD + d]), dk[d]);
      }
    }
    __nv_bfloat16* dkp = dqkv + (static_cast<size_t>(n) * L + j) * rs + C + h * HD;
    __nv_bfloat16* dvp = dkp + C;
#pragma unroll
    for (int d = 0; d < HD; ++d) {
      dkp[d] = __float2bfloat16(dk[d]);
      dvp[d] = __float2bfloat16(dv[d]);
    }
  }
}

int launch_attention_backward(const dmc_attn_bwd_desc& d, cudaStream_t st) {
  DMC_REQUIRE(d.qkv && d.out && d.dout && d.dqkv && d.B > 0 && d.heads > 0, "attention_backward: bad arguments");
  const int hd = d.C / d.heads;
  DMC_REQUIRE(d.C == d.heads * hd && hd == 64 && d.L >= 1 && d.L <= 256, "attention_backward: head dim 64 and L <= 256 (C=%d heads=%d L=%d)",
              d.C, d.heads, d.L);
  {
    const char* e = getenv("DMC_ATTN_BWD_IMPL");  // 1: the CUDA-core fp32 kernel below (debug / A-B)
    if (!(e && e[0] == '1')) return launch_attention_backward_mma(d, st);
  }
  const size_t smem = static_cast<size_t>(4) * d.L * hd * 2 + static_cast<size_t>(3) * d.L * 4;
  static DeviceOnce attr;
  int attr_dev = 0;
  if (attr.need(&attr_dev)) {
    DMC_CUDA_OK(cudaFuncSetAttribute(attention_bwd_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
    attr.done(attr_dev);
  }
  const int threads = d.L >= 256 ? 256 : (d.L >= 128 ? 128 : 64);
  attention_bwd_kernel<64><<<dim3(d.heads, d.B), threads, smem, st>>>(
      reinterpret_cast<const __nv_bfloat16*>(d.qkv), reinterpret_cast<const __nv_bfloat16*>(d.out),
      reinterpret_cast<const __nv_bfloat16*>(d.dout), reinterpret_cast<__nv_bfloat16*>(d.dqkv), d.L, d.C,
      1.0f / sqrtf(static_cast<float>(hd)));
  DMC_CUDA_OK(cudaGetLastError());
  return 0;
}

// =============================================================================================
// Small reductions and layout kernels
// =============================================================================================
// out[n][c] (+)= sum over the pixels of src[n, p, c]  (per-image conditioning-row gradients); the bias gradient adds the
// images in index order in a second small kernel (deterministic).  One CTA = one image x 128 channels; 16 pixel lanes x
// 16 channel vectors of 16 bytes.
__global__ void __launch_bounds__(256) channel_sum_image_kernel(const uint4* __restrict__ src, float* __restrict__ out, int HW,
                                                                int C, int accumulate) {
  __shared__ float red[16][129];
  const int C8 = C >> 3;
  const int n = blockIdx.y, tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int v8 = blockIdx.x * 16 + tx;
  float acc[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[j] = 0.f;
  if (v8 < C8) {
    const uint4* base = src + static_cast<size_t>(n) * HW * C8 + v8;
    int p = ty;
    for (; p + 48 < HW; p += 64) {  // four independent loads in flight per thread
      uint4 v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) v[u] = __ldg(base + static_cast<size_t>(p + 16 * u) * C8);
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const uint32_t w[4] = {v[u].x, v[u].y, v[u].z, v[u].w};
#pragma unroll
        for (int j = 0; j < 4; ++j) { const float2 f = unpack_bf16x2(w[j]); acc[2 * j] += f.x; acc[2 * j + 1] += f.y; }
      }
    }
    for (; p < HW; p += 16) {
      const uint4 v = __ldg(base + static_cast<size_t>(p) * C8);
      const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) { const float2 f = unpack_bf16x2(w[j]); acc[2 * j] += f.x; acc[2 * j + 1] += f.y; }
    }
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) red[ty][tx * 8 + j] = acc[j];
  __syncthreads();
  if (threadIdx.x < 128) {
    const int c = blockIdx.x * 128 + threadIdx.x;
    if (c < C) {
      float s2 = 0.f;
#pragma unroll
      for (int r = 0; r < 16; ++r) s2 += red[r][threadIdx.x];
      float* o = out + static_cast<size_t>(n) * C + c;
      *o = accumulate ? *o + s2 : s2;
    }
  }
}

__global__ void __launch_bounds__(1024) channel_sum_total_kernel(const float* __restrict__ per_image, float* __restrict__ out, int B,
                                                                 int C, int accumulate) {
  __shared__ float red[32][33];
  const int cl = threadIdx.x & 31, nl = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + cl;
  float s2 = 0.f;
  if (c < C)
    for (int n = nl; n < B; n += 32) s2 += per_image[static_cast<size_t>(n) * C + c];
  red[nl][cl] = s2;
  __syncthreads();
  if (nl == 0 && c < C) {
    s2 = 0.f;
#pragma unroll
    for (int k = 0; k < 32; ++k) s2 += red[k][cl];
    out[c] = accumulate ? out[c] + s2 : s2;
  }
}

int launch_channel_sum(const void* src, float* out, int B, int HW, int C, int per_image, int accumulate, float* scratch,
                       cudaStream_t st) {
  DMC_REQUIRE(src && out && B > 0 && HW > 0 && C > 0 && C % 8 == 0, "channel_sum: bad arguments");
  DMC_REQUIRE(per_image || scratch, "channel_sum: the all-image sum needs a [B, C] fp32 scratch buffer");
  dim3 grid((C + 127) / 128, B);
  channel_sum_image_kernel<<<grid, 256, 0, st>>>(reinterpret_cast<const uint4*>(src), per_image ? out : scratch, HW, C,
                                                 per_image ? accumulate : 0);
  DMC_CUDA_OK(cudaGetLastError());
  if (!per_image) {
    channel_sum_total_kernel<<<(C + 31) / 32, 1024, 0, st>>>(scratch, out, B, C, accumulate);
    DMC_CUDA_OK(cudaGetLastError());
  }
  return 0;
}

// dst[n, 2i, 2j, :] = src[n, i, j, :], zeros elsewhere: the gradient of a stride-2 convolution's output spread onto the
// input grid, so that its input gradient is an ordinary stride-1 3x3 convolution (tensor-core kernel) with flipped weights
__global__ void __launch_bounds__(256) dilate2x_kernel(const uint4* __restrict__ src, uint4* __restrict__ dst, int B, int h, int w,
                                                       int C8) {
  const size_t total = static_cast<size_t>(B) * 4 * h * w * C8;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int cb = static_cast<int>(i % C8);
    size_t pix = i / C8;
    const int x = static_cast<int>(pix % (2 * w)), y = static_cast<int>((pix / (2 * w)) % (2 * h));
    const size_t n = pix / (static_cast<size_t>(4) * w * h);
    uint4 v = make_uint4(0, 0, 0, 0);
    if (((x | y) & 1) == 0) v = __ldg(src + ((n * h + (y >> 1)) * w + (x >> 1)) * C8 + cb);
    dst[i] = v;
  }
}

int launch_dilate2x(const void* src, void* dst, int B, int h, int w, int C, cudaStream_t st) {
  DMC_REQUIRE(src && dst && B > 0 && h > 0 && w > 0 && C % 8 == 0, "dilate2x: bad arguments");
  const size_t total = static_cast<size_t>(B) * 4 * h * w * (C / 8);
  const int blocks = static_cast<int>(std::min<size_t>((total + 255) / 256, static_cast<size_t>(num_sms()) * 16));
  dilate2x_kernel<<<blocks, 256, 0, st>>>(reinterpret_cast<const uint4*>(src), reinterpret_cast<uint4*>(dst), B, h, w, C / 8);
  DMC_CUDA_OK(cudaGetLastError());
  return 0;
}

// dlow[n, i, j, c] (+)= sum of the 2x2 block of dhigh (backward of the nearest 2x upsample, models/unet.py:119)
__global__ void __launch_bounds__(256) block_sum2x2_kernel(const uint4* __restrict__ dhigh, uint4* __restrict__ dlow, int B,
                                                           int H, int W, int C8, int accumulate) {
  const size_t total = static_cast<size_t>(B) * H * W * C8;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int cb = static_cast<int>(i % C8);
    size_t pix = i / C8;
    const int w = static_cast<int>(pix % W), h = static_cast<int>((pix / W) % H);
    const size_t n = pix / (static_cast<size_t>(W) * H);
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = 0.f;
    if (accumulate) {
      const uint4 v = dlow[i];
      const uint32_t u[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) { const float2 f = unpack_bf16x2(u[j]); acc[2 * j] = f.x; acc[2 * j + 1] = f.y; }
    }
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
      for (int b = 0; b < 2; ++b) {
        const uint4 v = dhigh[((n * 2 * H + 2 * h + a) * 2 * W + 2 * w + b) * C8 + cb];
        const uint32_t u[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) { const float2 f = unpack_bf16x2(u[j]); acc[2 * j] += f.x; acc[2 * j + 1] += f.y; }
      }
    dlow[i] = make_uint4(pack_bf16x2(acc[0], acc[1]), pack_bf16x2(acc[2], acc[3]), pack_bf16x2(acc[4], acc[5]),
                         pack_bf16x2(acc[6], acc[7]));
  }
}

int launch_block_sum2x2(const void* dhigh, void* dlow, int B, int H, int W, int C, int accumulate, cudaStream_t st) {
  DMC_REQUIRE(dhigh && dlow && C % 8 == 0 && B > 0, "block_sum2x2: bad arguments");
  const size_t total = static_cast<size_t>(B) * H * W * (C / 8);
  const int blocks = static_cast<int>(std::min<size_t>((total + 255) / 256, static_cast<size_t>(num_sms()) * 16));
  block_sum2x2_kernel<<<blocks, 256, 0, st>>>(reinterpret_cast<const uint4*>(dhigh), reinterpret_cast<uint4*>(dlow), B, H, W,
                                              C / 8, accumulate);
  DMC_CUDA_OK(cudaGetLastError());
  return 0;
}

// Re-pack of the convolution weights after an optimizer step: fp32 [Cout, Cin, kh, kw] parameters -> the bf16 GEMM operands of
// the forward convolution (mode 0: [Cout][tap][ci], a column range of a possibly wider K-concatenated matrix) and of the
// input-gradient convolution (mode 1: [ci][flipped tap][co], zero-padded co).  ONE launch for all layers: blockIdx.y = item,
// 32 x 32 (co, ci) tiles staged through shared memory so that both the fp32 reads and the bf16 writes are contiguous runs.
__global__ void __launch_bounds__(256) pack_weights_kernel(const dmc_pack_item* __restrict__ items) {
  __shared__ float tile[32][32 * 9 + 1];
  const dmc_pack_item it = items[blockIdx.y];
  __nv_bfloat16* dst = reinterpret_cast<__nv_bfloat16*>(it.dst);
  const int tci = (it.cin + 31) / 32, ntile = ((it.cout + 31) / 32) * tci;
  for (int tl = blockIdx.x; tl < ntile; tl += gridDim.x) {
    const int co0 = (tl / tci) * 32, cc0 = (tl % tci) * 32;
    const int nci = min(32, it.cin - cc0), ncol = nci * it.taps;
    for (int e = threadIdx.x; e < 32 * ncol; e += 256) {
      const int r = e / ncol, c = e - r * ncol, co = co0 + r;
      tile[r][c] = co < it.cout ? __ldg(it.src + (static_cast<size_t>(co) * it.cin_total + it.ci0 + cc0) * it.taps + c) : 0.f;
    }
    __syncthreads();
    if (it.mode == 0) {
      for (int e = threadIdx.x; e < 32 * ncol; e += 256) {
        const int ci = e % nci, tap = (e / nci) % it.taps, r = e / ncol, co = co0 + r;
        if (co < it.cout)
          dst[static_cast<size_t>(co) * it.ld + it.col0 + tap * it.cin + cc0 + ci] = __float2bfloat16_rn(tile[r][ci * it.taps + tap]);
      }
    } else {
      for (int e = threadIdx.x; e < 32 * ncol; e += 256) {
        const int r = e & 31, tap = (e >> 5) % it.taps, ci = (e >> 5) / it.taps, co = co0 + r;
        if (co < it.cout)
          dst[static_cast<size_t>(cc0 + ci) * it.ld + it.col0 + (it.taps - 1 - tap) * it.cpad + co] =
              __float2bfloat16_rn(tile[r][ci * it.taps + tap]);
      }
    }
    __syncthreads();
  }
}

int launch_pack_weights(const dmc_pack_item* items_dev, int n, cudaStream_t st) {
  DMC_REQUIRE(items_dev && n > 0 && n <= 65535, "pack_weights: bad arguments (n=%d)", n);
  pack_weights_kernel<<<dim3(32, n), 256, 0, st>>>(items_dev);
  DMC_CUDA_OK(cudaGetLastError());
  return 0;
}

// dst (+)= src over n bf16 elements (n % 8 == 0): gradient of an identity residual branch / fan-out accumulation
__global__ void __launch_bounds__(256) add_bf16_kernel(uint4* __restrict__ dst, const uint4* __restrict__ src, size_t n8,
                                                       int accumulate) {
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n8; i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    uint4 s = src[i];
    if (accumulate) {
      uint4 d = dst[i];
      const __nv_bfloat162* sp = reinterpret_cast<const __nv_bfloat162*>(&s);
      __nv_bfloat162* dp = reinterpret_cast<__nv_bfloat162*>(&d);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 a = __bfloat1622float2(sp[j]), b = __bfloat1622float2(dp[j]);
        dp[j] = __floats2bfloat162_rn(a.x + b.x, a.y + b.y);
      }
      s = d;
    }
    dst[i] = s;
  }
}

int launch_add_bf16(void* dst, const void* src, size_t n, int accumulate, cudaStream_t st) {
  DMC_REQUIRE(dst && src && n > 0 && n % 8 == 0, "add_bf16: bad arguments");
  const size_t n8 = n / 8;
  const int blocks = static_cast<int>(std::min<size_t>((n8 + 255) / 256, static_cast<size_t>(num_sms()) * 16));
  add_bf16_kernel<<<blocks, 256, 0, st>>>(reinterpret_cast<uint4*>(dst), reinterpret_cast<const uint4*>(src), n8, accumulate);
  DMC_CUDA_OK(cudaGetLastError());
  return 0;
}

// fp32 NCHW [B, Csrc, H, W] -> bf16 NHWC [B, H, W, Cdst] with zero-padded channels (head gradient, stem input)
__global__ void __launch_bounds__(256) nchw_to_nhwc_pad_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst,
                                                               int B, int Cs, int HW, int Cd) {
  const size_t total = static_cast<size_t>(B) * HW * Cd;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(i % Cd);
    const size_t pix = i / Cd;
    const int p = static_cast<int>(pix % HW);
    const size_t n = pix / HW;
    dst[i] = __float2bfloat16(c < Cs ? src[(n * Cs + c) * HW + p] : 0.f);
  }
}

int launch_nchw_to_nhwc_pad(const float* src, void* dst, int B, int Cs, int HW, int Cd, cudaStream_t st) {
  DMC_REQUIRE(src && dst && B > 0 && Cs > 0 && Cd >= Cs, "nchw_to_nhwc_pad: bad arguments");
  const size_t total = static_cast<size_t>(B) * HW * Cd;
  const int blocks = static_cast<int>(std::min<size_t>((total + 255) / 256, static_cast<size_t>(num_sms()) * 16));
  nchw_to_nhwc_pad_kernel<<<blocks, 256, 0, st>>>(src, reinterpret_cast<__nv_bfloat16*>(dst), B, Cs, HW, Cd);
  DMC_CUDA_OK(cudaGetLastError());
  return 0;
}

// Input gradient of a strided 3x3 (padding 1) convolution on the CUDA cores (the three Downsample layers, 0.6 % of the
// FLOPs): dx[n, y, x, ci] (+)= sum_{co, r, s : (y + 1 - r) % stride == 0 ...} dy[n, (y+1-r)/stride, (x+1-s)/stride, co] w[co, ci, r, s]
__global__ void __launch_bounds__(256) conv_dgrad_strided_kernel(const __nv_bfloat16* __restrict__ dy,
                                                                 const float* __restrict__ w, __nv_bfloat16* __restrict__ dx,
                                                                 int B, int Hin, int Win, int Cin, int Cout, int stride,
                                                                 int accumulate) {
  const int Ho = Hin / stride, Wo = Win / stride;
  const size_t total = static_cast<size_t>(B) * Hin * Win * Cin;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int ci = static_cast<int>(i % Cin);
    size_t pix = i / Cin;
    const int x = static_cast<int>(pix % Win), y = static_cast<int>((pix / Win) % Hin);
    const size_t n = pix / (static_cast<size_t>(Win) * Hin);
    float acc = accumulate ? __bfloat162float(dx[i]) : 0.f;
    for (int r = 0; r < 3; ++r) {
      const int ty = y + 1 - r;
      if (ty < 0 || ty % stride != 0 || ty / stride >= Ho) continue;
      for (int s = 0; s < 3; ++s) {
        const int tx = x + 1 - s;
        if (tx < 0 || tx % stride != 0 || tx / stride >= Wo) continue;
        const __nv_bfloat16* dp = dy + ((n * Ho + ty / stride) * Wo + tx / stride) * Cout;
        const float* wp = w + (static_cast<size_t>(ci) * 9 + r * 3 + s);
        for (int co = 0; co < Cout; ++co) acc = fmaf(__bfloat162float(dp[co]), wp[static_cast<size_t>(co) * Cin * 9], acc);
      }
    }
    dx[i] = __float2bfloat16(acc);
  }
}

int launch_conv_dgrad_strided(const void* dy, const float* w, void* dx, int B, int Hin, int Win, int Cin, int Cout, int stride,
                              int accumulate, cudaStream_t st) {
  DMC_REQUIRE(dy && w && dx && B > 0 && stride >= 1 && Hin % stride == 0 && Win % stride == 0, "conv_dgrad_strided: bad arguments");
  const size_t total = static_cast<size_t>(B) * Hin * Win * Cin;
  const int blocks = static_cast<int>(std::min<size_t>((total + 255) / 256, static_cast<size_t>(num_sms()) * 32));
  conv_dgrad_strided_kernel<<<blocks, 256, 0, st>>>(reinterpret_cast<const __nv_bfloat16*>(dy), w,
                                                    reinterpret_cast<__nv_bfloat16*>(dx), B, Hin, Win, Cin, Cout, stride,
                                                    accumulate);
  DMC_CUDA_OK(cudaGetLastError());
  return 0;
}

}  // namespace dmc
