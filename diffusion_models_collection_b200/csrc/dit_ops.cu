// Memory-bound kernels of the DiT forward (/root/reference/models/dit.py): adaLN conditioning table, patch embedding
// (+ positional embedding) into the fp32 token stream, LayerNorm + adaLN modulate into the bf16 GEMM operand.
// The GEMMs (qkv / out_proj / MLP / final linear) run on conv_umma.cu as 1x1 "convolutions" over the token grid with
// GELU / gate / fp32-residual / unpatchify epilogues; attention runs on attention_umma.cu.
#include "common.cuh"
#include "kernels.h"

namespace dmc {

// =============================================================================================
// Conditioning rows.  Row r is one distinct (t, label) pair:
//   uniform_t == 0: r = image index          (t_r = t[r], label_r = clamp(y[r]))
//   uniform_t == 1: r = label index          (t_r = t[0], label_r = r)          -- sampling: num_classes + 1 rows
// h1[r, :] = SiLU(W1 . [cos(t f) | sin(t f)] + b1)            models/dit.py:42-50 (cos first, divisor half), :34-36
// =============================================================================================
__global__ void __launch_bounds__(256) dit_hidden_kernel(const int64_t* __restrict__ t, const float* __restrict__ freqs,
                                                         const float* __restrict__ w1, const float* __restrict__ b1,
                                                         float* __restrict__ h1, int half, int hidden, int uniform_t) {
  extern __shared__ float emb[];  // 2*half
  const int r = blockIdx.x;
  const float tf = static_cast<float>(t[uniform_t ? 0 : r]);  // t[:, None].float() * freqs (dit.py:46)
  for (int i = threadIdx.x; i < half; i += blockDim.x) {
    const float arg = __fmul_rn(tf, freqs[i]);
    emb[i] = cosf(arg);
    emb[half + i] = sinf(arg);
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  const int K = 2 * half;
  for (int j = blockIdx.y * nw + warp; j < hidden; j += gridDim.y * nw) {
    float acc = 0.f;
    for (int k = lane; k < K; k += 32) acc = fmaf(w1[static_cast<size_t>(j) * K + k], emb[k], acc);
#pragma unroll
    for (int o = 16; o; o >>= 1) acc += __shfl_xor_sync(0xFFFFFFFFu, acc, o);
    if (lane == 0) {
      const float v = acc + b1[j];
      h1[static_cast<size_t>(r) * hidden + j] = v / (1.0f + expf(-v));
    }
  }
}

// sc[r, j] = SiLU(b2[j] + W2[j, :] . h1[r, :] + emb[label_r, j])   (c = t_emb + y_emb, then the SiLU every adaLN applies)
__global__ void __launch_bounds__(256) dit_c_kernel(const float* __restrict__ h1, const float* __restrict__ w2,
                                                    const float* __restrict__ b2, const float* __restrict__ emb,
                                                    const int64_t* __restrict__ y, float* __restrict__ sc, int hidden,
                                                    int uniform_t, int num_classes) {
  const int r = blockIdx.x;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  const float* row_in = h1 + static_cast<size_t>(r) * hidden;
  long long lab = 0;
  if (emb != nullptr) {
    lab = uniform_t ? r : y[r];
    lab = lab < 0 ? 0 : (lab > num_classes ? num_classes : lab);  // torch.clamp(y, 0, num_classes), dit.py:280
  }
  for (int j = blockIdx.y * nw + warp; j < hidden; j += gridDim.y * nw) {
    float acc = 0.f;
    for (int k = lane; k < hidden; k += 32) acc = fmaf(w2[static_cast<size_t>(j) * hidden + k], row_in[k], acc);
#pragma unroll
    for (int o = 16; o; o >>= 1) acc += __shfl_xor_sync(0xFFFFFFFFu, acc, o);
    if (lane == 0) {
      float v = acc + b2[j];
      if (emb != nullptr) v += emb[static_cast<size_t>(lab) * hidden + j];
      sc[static_cast<size_t>(r) * hidden + j] = v / (1.0f + expf(-v));
    }
  }
}

// rows[r, j] = b[j] + W[j, :] . sc[r, :]: one warp per output column keeps its weight row in registers and walks the rows
template <int KPL>  // hidden / 32
__global__ void __launch_bounds__(256) dit_project_kernel(const float* __restrict__ sc, const float* __restrict__ w,
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
  for (int r = blockIdx.y; r < R; r += gridDim.y) {
    const float* s = sc + static_cast<size_t>(r) * K;
    float acc = 0.f;
#pragma unroll
    for (int i = 0; i < KPL; ++i) acc = fmaf(wr[i], s[lane + 32 * i], acc);
#pragma unroll
    for (int o = 16; o; o >>= 1) acc += __shfl_xor_sync(0xFFFFFFFFu, acc, o);
    if (lane == 0) out[static_cast<size_t>(r) * N + j] = acc + bj;
  }
}

__global__ void __launch_bounds__(256) dit_project_generic_kernel(const float* __restrict__ sc, const float* __restrict__ w,
                                                                  const float* __restrict__ b, float* __restrict__ out,
                                                                  int K, int N) {
  const int r = blockIdx.x;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  const float* row_in = sc + static_cast<size_t>(r) * K;
  for (int j = blockIdx.y * nw + warp; j < N; j += gridDim.y * nw) {
    float acc = 0.f;
    for (int k = lane; k < K; k += 32) acc = fmaf(w[static_cast<size_t>(j) * K + k], row_in[k], acc);
#pragma unroll
    for (int o = 16; o; o >>= 1) acc += __shfl_xor_sync(0xFFFFFFFFu, acc, o);
    if (lane == 0) out[static_cast<size_t>(r) * N + j] = acc + b[j];
  }
}

// mod[n, :] = rows[row_of(n), :]
__global__ void __launch_bounds__(256) dit_gather_kernel(const float* __restrict__ rows, const int64_t* __restrict__ y,
                                                         float* __restrict__ mod, int B, int ncols, int uniform_t,
                                                         int has_emb, int num_classes) {
  const int n4 = ncols >> 2;
  const size_t total = static_cast<size_t>(B) * n4;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int n = static_cast<int>(i / n4), c = static_cast<int>(i % n4);
    long long r = n;
    if (uniform_t) {
      r = 0;
      if (has_emb) {
        r = y[n];
        r = r < 0 ? 0 : (r > num_classes ? num_classes : r);
      }
    }
    reinterpret_cast<float4*>(mod + static_cast<size_t>(n) * ncols)[c] =
        __ldg(reinterpret_cast<const float4*>(rows + static_cast<size_t>(r) * ncols) + c);
  }
}

int dit_cond_num_launches(const dmc_dit_cond_desc&) { return 4; }

int launch_dit_cond(const dmc_dit_cond_desc& d, cudaStream_t st) {
  DMC_REQUIRE(d.t && d.freqs && d.w1 && d.b1 && d.w2 && d.b2 && d.w_all && d.b_all && d.scratch && d.mod,
              "dit_cond: null pointer argument");
  DMC_REQUIRE(d.B > 0 && d.ncols % 4 == 0 && d.hidden % 32 == 0 && d.freq_dim % 2 == 0 && d.freq_dim > 0,
              "dit_cond: unsupported shape (B=%d ncols=%d hidden=%d freq_dim=%d)", d.B, d.ncols, d.hidden, d.freq_dim);
  const bool has_emb = d.emb != nullptr && d.y != nullptr;
  const int R = d.uniform_t ? (has_emb ? d.num_classes + 1 : 1) : d.B;
  float* h1 = d.scratch;
  float* sc = h1 + static_cast<size_t>(R) * d.hidden;
  float* rows = sc + static_cast<size_t>(R) * d.hidden;
  dit_hidden_kernel<<<dim3(R, 4), 256, d.freq_dim * sizeof(float), st>>>(d.t, d.freqs, d.w1, d.b1, h1, d.freq_dim / 2,
                                                                         d.hidden, d.uniform_t);
  dit_c_kernel<<<dim3(R, 4), 256, 0, st>>>(h1, d.w2, d.b2, has_emb ? d.emb : nullptr, d.y, sc, d.hidden, d.uniform_t,
                                           d.num_classes);
  const dim3 pg((d.ncols + 7) / 8, std::min(R, 8));
  switch (d.hidden) {
    case 256: dit_project_kernel<8><<<pg, 256, 0, st>>>(sc, d.w_all, d.b_all, rows, R, d.ncols); break;
    case 384: dit_project_kernel<12><<<pg, 256, 0, st>>>(sc, d.w_all, d.b_all, rows, R, d.ncols); break;
    case 512: dit_project_kernel<16><<<pg, 256, 0, st>>>(sc, d.w_all, d.b_all, rows, R, d.ncols); break;
    case 768: dit_project_kernel<24><<<pg, 256, 0, st>>>(sc, d.w_all, d.b_all, rows, R, d.ncols); break;
    case 1024: dit_project_kernel<32><<<pg, 256, 0, st>>>(sc, d.w_all, d.b_all, rows, R, d.ncols); break;
    default: dit_project_generic_kernel<<<dim3(R, 16), 256, 0, st>>>(sc, d.w_all, d.b_all, rows, d.hidden, d.ncols); break;
  }
  const size_t total = static_cast<size_t>(d.B) * (d.ncols / 4);
  const int blocks = static_cast<int>(std::min<size_t>((total + 255) / 256, static_cast<size_t>(num_sms()) * 8));
  dit_gather_kernel<<<blocks, 256, 0, st>>>(rows, d.y, d.mod, d.B, d.ncols, d.uniform_t, has_emb ? 1 : 0, d.num_classes);
  DMC_CUDA_OK(cudaGetLastError());
  return 0;
}

// =============================================================================================
// PatchEmbed (conv p x p, stride p) + bias + positional embedding: fp32 NCHW image -> fp32 token stream
// models/dit.py:21-27, 265.  K = Cin * p * p is tiny (12): memory-bound (fp32 out = hidden * 4 B per token).
// One warp per token: the K patch values are warp-uniform registers, lanes own 4 consecutive output channels.
// =============================================================================================
template <int KMAX, int T, int PT>  // T tokens per warp; PT: patch size known at compile time (0: run-time P)
__global__ void __launch_bounds__(256) patch_embed_kernel(const float* __restrict__ x, const float* __restrict__ wT,
                                                          const float* __restrict__ bias, const float* __restrict__ pos,
                                                          float* __restrict__ out, int x_batch, int B, int Cin, int H, int W,
                                                          int P, int hidden) {
  // T = 4 tokens per warp (K <= 16): the K weight rows of a 128-channel chunk are loaded once and used for four tokens (one token per warp
  // re-read all K * hidden weights per token through the LSU: 0.55 ms for 1024 images, 8x its HBM time).  Every output is still
  // the same chain of K fused multiply-adds in k order: bit-identical to the one-token form.
  if (PT > 0) P = PT;  // the (ci, pi, qi) of every k become constants of the unrolled loops: with a run-time P the three
                       // integer divisions per patch value cost more than the whole arithmetic of the kernel
  const int Ht = H / P, Wt = W / P, L = Ht * Wt, K = Cin * P * P;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const size_t total = static_cast<size_t>(B) * L;
  const size_t tok0 = (static_cast<size_t>(blockIdx.x) * 8 + warp) * T;
  if (tok0 >= total) return;
  float pv[T][KMAX];
#pragma unroll
  for (int t = 0; t < T; ++t) {
    const size_t tok = min(tok0 + t, total - 1);
    const int n = static_cast<int>(tok / L), l = static_cast<int>(tok % L);
    const int ti = l / Wt, tj = l % Wt;
    const float* xin = x + static_cast<size_t>(n % x_batch) * Cin * H * W;
#pragma unroll
    for (int k = 0; k < KMAX; ++k) {
      if (k < K) {  // weight layout [hidden, Cin, P, P]: k = (ci * P + pi) * P + qi
        const int ci = k / (P * P), pi = (k / P) % P, qi = k % P;
        pv[t][k] = __ldg(xin + (static_cast<size_t>(ci) * H + ti * P + pi) * W + tj * P + qi);
      } else {
        pv[t][k] = 0.f;
      }
    }
  }
  for (int c = lane * 4; c < hidden; c += 128) {
    const float4 bv = __ldg(reinterpret_cast<const float4*>(bias + c));
    float4 acc[T];
#pragma unroll
    for (int t = 0; t < T; ++t) acc[t] = bv;
#pragma unroll
    for (int k = 0; k < KMAX; ++k) {
      if (k < K) {
        const float4 wv = __ldg(reinterpret_cast<const float4*>(wT + static_cast<size_t>(k) * hidden + c));
#pragma unroll
        for (int t = 0; t < T; ++t) {
          acc[t].x = fmaf(pv[t][k], wv.x, acc[t].x);
          acc[t].y = fmaf(pv[t][k], wv.y, acc[t].y);
          acc[t].z = fmaf(pv[t][k], wv.z, acc[t].z);
          acc[t].w = fmaf(pv[t][k], wv.w, acc[t].w);
        }
      }
    }
#pragma unroll
    for (int t = 0; t < T; ++t) {
      const size_t tok = tok0 + t;
      if (tok < total) {
        const int l = static_cast<int>(tok % L);
        const float4 pp = __ldg(reinterpret_cast<const float4*>(pos + static_cast<size_t>(l) * hidden + c));
        // (conv + bias) + pos_embed, the reference's order (dit.py:265)
        *reinterpret_cast<float4*>(out + tok * hidden + c) =
            make_float4(acc[t].x + pp.x, acc[t].y + pp.y, acc[t].z + pp.z, acc[t].w + pp.w);
      }
    }
  }
}

int launch_patch_embed(const dmc_patch_embed_desc& d, cudaStream_t st) {
  DMC_REQUIRE(d.x && d.weight && d.bias && d.pos && d.out, "patch_embed: null pointer argument");
  DMC_REQUIRE(d.B > 0 && d.x_batch > 0 && d.patch > 0 && d.H % d.patch == 0 && d.W % d.patch == 0 && d.hidden % 4 == 0,
              "patch_embed: unsupported shape");
  const int K = d.Cin * d.patch * d.patch;
  DMC_REQUIRE(K <= 64, "patch_embed: Cin * patch^2 = %d > 64 is not supported", K);
  const size_t tokens = static_cast<size_t>(d.B) * (d.H / d.patch) * (d.W / d.patch);
  // `weight` here is the TRANSPOSED [K, hidden] copy the host packs once (coalesced per-channel reads)
  if (K <= 16 && d.patch == 2)  // the shipped configs (3 x 2 x 2): 8 warps x 4 tokens per block
    patch_embed_kernel<16, 4, 2><<<static_cast<int>((tokens + 31) / 32), 256, 0, st>>>(d.x, d.weight, d.bias, d.pos, d.out, d.x_batch,
                                                                                        d.B, d.Cin, d.H, d.W, d.patch, d.hidden);
  else if (K <= 16)
    patch_embed_kernel<16, 4, 0><<<static_cast<int>((tokens + 31) / 32), 256, 0, st>>>(d.x, d.weight, d.bias, d.pos, d.out, d.x_batch,
                                                                                        d.B, d.Cin, d.H, d.W, d.patch, d.hidden);
  else
    patch_embed_kernel<64, 1, 0><<<static_cast<int>((tokens + 7) / 8), 256, 0, st>>>(d.x, d.weight, d.bias, d.pos, d.out, d.x_batch, d.B,
                                                                                      d.Cin, d.H, d.W, d.patch, d.hidden);
  DMC_CUDA_OK(cudaGetLastError());
  return 0;
}

// =============================================================================================
// LayerNorm (no affine) + adaLN modulate: fp32 token stream -> bf16 GEMM operand.  One warp per token, the row lives
// in registers (two-pass mean / variance like ATen's, fp32), 128-bit loads, 64-bit bf16x4 stores.
// =============================================================================================
template <int V4>  // float4 per lane: C = 128 * V4
__global__ void __launch_bounds__(256) ln_modulate_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ out,
                                                          __nv_bfloat16* __restrict__ out_lo,
                                                          const float* __restrict__ shift, const float* __restrict__ scale,
                                                          int mod_stride, size_t tokens, int L, float eps) {
  // TWO token rows per warp (all loads of both rows issued before the first reduction: twice the bytes in flight per warp -- one
  // row per warp ran at 4.3 TB/s, latency-bound on its load -> shuffle -> shuffle -> store chain)
  constexpr int C = 128 * V4;
  constexpr int R = 2;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const size_t tok0 = (static_cast<size_t>(blockIdx.x) * 8 + warp) * R;
  if (tok0 >= tokens) return;
  float4 v[R][V4];
  float s[R];
#pragma unroll
  for (int r = 0; r < R; ++r) {
    const size_t tok = min(tok0 + r, tokens - 1);  // (an odd tail row is computed twice and stored once)
    const float4* xr = reinterpret_cast<const float4*>(x + tok * C);
#pragma unroll
    for (int i = 0; i < V4; ++i) v[r][i] = xr[lane + 32 * i];
  }
#pragma unroll
  for (int r = 0; r < R; ++r) {
    s[r] = 0.f;
#pragma unroll
    for (int i = 0; i < V4; ++i) s[r] += (v[r][i].x + v[r][i].y) + (v[r][i].z + v[r][i].w);
  }
#pragma unroll
  for (int o = 16; o; o >>= 1) {
#pragma unroll
    for (int r = 0; r < R; ++r) s[r] += __shfl_xor_sync(0xFFFFFFFFu, s[r], o);
  }
  float mean[R], ss[R];
#pragma unroll
  for (int r = 0; r < R; ++r) {
    mean[r] = s[r] * (1.0f / C);
    ss[r] = 0.f;
#pragma unroll
    for (int i = 0; i < V4; ++i) {
      const float a = v[r][i].x - mean[r], b = v[r][i].y - mean[r], c = v[r][i].z - mean[r], e = v[r][i].w - mean[r];
      ss[r] += (a * a + b * b) + (c * c + e * e);
    }
  }
#pragma unroll
  for (int o = 16; o; o >>= 1) {
#pragma unroll
    for (int r = 0; r < R; ++r) ss[r] += __shfl_xor_sync(0xFFFFFFFFu, ss[r], o);
  }
#pragma unroll
  for (int r = 0; r < R; ++r) {
    const size_t tok = tok0 + r;
    if (tok >= tokens) break;
    const int n = static_cast<int>(tok / L);
    const float rstd = rsqrtf(ss[r] * (1.0f / C) + eps);
    const float4* sh4 = reinterpret_cast<const float4*>(shift + static_cast<size_t>(n) * mod_stride);
    const float4* sc4 = reinterpret_cast<const float4*>(scale + static_cast<size_t>(n) * mod_stride);
    uint2* o2 = reinterpret_cast<uint2*>(out + tok * C);
#pragma unroll
    for (int i = 0; i < V4; ++i) {
      const float4 sh = __ldg(sh4 + lane + 32 * i), sc = __ldg(sc4 + lane + 32 * i);
      const float y0 = fmaf((v[r][i].x - mean[r]) * rstd, 1.0f + sc.x, sh.x);
      const float y1 = fmaf((v[r][i].y - mean[r]) * rstd, 1.0f + sc.y, sh.y);
      const float y2 = fmaf((v[r][i].z - mean[r]) * rstd, 1.0f + sc.z, sh.z);
      const float y3 = fmaf((v[r][i].w - mean[r]) * rstd, 1.0f + sc.w, sh.w);
      const uint2 hi = make_uint2(pack_bf16x2(y0, y1), pack_bf16x2(y2, y3));
      o2[lane + 32 * i] = hi;
      if (out_lo != nullptr) {  // split-bf16 mode: the rounding remainder
        const float2 h0 = unpack_bf16x2(hi.x), h1 = unpack_bf16x2(hi.y);
        reinterpret_cast<uint2*>(out_lo + tok * C)[lane + 32 * i] =
            make_uint2(pack_bf16x2(y0 - h0.x, y1 - h0.y), pack_bf16x2(y2 - h1.x, y3 - h1.y));
      }
    }
  }
}

int launch_ln_modulate(const dmc_ln_mod_desc& d, cudaStream_t st) {
  DMC_REQUIRE(d.x && d.out && d.shift && d.scale, "ln_modulate: null pointer argument");
  DMC_REQUIRE(d.B > 0 && d.L > 0 && d.C % 128 == 0 && d.C >= 128 && d.C <= 1024 && d.mod_stride % 4 == 0,
              "ln_modulate: C=%d must be a multiple of 128 in [128, 1024]", d.C);
  const size_t tokens = static_cast<size_t>(d.B) * d.L;
  const int blocks = static_cast<int>((tokens + 15) / 16);  // 8 warps x 2 rows
  __nv_bfloat16* out = reinterpret_cast<__nv_bfloat16*>(d.out);
  __nv_bfloat16* out_lo = reinterpret_cast<__nv_bfloat16*>(d.out_lo);
#define LNM(V) ln_modulate_kernel<V><<<blocks, 256, 0, st>>>(d.x, out, out_lo, d.shift, d.scale, d.mod_stride, tokens, d.L, d.eps)
  switch (d.C / 128) {
    case 1: LNM(1); break;
    case 2: LNM(2); break;
    case 3: LNM(3); break;
    case 4: LNM(4); break;
    case 5: LNM(5); break;
    case 6: LNM(6); break;
    case 7: LNM(7); break;
    default: LNM(8); break;
  }
#undef LNM
  DMC_CUDA_OK(cudaGetLastError());
  return 0;
}

}  // namespace dmc
