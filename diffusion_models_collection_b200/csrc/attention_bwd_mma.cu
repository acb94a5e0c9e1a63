// Attention backward on tensor cores (training step; reference: autograd of models/unet.py:88-96).
//
//   S = Q K^T * scale,  P = softmax(S),  O = P V          (forward, head dim 64, L <= 256 tokens)
//   dV = P^T dO,  dP = dO V^T,  dS = P o (dP - D) * scale with D_i = dO_i . O_i,  dQ = dS K,  dK = dS^T Q
//
// One CTA per (image, head).  Q, K, V and dO of the head stay resident in shared memory (4 x L x 64 bf16); the kernel makes
// three passes over them with warp-level mma.sync.m16n8k16 (bf16 in, fp32 accumulate), each warp owning 32 rows so that no
// accumulator is ever shared between warps (no atomics, bit-reproducible):
//   pass 1  rows = queries: log-sum-exp of every score row (the forward pass does not save it)
//   pass 2  rows = queries: recompute S, P, dP block by block (64 keys), dQ += dS K
//   pass 3  rows = keys:    recompute S^T, P^T, dP^T block by block (64 queries), dV += P^T dO, dK += dS^T Q
// P and dS are rounded to bf16 for the second GEMM of each pair (the usual flash-attention-2 numerics); everything else fp32.
// The tcgen05 form of this kernel is future work: at the UNet's sizes (L = 256 / 64 on 16x16 / 8x8 maps) the backward of all
// attention layers is < 5 % of a training step once it runs on tensor cores at all.
#include "common.cuh"
#include "kernels.h"

namespace dmc {

constexpr int AB_LD = 72;        // bf16 elements per shared-memory row (64 + 8 pad: conflict-free ldmatrix)
constexpr int AB_LMAX = 256;
constexpr int AB_THREADS = 256;
constexpr size_t AB_SMEM = static_cast<size_t>(4) * AB_LMAX * AB_LD * 2 + 2 * AB_LMAX * 4;

__device__ __forceinline__ void ldsm_x4(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void mma_16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// acc[8][4] (16 rows x 64 columns) = A[rowbase .. +16][0..64) * Bm[colbase .. +64][0..64)^T, both row-major [row][64] in smem
__device__ __forceinline__ void gemm_rows_x_rowsT(float (&acc)[8][4], uint32_t sA, int rowbase, uint32_t sB, int colbase, int lane) {
#pragma unroll
  for (int nt = 0; nt < 8; ++nt)
#pragma unroll
    for (int e = 0; e < 4; ++e) acc[nt][e] = 0.f;
  const int a_row = rowbase + (lane & 15), a_col = (lane >> 4) * 8;
  const int b_row = ((lane >> 4) * 8) + (lane & 7), b_col = ((lane >> 3) & 1) * 8;
#pragma unroll
  for (int ks = 0; ks < 4; ++ks) {
    uint32_t a[4];
    ldsm_x4(a, sA + static_cast<uint32_t>((a_row * AB_LD + ks * 16 + a_col) * 2));
#pragma unroll
    for (int np = 0; np < 4; ++np) {
      uint32_t b[4];
      ldsm_x4(b, sB + static_cast<uint32_t>(((colbase + np * 16 + b_row) * AB_LD + ks * 16 + b_col) * 2));
      mma_16816(acc[2 * np], a, b[0], b[1]);
      mma_16816(acc[2 * np + 1], a, b[2], b[3]);
    }
  }
}

// acc[8][4] (16 x 64) += Areg (16 x 64 as four k16 A fragments) * Bm[kbase .. +64][0..64), Bm row-major [k][64] in smem
__device__ __forceinline__ void gemm_regs_x_rows(float (&acc)[8][4], const uint32_t (&a)[4][4], uint32_t sB, int kbase, int lane) {
  const int b_row = ((lane >> 3) & 1) * 8 + (lane & 7), b_col = (lane >> 4) * 8;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
#pragma unroll
    for (int np = 0; np < 4; ++np) {
      uint32_t b[4];
      ldsm_x4_t(b, sB + static_cast<uint32_t>(((kbase + j * 16 + b_row) * AB_LD + np * 16 + b_col) * 2));
      mma_16816(acc[2 * np], a[j], b[0], b[1]);
      mma_16816(acc[2 * np + 1], a[j], b[2], b[3]);
    }
  }
}

// C fragments of a 16 x 64 tile -> four k16 A fragments (bf16)
__device__ __forceinline__ void pack_a(uint32_t (&a)[4][4], const float (&c)[8][4]) {
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    a[j][0] = pack_bf16x2(c[2 * j][0], c[2 * j][1]);
    a[j][1] = pack_bf16x2(c[2 * j][2], c[2 * j][3]);
    a[j][2] = pack_bf16x2(c[2 * j + 1][0], c[2 * j + 1][1]);
    a[j][3] = pack_bf16x2(c[2 * j + 1][2], c[2 * j + 1][3]);
  }
}

__device__ __forceinline__ void store_rows(__nv_bfloat16* dst, size_t row_stride, int rowbase, int L, const float (&c)[8][4], int lane) {
  const int g = lane >> 2, t = lane & 3;
#pragma unroll
  for (int half = 0; half < 2; ++half) {
    const int row = rowbase + g + half * 8;
    if (row < L) {
      __nv_bfloat16* p = dst + static_cast<size_t>(row) * row_stride + 2 * t;
#pragma unroll
      for (int nt = 0; nt < 8; ++nt)
        *reinterpret_cast<uint32_t*>(p + nt * 8) = pack_bf16x2(c[nt][2 * half], c[nt][2 * half + 1]);
    }
  }
}

__global__ void __launch_bounds__(AB_THREADS, 1)
attention_bwd_mma_kernel(const __nv_bfloat16* __restrict__ qkv, const __nv_bfloat16* __restrict__ out,
                         const __nv_bfloat16* __restrict__ dout, __nv_bfloat16* __restrict__ dqkv, int L, int heads, int C,
                         float scale) {
  extern __shared__ __align__(16) uint8_t ab_smem[];
  __nv_bfloat16* sQ = reinterpret_cast<__nv_bfloat16*>(ab_smem);
  __nv_bfloat16* sK = sQ + AB_LMAX * AB_LD;
  __nv_bfloat16* sV = sK + AB_LMAX * AB_LD;
  __nv_bfloat16* sO = sV + AB_LMAX * AB_LD;  // dO
  float* sLse = reinterpret_cast<float*>(sO + AB_LMAX * AB_LD);
  float* sD = sLse + AB_LMAX;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int n = blockIdx.x / heads, h = blockIdx.x % heads;
  const int Lp = (L + 63) & ~63;
  const size_t tok0 = static_cast<size_t>(n) * L;

  // ---- stage Q, K, V, dO (zero rows beyond L) and D_i = dO_i . O_i
  for (int idx = tid; idx < Lp * 8; idx += AB_THREADS) {
    const int r = idx >> 3, c8 = (idx & 7) * 8;
    uint4 q = make_uint4(0, 0, 0, 0), k = q, v = q, d = q;
    float dd = 0.f;
    if (r < L) {
      const __nv_bfloat16* base = qkv + (tok0 + r) * 3 * C + h * 64 + c8;
      q = *reinterpret_cast<const uint4*>(base);
      k = *reinterpret_cast<const uint4*>(base + C);
      v = *reinterpret_cast<const uint4*>(base + 2 * C);
      d = *reinterpret_cast<const uint4*>(dout + (tok0 + r) * C + h * 64 + c8);
      const uint4 o = *reinterpret_cast<const uint4*>(out + (tok0 + r) * C + h * 64 + c8);
      const uint32_t du[4] = {d.x, d.y, d.z, d.w}, ou[4] = {o.x, o.y, o.z, o.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 a = unpack_bf16x2(du[j]), b = unpack_bf16x2(ou[j]);
        dd += a.x * b.x + a.y * b.y;
      }
    }
    *reinterpret_cast<uint4*>(sQ + r * AB_LD + c8) = q;
    *reinterpret_cast<uint4*>(sK + r * AB_LD + c8) = k;
    *reinterpret_cast<uint4*>(sV + r * AB_LD + c8) = v;
    *reinterpret_cast<uint4*>(sO + r * AB_LD + c8) = d;
    dd += __shfl_xor_sync(0xffffffffu, dd, 1);
    dd += __shfl_xor_sync(0xffffffffu, dd, 2);
    dd += __shfl_xor_sync(0xffffffffu, dd, 4);
    if ((idx & 7) == 0) sD[r] = dd;
  }
  __syncthreads();

  const uint32_t aQ = smem_u32(sQ), aK = smem_u32(sK), aV = smem_u32(sV), aO = smem_u32(sO);
  const int g = lane >> 2, t = lane & 3;
  const float sl2 = scale * 1.4426950408889634f;
  const int nblk = Lp / 64;
  const int r0 = warp * 32;

  // ---- pass 1: base-2 log-sum-exp of every score row
  if (r0 < Lp) {
    for (int mt = 0; mt < 2; ++mt) {
      const int rb = r0 + mt * 16;
      float mx[2] = {-INFINITY, -INFINITY}, sum[2] = {0.f, 0.f};
      for (int kb = 0; kb < nblk; ++kb) {
        float s[8][4];
        gemm_rows_x_rowsT(s, aQ, rb, aK, kb * 64, lane);
        float bm[2] = {-INFINITY, -INFINITY};
#pragma unroll
        for (int nt = 0; nt < 8; ++nt)
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const int col = kb * 64 + nt * 8 + 2 * t + (e & 1);
            s[nt][e] = col < L ? s[nt][e] * sl2 : -INFINITY;
            bm[e >> 1] = fmaxf(bm[e >> 1], s[nt][e]);
          }
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
          bm[hh] = fmaxf(bm[hh], __shfl_xor_sync(0xffffffffu, bm[hh], 1));
          bm[hh] = fmaxf(bm[hh], __shfl_xor_sync(0xffffffffu, bm[hh], 2));
          const float mnew = fmaxf(mx[hh], bm[hh]);
          sum[hh] *= exp2f(mx[hh] - mnew);
          mx[hh] = mnew;
        }
#pragma unroll
        for (int nt = 0; nt < 8; ++nt)
#pragma unroll
          for (int e = 0; e < 4; ++e) sum[e >> 1] += exp2f(s[nt][e] - mx[e >> 1]);
      }
#pragma unroll
      for (int hh = 0; hh < 2; ++hh) {
        sum[hh] += __shfl_xor_sync(0xffffffffu, sum[hh], 1);
        sum[hh] += __shfl_xor_sync(0xffffffffu, sum[hh], 2);
        if (t == 0) sLse[rb + g + hh * 8] = mx[hh] + log2f(sum[hh]);
      }
    }
  }
  __syncthreads();

  // ---- pass 2: dQ (rows = queries)
  if (r0 < L) {
    for (int mt = 0; mt < 2; ++mt) {
      const int rb = r0 + mt * 16;
      if (rb >= L) break;
      const float lse[2] = {sLse[rb + g], sLse[rb + g + 8]};
      const float dr[2] = {sD[rb + g], sD[rb + g + 8]};
      float dq[8][4];
#pragma unroll
      for (int nt = 0; nt < 8; ++nt)
#pragma unroll
        for (int e = 0; e < 4; ++e) dq[nt][e] = 0.f;
      for (int kb = 0; kb < nblk; ++kb) {
        float s[8][4], dp[8][4];
        gemm_rows_x_rowsT(s, aQ, rb, aK, kb * 64, lane);
        gemm_rows_x_rowsT(dp, aO, rb, aV, kb * 64, lane);
#pragma unroll
        for (int nt = 0; nt < 8; ++nt)
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const int col = kb * 64 + nt * 8 + 2 * t + (e & 1);
            const float p = col < L ? exp2f(s[nt][e] * sl2 - lse[e >> 1]) : 0.f;
            s[nt][e] = p * (dp[nt][e] - dr[e >> 1]) * scale;
          }
        uint32_t a[4][4];
        pack_a(a, s);
        gemm_regs_x_rows(dq, a, aK, kb * 64, lane);
      }
      store_rows(dqkv + tok0 * 3 * C + h * 64, static_cast<size_t>(3) * C, rb, L, dq, lane);
    }
  }

  // ---- pass 3: dK, dV (rows = keys)
  if (r0 < L) {
    for (int mt = 0; mt < 2; ++mt) {
      const int rb = r0 + mt * 16;
      if (rb >= L) break;
      float dk[8][4], dv[8][4];
#pragma unroll
      for (int nt = 0; nt < 8; ++nt)
#pragma unroll
        for (int e = 0; e < 4; ++e) dk[nt][e] = dv[nt][e] = 0.f;
      for (int qb = 0; qb < nblk; ++qb) {
        float st[8][4], dpt[8][4];
        gemm_rows_x_rowsT(st, aK, rb, aQ, qb * 64, lane);    // S^T  = K_j Q^T
        gemm_rows_x_rowsT(dpt, aV, rb, aO, qb * 64, lane);   // dP^T = V_j dO^T
#pragma unroll
        for (int nt = 0; nt < 8; ++nt)
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const int col = qb * 64 + nt * 8 + 2 * t + (e & 1);  // query index
            const float p = col < L ? exp2f(st[nt][e] * sl2 - sLse[col]) : 0.f;
            st[nt][e] = p;
            dpt[nt][e] = p * (dpt[nt][e] - sD[col]) * scale;
          }
        uint32_t a[4][4];
        pack_a(a, st);
        gemm_regs_x_rows(dv, a, aO, qb * 64, lane);  // dV += P^T dO
        pack_a(a, dpt);
        gemm_regs_x_rows(dk, a, aQ, qb * 64, lane);  // dK += dS^T Q
      }
      store_rows(dqkv + tok0 * 3 * C + C + h * 64, static_cast<size_t>(3) * C, rb, L, dk, lane);
      store_rows(dqkv + tok0 * 3 * C + 2 * C + h * 64, static_cast<size_t>(3) * C, rb, L, dv, lane);
    }
  }
}

int launch_attention_backward_mma(const dmc_attn_bwd_desc& d, cudaStream_t st) {
  static DeviceOnce attr;
  int attr_dev = 0;
  if (attr.need(&attr_dev)) {
    DMC_CUDA_OK(cudaFuncSetAttribute(attention_bwd_mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(AB_SMEM)));
    attr.done(attr_dev);
  }
  attention_bwd_mma_kernel<<<d.B * d.heads, AB_THREADS, AB_SMEM, st>>>(
      reinterpret_cast<const __nv_bfloat16*>(d.qkv), reinterpret_cast<const __nv_bfloat16*>(d.out),
      reinterpret_cast<const __nv_bfloat16*>(d.dout), reinterpret_cast<__nv_bfloat16*>(d.dqkv), d.L, d.heads, d.C, 0.125f);
  DMC_CUDA_OK(cudaGetLastError());
  return 0;
}

}  // namespace dmc
