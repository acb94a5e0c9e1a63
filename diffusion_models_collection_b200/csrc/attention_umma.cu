// Multi-head self-attention core softmax(Q K^T / sqrt(hd)) V on the 5th-generation tensor cores, head dim 64.
// Replaces /root/reference/models/unet.py:88-96 (2 bmm + softmax + 2 permute copies, the L x L fp32 score matrix
// materialised in HBM) and the attention core of nn.MultiheadAttention in models/dit.py:94,123.
//
// Data: qkv is the bf16 token matrix [B*L, 3C] written by the qkv GEMM, channel order [q|k|v][head][64]; out is
// [B*L, C].  One work item = 128 consecutive token rows x one head:
//   L >= 128 : the 128 queries of one image attend to all L keys of that image (L <= 256: ONE key tile, so the
//              softmax is exact in one pass -- no online rescaling needed).  Items of one (image, head) are processed
//              back to back by the same CTA and share the K / V tiles in shared memory.
//   L <  128 : the tile holds 128/L whole images; scores are computed for the 128 x 128 block and everything outside
//              the block diagonal (other images) is masked to probability 0.
// Pipeline of one CTA (192 threads):
//   warp 4 lane 0 : TMA producer  (Q tile {64 x 128}, K and V tiles {64 x keys}, SWIZZLE_128B)
//   warp 5 lane 0 : tcgen05.mma issuer:  S[128 x keys] = Q K^T  (A, B K-major)   -> TMEM columns [0, keys)
//                                        O[128 x 64]   = P V    (A = P from smem, B = V MN-major) -> columns [256, 320)
//   warps 0..3    : softmax: tcgen05.ld S row (thread = query row), max / exp2 / sum in fp32, P as bf16 into shared
//                   memory in the K-major SWIZZLE_128B operand layout; then O * (1/sum) -> bf16 -> global.
#include "common.cuh"
#include "kernels.h"

#include <new>

namespace dmc {

constexpr int AT_M = 128;
constexpr int AT_HD = 64;
constexpr int AT_MAXKEYS = 256;
constexpr int AT_Q_BYTES = AT_M * AT_HD * 2;             // 16 KB
constexpr int AT_KV_BYTES = AT_MAXKEYS * AT_HD * 2;      // 32 KB each
constexpr int AT_P_BYTES = AT_M * AT_MAXKEYS * 2;        // 64 KB
constexpr int AT_O_COL = 256;
constexpr size_t AT_SMEM = AT_Q_BYTES + 2 * AT_KV_BYTES + AT_P_BYTES + 1024 + 256;

struct AttnParams {
  int L, heads, C;
  int keys;            // key rows per tile: max(L, 128)
  int qtiles;          // q tiles per item group: L / 128 (>= 1)
  int groups;          // item groups: L >= 128 ? B * heads : ceil(B*L / 128) * heads
  int total_rows;      // B * L
  float scale_log2e;
  __nv_bfloat16* out;
};

struct AttnPrepared {
  CUtensorMap tmQ, tmKV;
  AttnParams p;
  int grid;
};

// MN-major SWIZZLE_128B descriptor (operand rows = K index, 128-byte rows of 64 MN elements): 8-row groups 1024 B apart
__device__ __forceinline__ uint64_t umma_desc_mn_sw128(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFF) >> 4);
  d |= static_cast<uint64_t>(1) << 16;            // LBO: stride between 64-element MN blocks (single block: unused)
  d |= static_cast<uint64_t>(1024 >> 4) << 32;    // SBO: stride between groups of 8 K rows
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(2) << 61;
  return d;
}

__global__ void __launch_bounds__(192, 1)
attention_umma_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmKV,
                      const __grid_constant__ AttnParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint8_t* sQ = smem;
  uint8_t* sK = sQ + AT_Q_BYTES;
  uint8_t* sV = sK + AT_KV_BYTES;
  uint8_t* sP = sV + AT_KV_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sP + AT_P_BYTES);
  uint64_t* kv_full = bars + 0;
  uint64_t* kv_empty = bars + 1;
  uint64_t* q_full = bars + 2;
  uint64_t* q_empty = bars + 3;
  uint64_t* s_full = bars + 4;
  uint64_t* s_empty = bars + 5;
  uint64_t* p_full = bars + 6;
  uint64_t* o_full = bars + 7;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 4 && lane == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmKV);
    mbar_init(kv_full, 1);
    mbar_init(kv_empty, 1);
    mbar_init(q_full, 1);
    mbar_init(q_empty, 1);
    mbar_init(s_full, 1);
    mbar_init(s_empty, 128);
    mbar_init(p_full, 128);
    mbar_init(o_full, 1);
    mbar_fence_init();
  }
  if (warp == 5) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int keys = p.keys;
  const bool small = p.L < AT_M;

  if (warp == 4) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      uint32_t it = 0, qi = 0;
      for (int g = blockIdx.x; g < p.groups; g += gridDim.x, ++it) {
        const int h = g % p.heads;
        const int blk = g / p.heads;  // image (L >= 128) or 128-row block (L < 128)
        const int key_row0 = small ? blk * AT_M : blk * p.L;
        mbar_wait(kv_empty, (it & 1u) ^ 1u);
        mbar_expect_tx(kv_full, static_cast<uint32_t>(2 * keys * AT_HD * 2));
        tma_load_2d(sK, &tmKV, kv_full, p.C + h * AT_HD, key_row0);
        tma_load_2d(sV, &tmKV, kv_full, 2 * p.C + h * AT_HD, key_row0);
        for (int qt = 0; qt < p.qtiles; ++qt, ++qi) {
          mbar_wait(q_empty, (qi & 1u) ^ 1u);
          mbar_expect_tx(q_full, AT_Q_BYTES);
          tma_load_2d(sQ, &tmQ, q_full, h * AT_HD, key_row0 + qt * AT_M);
        }
      }
    }
  } else if (warp == 5) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      const uint32_t idesc_s = umma_idesc_bf16(AT_M, keys);
      const uint32_t idesc_o = umma_idesc_bf16(AT_M, AT_HD, /*b_mn_major=*/1);
      const uint64_t qdesc = umma_desc_k_sw128(smem_u32(sQ));
      const uint64_t kdesc = umma_desc_k_sw128(smem_u32(sK));
      uint32_t it = 0, qi = 0;
      for (int g = blockIdx.x; g < p.groups; g += gridDim.x, ++it) {
        mbar_wait(kv_full, it & 1u);
        for (int qt = 0; qt < p.qtiles; ++qt, ++qi) {
          mbar_wait(q_full, qi & 1u);
          mbar_wait(s_empty, (qi & 1u) ^ 1u);
          tc_fence_after();
#pragma unroll
          for (int k = 0; k < AT_HD / 16; ++k) umma_bf16(tmem_base, qdesc + 2 * k, kdesc + 2 * k, idesc_s, k != 0 ? 1u : 0u);
          umma_commit(q_empty);
          umma_commit(s_full);
          mbar_wait(p_full, qi & 1u);
          tc_fence_after();
          for (int j = 0; j < keys / 16; ++j) {
            const uint64_t pdesc = umma_desc_k_sw128(smem_u32(sP) + (j >> 2) * (AT_M * 128)) + 2 * (j & 3);
            const uint64_t vdesc = umma_desc_mn_sw128(smem_u32(sV) + j * 16 * 128);
            umma_bf16(tmem_base + AT_O_COL, pdesc, vdesc, idesc_o, j != 0 ? 1u : 0u);
          }
          umma_commit(o_full);
          if (qt == p.qtiles - 1) umma_commit(kv_empty);
        }
      }
    }
  } else {
    // ===================== softmax + output (thread = query row) =====================
    const int row = warp * 32 + lane;
    const uint32_t lane_addr = tmem_base + (static_cast<uint32_t>(warp * 32) << 16);
    uint32_t qi = 0;
    for (int g = blockIdx.x; g < p.groups; g += gridDim.x) {
      const int h = g % p.heads;
      const int blk = g / p.heads;
      const int key_row0 = small ? blk * AT_M : blk * p.L;
      // valid key columns of this row: the block-diagonal segment when several images share the tile
      const int c_lo = small ? (row / p.L) * p.L : 0;
      const int c_hi = small ? c_lo + p.L : keys;
      // warp-uniform hull of the valid columns (tcgen05.ld is warp-collective: the skip test below must not diverge)
      const int w_lo = small ? ((warp * 32) / p.L) * p.L : 0;
      const int w_hi = small ? ((warp * 32 + 31) / p.L + 1) * p.L : keys;
      for (int qt = 0; qt < p.qtiles; ++qt, ++qi) {
        mbar_wait(s_full, qi & 1u);
        tc_fence_after();
        // pass 1: row max
        float m = -INFINITY;
        for (int c0 = 0; c0 < keys; c0 += 32) {
          if (c0 + 32 <= w_lo || c0 >= w_hi) continue;
          uint32_t r[32];
          tmem_ld_32x32(lane_addr + c0, r);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const int c = c0 + j;
            if (c >= c_lo && c < c_hi) m = fmaxf(m, __uint_as_float(r[j]));
          }
        }
        const float ms = m * p.scale_log2e;
        // pass 2: p = exp2(s * scale - max * scale), row sum, bf16 P into the swizzled operand layout
        float sum = 0.f;
        for (int c0 = 0; c0 < keys; c0 += 32) {
          uint32_t r[32];
          tmem_ld_32x32(lane_addr + c0, r);
          tmem_ld_wait();
          uint32_t pk[16];
#pragma unroll
          for (int j = 0; j < 32; j += 2) {
            const int c = c0 + j;
            float e0 = (c >= c_lo && c < c_hi) ? exp2f(fmaf(__uint_as_float(r[j]), p.scale_log2e, -ms)) : 0.f;
            float e1 = (c + 1 >= c_lo && c + 1 < c_hi) ? exp2f(fmaf(__uint_as_float(r[j + 1]), p.scale_log2e, -ms)) : 0.f;
            // the row sum uses the bf16-rounded probabilities the tensor core will actually multiply
            __nv_bfloat162 b = __floats2bfloat162_rn(e0, e1);
            float2 f = __bfloat1622float2(b);
            sum += f.x + f.y;
            pk[j >> 1] = *reinterpret_cast<uint32_t*>(&b);
          }
          uint8_t* sub = sP + (c0 >> 6) * (AT_M * 128) + row * 128;
          const int chunk0 = (c0 & 63) >> 3;  // first 16-byte chunk of these 32 keys inside the 128-byte row
#pragma unroll
          for (int q4 = 0; q4 < 4; ++q4) {
            const int chunk = (chunk0 + q4) ^ (row & 7);
            *reinterpret_cast<uint4*>(sub + chunk * 16) = make_uint4(pk[4 * q4], pk[4 * q4 + 1], pk[4 * q4 + 2], pk[4 * q4 + 3]);
          }
        }
        fence_proxy_async_smem();
        tc_fence_before();
        mbar_arrive(p_full);
        // O = P V done -> normalise, store
        mbar_wait(o_full, qi & 1u);
        tc_fence_after();
        const float inv = 1.0f / sum;
        const int grow = key_row0 + qt * AT_M + row;
        __nv_bfloat16* op = p.out + static_cast<size_t>(grow) * p.C + h * AT_HD;
#pragma unroll
        for (int c0 = 0; c0 < AT_HD; c0 += 32) {
          uint32_t r[32];
          tmem_ld_32x32(lane_addr + AT_O_COL + c0, r);
          tmem_ld_wait();
          if (grow < p.total_rows) {
#pragma unroll
            for (int q4 = 0; q4 < 4; ++q4) {
              uint4 u;
              u.x = pack_bf16x2(__uint_as_float(r[8 * q4]) * inv, __uint_as_float(r[8 * q4 + 1]) * inv);
              u.y = pack_bf16x2(__uint_as_float(r[8 * q4 + 2]) * inv, __uint_as_float(r[8 * q4 + 3]) * inv);
              u.z = pack_bf16x2(__uint_as_float(r[8 * q4 + 4]) * inv, __uint_as_float(r[8 * q4 + 5]) * inv);
              u.w = pack_bf16x2(__uint_as_float(r[8 * q4 + 6]) * inv, __uint_as_float(r[8 * q4 + 7]) * inv);
              reinterpret_cast<uint4*>(op + c0)[q4] = u;
            }
          }
        }
        tc_fence_before();
        mbar_arrive(s_empty);
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 5) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

static int encode2d(CUtensorMap* m, const void* base, uint64_t cols, uint64_t rows, uint32_t box_rows) {
  EncodeTiledFn fn = encode_tiled_fn();
  DMC_REQUIRE(fn != nullptr, "attention: cuTensorMapEncodeTiled unavailable -- call dmc_init()");
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {cols * 2};
  cuuint32_t box[2] = {AT_HD, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  DMC_REQUIRE(r == CUDA_SUCCESS, "attention: cuTensorMapEncodeTiled failed with CUresult %d", static_cast<int>(r));
  return 0;
}

bool attention_umma_supported(const dmc_attn_desc& d) {
  if (d.heads <= 0 || d.C != d.heads * AT_HD) return false;
  const int L = d.L;
  if (L >= AT_M) return L == 128 || L == 256;
  return L >= 1 && (AT_M % L) == 0 && L >= 16;  // L in {16, 32, 64}
}

int attention_prepare(const dmc_attn_desc& d, AttnPrepared** out) {
  DMC_REQUIRE(d.qkv && d.out && d.B > 0 && d.L > 0, "attention: bad arguments");
  DMC_REQUIRE(attention_umma_supported(d), "attention: unsupported shape for the tcgen05 kernel (L=%d C=%d heads=%d)", d.L,
              d.C, d.heads);
  AttnPrepared* P = new (std::nothrow) AttnPrepared();
  DMC_REQUIRE(P != nullptr, "attention: out of host memory");
  AttnParams& p = P->p;
  p.L = d.L; p.heads = d.heads; p.C = d.C;
  p.keys = d.L < AT_M ? AT_M : d.L;
  p.qtiles = d.L < AT_M ? 1 : d.L / AT_M;
  p.total_rows = d.B * d.L;
  p.groups = (d.L < AT_M ? (p.total_rows + AT_M - 1) / AT_M : d.B) * d.heads;
  p.scale_log2e = 1.4426950408889634f / sqrtf(static_cast<float>(AT_HD));
  p.out = reinterpret_cast<__nv_bfloat16*>(d.out);
  const uint64_t cols = 3ull * d.C, rows = static_cast<uint64_t>(p.total_rows);
  if (encode2d(&P->tmQ, d.qkv, cols, rows, AT_M) != 0 || encode2d(&P->tmKV, d.qkv, cols, rows, p.keys) != 0) {
    delete P;
    return -1;
  }
  P->grid = std::min(p.groups, num_sms());
  *out = P;
  return 0;
}

void attention_release(AttnPrepared* p) { delete p; }

int launch_attention_umma(const AttnPrepared* P, cudaStream_t st) {
  static bool attr_set = false;
  if (!attr_set) {
    DMC_CUDA_OK(cudaFuncSetAttribute(attention_umma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     static_cast<int>(AT_SMEM)));
    attr_set = true;
  }
  attention_umma_kernel<<<P->grid, 192, AT_SMEM, st>>>(P->tmQ, P->tmKV, P->p);
  DMC_CUDA_OK(cudaGetLastError());
  return 0;
}

}  // namespace dmc
