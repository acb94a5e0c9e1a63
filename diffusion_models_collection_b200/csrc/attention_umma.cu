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
// Pipeline of one CTA (320 threads), two q-tiles ("slots") in flight:
//   warp 8 lane 0 : TMA producer  (Q tile {64 x 128} per slot, K and V tiles {64 x keys}, SWIZZLE_128B)
//   warp 9 lane 0 : tcgen05.mma issuer:  S[128 x keys] = Q K^T  (A, B K-major)
//                                        O[128 x 64]   = P V    (A = P from smem, B = V MN-major)
//   warps 0..3 / 4..7 : softmax of slot 0 / 1: tcgen05.ld S row (thread = query row), max / exp2 / sum in fp32, P as
//                   bf16 into shared memory in the K-major SWIZZLE_128B operand layout; then O * (1/sum) -> bf16.
#include "common.cuh"
#include "kernels.h"

#include <stdlib.h>

#include <new>

namespace dmc {

constexpr int AT_M = 128;
constexpr int AT_HD = 64;
constexpr int AT_MAXKEYS = 256;
constexpr int AT_Q_BYTES = AT_M * AT_HD * 2;             // 16 KB per slot
constexpr int AT_KV_BYTES = AT_MAXKEYS * AT_HD * 2;      // 32 KB each (K, V): one 256-key tile or two 128-key tiles
constexpr int AT_P_BYTES = AT_M * AT_MAXKEYS * 2;        // 64 KB per slot
constexpr int AT_THREADS = 320;                          // warps 0-3 softmax slot 0, 4-7 softmax slot 1, 8 TMA, 9 MMA
constexpr size_t AT_SMEM = 2 * AT_Q_BYTES + 2 * AT_KV_BYTES + 2 * AT_P_BYTES + 1024 + 256;

struct AttnParams {
  int L, heads, C;
  int keys;            // key rows per q-tile: max(L, 128)
  int tiles;           // q-tiles (128 token rows x one head) in total
  int items;           // work items = pairs of q-tiles (slot 0, slot 1): ceil(tiles / 2)
  int total_rows;      // B * L
  float scale_log2e;
  __nv_bfloat16* out;
  int debug;           // DMC_ATTN_DEBUG timing switches of the ping-pong kernel (results are WRONG when set; profiling only)
};

struct AttnPrepared {
  CUtensorMap tmQ, tmKV, tmO;
  AttnParams p;
  int grid;
  int pp;   // L == 256: 2 = P in tensor memory (attention_ts_kernel), 1 = ping-pong kernel with P in shared memory (DMC_ATTN_PP=1),
            // 0 = the one-warpgroup-per-tile kernel (DMC_ATTN_PP=0); the older forms stay for A/B runs
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

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// q-tile t -> (head, first token row of its queries, first token row of its keys)
__device__ __forceinline__ void tile_coords(const AttnParams& p, int t, int& h, int& q_row0, int& key_row0) {
  if (p.L >= AT_M) {
    const int qpi = p.L / AT_M;  // q-tiles per (image, head); the two tiles of a pair share K / V when qpi == 2
    const int qt = t % qpi;
    const int g = t / qpi;
    h = g % p.heads;
    key_row0 = (g / p.heads) * p.L;
    q_row0 = key_row0 + qt * AT_M;
  } else {
    h = t % p.heads;
    key_row0 = (t / p.heads) * AT_M;
    q_row0 = key_row0;
  }
}

// Two q-tiles are in flight per CTA (slot 0 / slot 1, one softmax warpgroup each): while one warpgroup runs its
// exponentials the tensor core computes the other slot's S = Q K^T or O = P V, and the TMA warp prefetches the next
// pair's Q / K / V.  TMEM: slot s owns columns [256 s, 256 s + 256): S first, then (S is dead once P is in shared
// memory) O in its first 64 columns.
template <bool SMALL>
__global__ void __launch_bounds__(AT_THREADS, 1)
attention_umma_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmKV,
                      const __grid_constant__ AttnParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_align1024(smem_raw);
  uint8_t* sQ = smem;                      // 2 x 16 KB
  uint8_t* sK = sQ + 2 * AT_Q_BYTES;       // 32 KB
  uint8_t* sV = sK + AT_KV_BYTES;          // 32 KB
  uint8_t* sP = sV + AT_KV_BYTES;          // 2 x 64 KB
  uint64_t* bars = reinterpret_cast<uint64_t*>(sP + 2 * AT_P_BYTES);
  uint64_t* k_full = bars + 0;
  uint64_t* k_empty = bars + 1;
  uint64_t* v_full = bars + 2;
  uint64_t* v_empty = bars + 3;
  uint64_t* q_full = bars + 4;    // [2]
  uint64_t* q_empty = bars + 6;   // [2]
  uint64_t* s_full = bars + 8;    // [2]
  uint64_t* s_empty = bars + 10;  // [2]
  uint64_t* p_full = bars + 12;   // [2]
  uint64_t* o_full = bars + 14;   // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 16);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 8 && lane == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmKV);
    mbar_init(k_full, 1);
    mbar_init(k_empty, 1);
    mbar_init(v_full, 1);
    mbar_init(v_empty, 1);
    for (int s = 0; s < 2; ++s) {
      mbar_init(&q_full[s], 1);
      mbar_init(&q_empty[s], 1);
      mbar_init(&s_full[s], 1);
      mbar_init(&s_empty[s], 128);
      mbar_init(&p_full[s], 128);
      mbar_init(&o_full[s], 1);
    }
    mbar_fence_init();
  }
  if (warp == 9) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int keys = p.keys;
  const bool shared_kv = (p.L == 2 * AT_M);            // both slots of an item read the same 256-key K / V tiles
  const int kv_slot_bytes = shared_kv ? 0 : AT_KV_BYTES / 2;

  if (warp == 8) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      uint32_t it = 0;
      for (int item = blockIdx.x; item < p.items; item += gridDim.x, ++it) {
        const uint32_t ph = it & 1u;
        const int t0 = 2 * item;
        const bool two = t0 + 1 < p.tiles;
        int h[2], qr[2], kr[2];
        tile_coords(p, t0, h[0], qr[0], kr[0]);
        tile_coords(p, two ? t0 + 1 : t0, h[1], qr[1], kr[1]);
        const int nload = (shared_kv || !two) ? 1 : 2;
        const uint32_t kv_bytes = static_cast<uint32_t>(nload * keys * AT_HD * 2);
        mbar_wait(k_empty, ph ^ 1u);
        mbar_expect_tx(k_full, kv_bytes);
        for (int s = 0; s < nload; ++s) tma_load_2d(sK + s * kv_slot_bytes, &tmKV, k_full, p.C + h[s] * AT_HD, kr[s]);
        for (int s = 0; s < (two ? 2 : 1); ++s) {
          mbar_wait(&q_empty[s], ph ^ 1u);
          mbar_expect_tx(&q_full[s], AT_Q_BYTES);
          tma_load_2d(sQ + s * AT_Q_BYTES, &tmQ, &q_full[s], h[s] * AT_HD, qr[s]);
        }
        mbar_wait(v_empty, ph ^ 1u);
        mbar_expect_tx(v_full, kv_bytes);
        for (int s = 0; s < nload; ++s) tma_load_2d(sV + s * kv_slot_bytes, &tmKV, v_full, 2 * p.C + h[s] * AT_HD, kr[s]);
      }
    }
  } else if (warp == 9) {
    // ===================== MMA issuer =====================
    // Issue order (software-pipelined so that the two slots run HALF A PERIOD APART: while one slot's warpgroup runs its
    // exponentials the tensor core works for the other slot -- issued in lock-step, both warpgroups computed at the same
    // time and the tensor core idled meanwhile, then the warpgroups idled during the MMAs):
    //   S0(0);  for every item i:  S1(i);  PV0(i);  S0(i+1);  PV1(i)
    if (lane == 0) {
      const uint32_t idesc_s = umma_idesc_bf16(AT_M, keys);
      const uint32_t idesc_o = umma_idesc_bf16(AT_M, AT_HD, /*b_mn_major=*/1);
      auto issue_s = [&](int s, uint32_t ph) {  // S[s] = Q[s] K^T of the item whose phase is ph
        mbar_wait(&q_full[s], ph);
        mbar_wait(&s_empty[s], ph ^ 1u);
        tc_fence_after();
        const uint64_t qdesc = umma_desc_k_sw128(smem_u32(sQ + s * AT_Q_BYTES));
        const uint64_t kdesc = umma_desc_k_sw128(smem_u32(sK + s * kv_slot_bytes));
#pragma unroll
        for (int k = 0; k < AT_HD / 16; ++k)
          umma_bf16(tmem_base + s * 256, qdesc + 2 * k, kdesc + 2 * k, idesc_s, k != 0 ? 1u : 0u);
        umma_commit(&q_empty[s]);
        umma_commit(&s_full[s]);
      };
      auto issue_pv = [&](int s, uint32_t ph) {  // O[s] = P[s] V
        mbar_wait(&p_full[s], ph);
        tc_fence_after();
        const uint32_t pbase = smem_u32(sP + s * AT_P_BYTES);
        const uint32_t vbase = smem_u32(sV + s * kv_slot_bytes);
        for (int j = 0; j < keys / 16; ++j) {
          const uint64_t pdesc = umma_desc_k_sw128(pbase + (j >> 2) * (AT_M * 128)) + 2 * (j & 3);
          const uint64_t vdesc = umma_desc_mn_sw128(vbase + j * 16 * 128);
          umma_bf16(tmem_base + s * 256, pdesc, vdesc, idesc_o, j != 0 ? 1u : 0u);
        }
        umma_commit(&o_full[s]);
      };
      uint32_t it = 0;
      int item = blockIdx.x;
      if (item < p.items) {
        mbar_wait(k_full, 0u);
        issue_s(0, 0u);
      }
      for (; item < p.items; item += gridDim.x, ++it) {
        const uint32_t ph = it & 1u;
        const bool two = 2 * item + 1 < p.tiles;
        const int next = item + gridDim.x;
        if (two) issue_s(1, ph);
        umma_commit(k_empty);  // K(i) has been consumed by S0(i) and S1(i): the producer may load K(i+1)
        mbar_wait(v_full, ph);
        issue_pv(0, ph);
        if (next < p.items) {
          mbar_wait(k_full, ph ^ 1u);
          issue_s(0, ph ^ 1u);
        }
        if (two) issue_pv(1, ph);
        umma_commit(v_empty);
      }
    }
  } else {
    // ===================== softmax + output: warpgroup = slot, thread = query row =====================
    const int slot = warp >> 2;
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const uint32_t lane_addr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + slot * 256;
    uint8_t* sPs = sP + slot * AT_P_BYTES;
    // valid key columns of this row when several images share the tile (block diagonal), and their warp-uniform hull
    // (tcgen05.ld is warp-collective: whole 32-column chunks may only be skipped by all lanes together)
    const int c_lo = SMALL ? (row / p.L) * p.L : 0;
    const int c_hi = SMALL ? c_lo + p.L : keys;
    const int w_lo = SMALL ? ((q * 32) / p.L) * p.L : 0;
    const int w_hi = SMALL ? ((q * 32 + 31) / p.L + 1) * p.L : keys;
    uint32_t it = 0;
    for (int item = blockIdx.x; item < p.items; item += gridDim.x, ++it) {
      const uint32_t ph = it & 1u;
      const int t = 2 * item + slot;
      if (t >= p.tiles) continue;  // odd tail: slot 1 has no tile in the last item (warpgroup-uniform)
      int h, q_row0, key_row0;
      tile_coords(p, t, h, q_row0, key_row0);
      mbar_wait(&s_full[slot], ph);
      tc_fence_after();
      float m = -INFINITY, sum = 0.f;
      float ms;
      if (!SMALL) {
        // L = 128 / 256: every chunk is live.  TMEM loads are software-pipelined -- the load of chunk c+1 is in flight
        // while chunk c is processed (tcgen05.wait::ld waits for all outstanding loads, so exactly one is kept in
        // flight); with two warps per scheduler the ~150-clock load latency of 18 chunks per tile was mostly exposed.
        uint32_t ra[32], rb[32];
        // pass 1: row max
        tmem_ld_32x32(lane_addr, ra);
#pragma unroll 1
        for (int c0 = 0; c0 < keys; c0 += 64) {
          tmem_ld_wait();
          tmem_ld_32x32(lane_addr + c0 + 32, rb);
#pragma unroll
          for (int j = 0; j < 32; ++j) m = fmaxf(m, __uint_as_float(ra[j]));
          tmem_ld_wait();
          tmem_ld_32x32(lane_addr + ((c0 + 64 < keys) ? c0 + 64 : 0), ra);  // last iteration: chunk 0 again, for pass 2
#pragma unroll
          for (int j = 0; j < 32; ++j) m = fmaxf(m, __uint_as_float(rb[j]));
        }
        ms = m * p.scale_log2e;
        // pass 2: p = 2^(s * scale - max * scale), row sum, bf16 P into the K-major SWIZZLE_128B operand layout
        auto emit = [&](const uint32_t (&r)[32], int c0) {
          uint32_t pk[16];
#pragma unroll
          for (int j = 0; j < 32; j += 2) {
            const float e0 = ex2_approx(fmaf(__uint_as_float(r[j]), p.scale_log2e, -ms));
            const float e1 = ex2_approx(fmaf(__uint_as_float(r[j + 1]), p.scale_log2e, -ms));
            sum += e0 + e1;
            pk[j >> 1] = pack_bf16x2(e0, e1);
          }
          uint8_t* sub = sPs + (c0 >> 6) * (AT_M * 128) + row * 128;
          const int chunk0 = (c0 & 63) >> 3;  // first 16-byte chunk of these 32 keys inside the 128-byte row
#pragma unroll
          for (int q4 = 0; q4 < 4; ++q4) {
            const int chunk = (chunk0 + q4) ^ (row & 7);
            *reinterpret_cast<uint4*>(sub + chunk * 16) = make_uint4(pk[4 * q4], pk[4 * q4 + 1], pk[4 * q4 + 2], pk[4 * q4 + 3]);
          }
        };
#pragma unroll 1
        for (int c0 = 0; c0 < keys; c0 += 64) {
          tmem_ld_wait();
          tmem_ld_32x32(lane_addr + c0 + 32, rb);
          emit(ra, c0);
          tmem_ld_wait();
          if (c0 + 64 < keys) tmem_ld_32x32(lane_addr + c0 + 64, ra);
          emit(rb, c0 + 32);
        }
      } else {
      // pass 1: row max
      for (int c0 = w_lo & ~31; c0 < w_hi; c0 += 32) {
        uint32_t r[32];
        tmem_ld_32x32(lane_addr + c0, r);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const int c = c0 + j;
          if (c >= c_lo && c < c_hi) m = fmaxf(m, __uint_as_float(r[j]));
        }
      }
      ms = m * p.scale_log2e;
      // pass 2: p = 2^(s * scale - max * scale), row sum, bf16 P into the K-major SWIZZLE_128B operand layout
      for (int c0 = 0; c0 < keys; c0 += 32) {
        uint32_t pk[16];
        if (c0 + 32 <= w_lo || c0 >= w_hi) {  // chunk outside every row's block: probabilities are zero
#pragma unroll
          for (int j = 0; j < 16; ++j) pk[j] = 0u;
        } else {
          uint32_t r[32];
          tmem_ld_32x32(lane_addr + c0, r);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; j += 2) {
            float e0 = ex2_approx(fmaf(__uint_as_float(r[j]), p.scale_log2e, -ms));
            float e1 = ex2_approx(fmaf(__uint_as_float(r[j + 1]), p.scale_log2e, -ms));
            const int c = c0 + j;
            e0 = (c >= c_lo && c < c_hi) ? e0 : 0.f;
            e1 = (c + 1 >= c_lo && c + 1 < c_hi) ? e1 : 0.f;
            sum += e0 + e1;
            pk[j >> 1] = pack_bf16x2(e0, e1);
          }
        }
        uint8_t* sub = sPs + (c0 >> 6) * (AT_M * 128) + row * 128;
        const int chunk0 = (c0 & 63) >> 3;  // first 16-byte chunk of these 32 keys inside the 128-byte row
#pragma unroll
        for (int q4 = 0; q4 < 4; ++q4) {
          const int chunk = (chunk0 + q4) ^ (row & 7);
          *reinterpret_cast<uint4*>(sub + chunk * 16) = make_uint4(pk[4 * q4], pk[4 * q4 + 1], pk[4 * q4 + 2], pk[4 * q4 + 3]);
        }
      }
      }
      fence_proxy_async_smem();
      tc_fence_before();
      mbar_arrive(&p_full[slot]);
      // O = P V done -> normalise, store
      mbar_wait(&o_full[slot], ph);
      tc_fence_after();
      const float inv = 1.0f / sum;
      const int grow = q_row0 + row;
      __nv_bfloat16* op = p.out + static_cast<size_t>(grow) * p.C + h * AT_HD;
      uint32_t ro[2][32];
      tmem_ld_32x32(lane_addr, ro[0]);
      tmem_ld_32x32(lane_addr + 32, ro[1]);
      tmem_ld_wait();
#pragma unroll
      for (int c0 = 0; c0 < AT_HD; c0 += 32) {
        const uint32_t (&r)[32] = ro[c0 >> 5];
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
      mbar_arrive(&s_empty[slot]);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 9) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// ------------------------------------------------------------------------------------------------
// Long sequences (L a multiple of 256, L > 256: the shipped DiT config has 64 x 64 images = 1024 tokens): the same
// roles and buffers, flash-style over 256-key tiles.  A work item is the pair of 128-query tiles (slot 0 / 1) of one
// (image, head); for every key tile:  S = Q K_t^T  ->  running max m, P = 2^(S - m), l = l a + rowsum(P)  ->
// O_t = P V_t on the tensor core  ->  the thread that owns the row keeps the running output in REGISTERS,
// O = O a + O_t  (a = 2^(m_old - m_new); no TMEM read-modify-write).  Q stays in shared memory for the whole item.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(AT_THREADS, 1)
attention_long_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmKV,
                      const __grid_constant__ AttnParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_align1024(smem_raw);
  uint8_t* sQ = smem;
  uint8_t* sK = sQ + 2 * AT_Q_BYTES;
  uint8_t* sV = sK + AT_KV_BYTES;
  uint8_t* sP = sV + AT_KV_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sP + 2 * AT_P_BYTES);
  uint64_t* k_full = bars + 0;
  uint64_t* k_empty = bars + 1;
  uint64_t* v_full = bars + 2;
  uint64_t* v_empty = bars + 3;
  uint64_t* q_full = bars + 4;    // [2]
  uint64_t* q_empty = bars + 6;   // [2]
  uint64_t* s_full = bars + 8;    // [2]
  uint64_t* s_empty = bars + 10;  // [2]
  uint64_t* p_full = bars + 12;   // [2]
  uint64_t* o_full = bars + 14;   // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 16);
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  constexpr int KT = AT_MAXKEYS;  // keys per tile
  const int nkt = p.L / KT;

  if (warp == 8 && lane == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmKV);
    mbar_init(k_full, 1);
    mbar_init(k_empty, 1);
    mbar_init(v_full, 1);
    mbar_init(v_empty, 1);
    for (int s = 0; s < 2; ++s) {
      mbar_init(&q_full[s], 1);
      mbar_init(&q_empty[s], 1);
      mbar_init(&s_full[s], 1);
      mbar_init(&s_empty[s], 128);
      mbar_init(&p_full[s], 128);
      mbar_init(&o_full[s], 1);
    }
    mbar_fence_init();
  }
  if (warp == 9) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 8) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      uint32_t itq = 0, u = 0;
      for (int item = blockIdx.x; item < p.items; item += gridDim.x, ++itq) {
        int h, qr, kr;
        tile_coords(p, 2 * item, h, qr, kr);  // both q-tiles of the item belong to the same (image, head)
        for (int s = 0; s < 2; ++s) {
          mbar_wait(&q_empty[s], (itq & 1u) ^ 1u);
          mbar_expect_tx(&q_full[s], AT_Q_BYTES);
          tma_load_2d(sQ + s * AT_Q_BYTES, &tmQ, &q_full[s], h * AT_HD, qr + s * AT_M);
        }
        for (int kt = 0; kt < nkt; ++kt, ++u) {
          const uint32_t ph = u & 1u;
          mbar_wait(k_empty, ph ^ 1u);
          mbar_expect_tx(k_full, AT_KV_BYTES);
          tma_load_2d(sK, &tmKV, k_full, p.C + h * AT_HD, kr + kt * KT);
          mbar_wait(v_empty, ph ^ 1u);
          mbar_expect_tx(v_full, AT_KV_BYTES);
          tma_load_2d(sV, &tmKV, v_full, 2 * p.C + h * AT_HD, kr + kt * KT);
        }
      }
    }
  } else if (warp == 9) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      const uint32_t idesc_s = umma_idesc_bf16(AT_M, KT);
      const uint32_t idesc_o = umma_idesc_bf16(AT_M, AT_HD, /*b_mn_major=*/1);
      uint32_t itq = 0, u = 0;
      for (int item = blockIdx.x; item < p.items; item += gridDim.x, ++itq) {
        for (int kt = 0; kt < nkt; ++kt, ++u) {
          const uint32_t ph = u & 1u;
          mbar_wait(k_full, ph);
          for (int s = 0; s < 2; ++s) {
            if (kt == 0) mbar_wait(&q_full[s], itq & 1u);
            mbar_wait(&s_empty[s], ph ^ 1u);
            tc_fence_after();
            const uint64_t qdesc = umma_desc_k_sw128(smem_u32(sQ + s * AT_Q_BYTES));
            const uint64_t kdesc = umma_desc_k_sw128(smem_u32(sK));
#pragma unroll
            for (int k = 0; k < AT_HD / 16; ++k)
              umma_bf16(tmem_base + s * 256, qdesc + 2 * k, kdesc + 2 * k, idesc_s, k != 0 ? 1u : 0u);
            if (kt == nkt - 1) umma_commit(&q_empty[s]);  // Q is free once its last S product has been read
            umma_commit(&s_full[s]);
          }
          umma_commit(k_empty);
          mbar_wait(v_full, ph);
          for (int s = 0; s < 2; ++s) {
            mbar_wait(&p_full[s], ph);
            tc_fence_after();
            const uint32_t pbase = smem_u32(sP + s * AT_P_BYTES);
            const uint32_t vbase = smem_u32(sV);
            for (int j = 0; j < KT / 16; ++j) {
              const uint64_t pdesc = umma_desc_k_sw128(pbase + (j >> 2) * (AT_M * 128)) + 2 * (j & 3);
              const uint64_t vdesc = umma_desc_mn_sw128(vbase + j * 16 * 128);
              umma_bf16(tmem_base + s * 256, pdesc, vdesc, idesc_o, j != 0 ? 1u : 0u);
            }
            umma_commit(&o_full[s]);
          }
          umma_commit(v_empty);
        }
      }
    }
  } else {
    // ===================== softmax + running output: warpgroup = slot, thread = query row =====================
    const int slot = warp >> 2;
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const uint32_t lane_addr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + slot * 256;
    uint8_t* sPs = sP + slot * AT_P_BYTES;
    uint32_t u = 0;
    for (int item = blockIdx.x; item < p.items; item += gridDim.x) {
      int h, q_row0, key_row0;
      tile_coords(p, 2 * item + slot, h, q_row0, key_row0);
      float m = -INFINITY, l = 0.f;
      float o[AT_HD];
#pragma unroll
      for (int d = 0; d < AT_HD; ++d) o[d] = 0.f;
      for (int kt = 0; kt < nkt; ++kt, ++u) {
        const uint32_t ph = u & 1u;
        mbar_wait(&s_full[slot], ph);
        tc_fence_after();
        float mt = -INFINITY;
        for (int c0 = 0; c0 < KT; c0 += 32) {
          uint32_t r[32];
          tmem_ld_32x32(lane_addr + c0, r);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; ++j) mt = fmaxf(mt, __uint_as_float(r[j]));
        }
        const float m_new = fmaxf(m, mt);
        const float alpha = ex2_approx((m - m_new) * p.scale_log2e);  // first tile: 2^(-inf) = 0
        const float ms = m_new * p.scale_log2e;
        float sum = 0.f;
        for (int c0 = 0; c0 < KT; c0 += 32) {
          uint32_t r[32], pk[16];
          tmem_ld_32x32(lane_addr + c0, r);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; j += 2) {
            const float e0 = ex2_approx(fmaf(__uint_as_float(r[j]), p.scale_log2e, -ms));
            const float e1 = ex2_approx(fmaf(__uint_as_float(r[j + 1]), p.scale_log2e, -ms));
            sum += e0 + e1;
            pk[j >> 1] = pack_bf16x2(e0, e1);
          }
          uint8_t* sub = sPs + (c0 >> 6) * (AT_M * 128) + row * 128;
          const int chunk0 = (c0 & 63) >> 3;
#pragma unroll
          for (int q4 = 0; q4 < 4; ++q4) {
            const int chunk = (chunk0 + q4) ^ (row & 7);
            *reinterpret_cast<uint4*>(sub + chunk * 16) = make_uint4(pk[4 * q4], pk[4 * q4 + 1], pk[4 * q4 + 2], pk[4 * q4 + 3]);
          }
        }
        l = fmaf(l, alpha, sum);
        m = m_new;
        fence_proxy_async_smem();
        tc_fence_before();
        mbar_arrive(&p_full[slot]);
        // O_t = P V_t done: fold it into the running output
        mbar_wait(&o_full[slot], ph);
        tc_fence_after();
#pragma unroll
        for (int c0 = 0; c0 < AT_HD; c0 += 32) {
          uint32_t r[32];
          tmem_ld_32x32(lane_addr + c0, r);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; ++j) o[c0 + j] = fmaf(o[c0 + j], alpha, __uint_as_float(r[j]));
        }
        tc_fence_before();
        mbar_arrive(&s_empty[slot]);
      }
      const float inv = 1.0f / l;
      const int grow = q_row0 + row;
      if (grow < p.total_rows) {
        uint4* op = reinterpret_cast<uint4*>(p.out + static_cast<size_t>(grow) * p.C + h * AT_HD);
#pragma unroll
        for (int v = 0; v < AT_HD / 8; ++v) {
          uint4 w;
          w.x = pack_bf16x2(o[8 * v] * inv, o[8 * v + 1] * inv);
          w.y = pack_bf16x2(o[8 * v + 2] * inv, o[8 * v + 3] * inv);
          w.z = pack_bf16x2(o[8 * v + 4] * inv, o[8 * v + 5] * inv);
          w.w = pack_bf16x2(o[8 * v + 6] * inv, o[8 * v + 7] * inv);
          op[v] = w;
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 9) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// ------------------------------------------------------------------------------------------------
// L = 256 (the UNet's 16 x 16 blocks, every DiT-32 layer): the "ping-pong" form.  Measured on the kernel above (run 13,
// 2048 images): 9 400 clocks per (image, head) item against 4 096 clocks of exponentials at 16 lanes / clk / SM -- one
// warpgroup per q-tile leaves ONE warp per scheduler on each tile's serial chain (S -> max -> exp -> P -> PV -> O -> next S)
// and the O read-out sits on that chain because O aliases S in tensor memory.  Here
//   * all EIGHT softmax warps work on ONE q-tile at a time (thread = one query row x 128 of its 256 keys; the two threads of
//     a row exchange their maxima through shared memory around one 256-thread named barrier), so a tile's softmax takes half
//     as long and the two tiles of an item alternate: while the warps run tile B's exponentials the tensor core does tile A's
//     P V, then A's next S = Q K^T;
//   * FOUR further warps own the output: they read O, scale by 1 / rowsum, store, and hand the accumulator back
//     (s_empty) -- none of that is on the softmax warps' path any more;
//   * the partial row sums travel in TENSOR MEMORY (one dead S column per half, tcgen05.st), not shared memory: the 224 KB of
//     Q / K / V / P leave no room for them.
// 448 threads: warps 0-7 softmax, 8-11 output, 12 TMA producer, 13 MMA issuer (+ TMEM allocation).
// ------------------------------------------------------------------------------------------------
constexpr int PP_THREADS = 448;
constexpr int PP_DATA_BYTES = 2 * AT_Q_BYTES + 2 * AT_KV_BYTES + 2 * AT_P_BYTES;   // 224 KB
constexpr int PP_MX_BYTES = 2 * 2 * AT_M * 4;                                     // [slot][half][row] row maxima
constexpr int PP_NEED = PP_DATA_BYTES + PP_MX_BYTES + 256;                        // + barriers
constexpr size_t PP_SMEM = 232448;                                                // the whole opt-in window (227 KB)

__device__ __forceinline__ void tmem_st_1(uint32_t taddr, uint32_t v) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x1.b32 [%0], {%1};" ::"r"(taddr), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t tmem_ld_1(uint32_t taddr) {
  uint32_t v;
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(v) : "r"(taddr) : "memory");
  return v;
}

__global__ void __launch_bounds__(PP_THREADS, 1)
attention_pp_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmKV,
                    const __grid_constant__ AttnParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_align1024(smem_raw);
  if (threadIdx.x == 0 && (smem - smem_raw) + PP_NEED > static_cast<long long>(PP_SMEM)) {
    printf("dmc: attention_pp_kernel: dynamic shared memory base is not 1024-byte aligned enough (%d bytes lost)\n",
           static_cast<int>(smem - smem_raw));
    __trap();
  }
  uint8_t* sQ = smem;                      // 2 x 16 KB (tile A = rows 0..127, tile B = rows 128..255 of the image)
  uint8_t* sK = sQ + 2 * AT_Q_BYTES;       // 32 KB
  uint8_t* sV = sK + AT_KV_BYTES;          // 32 KB
  uint8_t* sP = sV + AT_KV_BYTES;          // 2 x 64 KB
  float* s_mx = reinterpret_cast<float*>(sP + 2 * AT_P_BYTES);   // [2 slots][2 halves][128 rows]
  uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(s_mx) + PP_MX_BYTES);
  uint64_t* k_full = bars + 0;
  uint64_t* k_empty = bars + 1;
  uint64_t* v_full = bars + 2;
  uint64_t* v_empty = bars + 3;
  uint64_t* q_full = bars + 4;    // [2]
  uint64_t* q_empty = bars + 6;   // [2]
  uint64_t* s_full = bars + 8;    // [2]
  uint64_t* s_empty = bars + 10;  // [2]  128 arrivals: the output warps have read O and the row sums
  uint64_t* p_full = bars + 12;   // [2]  256 arrivals: P is in shared memory, the row sums are in tensor memory
  uint64_t* o_full = bars + 14;   // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 16);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 12 && lane == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmKV);
    mbar_init(k_full, 1);
    mbar_init(k_empty, 1);
    mbar_init(v_full, 1);
    mbar_init(v_empty, 1);
    for (int s = 0; s < 2; ++s) {
      mbar_init(&q_full[s], 1);
      mbar_init(&q_empty[s], 1);
      mbar_init(&s_full[s], 1);
      mbar_init(&s_empty[s], 128);
      mbar_init(&p_full[s], 256);
      mbar_init(&o_full[s], 1);
    }
    mbar_fence_init();
  }
  if (warp == 13) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  constexpr int L = AT_MAXKEYS;   // 256 keys = 2 q-tiles per (image, head)

  if (warp == 12) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      uint32_t it = 0;
      for (int item = blockIdx.x; item < p.items; item += gridDim.x, ++it) {
        const uint32_t ph = it & 1u;
        const int h = item % p.heads;
        const int row0 = (item / p.heads) * L;
        mbar_wait(k_empty, ph ^ 1u);
        mbar_expect_tx(k_full, AT_KV_BYTES);
        tma_load_2d(sK, &tmKV, k_full, p.C + h * AT_HD, row0);
        for (int s = 0; s < 2; ++s) {
          mbar_wait(&q_empty[s], ph ^ 1u);
          mbar_expect_tx(&q_full[s], AT_Q_BYTES);
          tma_load_2d(sQ + s * AT_Q_BYTES, &tmQ, &q_full[s], h * AT_HD, row0 + s * AT_M);
        }
        mbar_wait(v_empty, ph ^ 1u);
        mbar_expect_tx(v_full, AT_KV_BYTES);
        tma_load_2d(sV, &tmKV, v_full, 2 * p.C + h * AT_HD, row0);
      }
    }
  } else if (warp == 13) {
    // ===================== MMA issuer =====================
    // order:  S_A(0) S_B(0);  per item i:  PV_A(i)  S_A(i+1)  PV_B(i)  S_B(i+1)
    if (lane == 0) {
      const uint32_t idesc_s = umma_idesc_bf16(AT_M, L);
      const uint32_t idesc_o = umma_idesc_bf16(AT_M, AT_HD, /*b_mn_major=*/1);
      const uint64_t kdesc = umma_desc_k_sw128(smem_u32(sK));
      auto issue_s = [&](int s) {
        const uint64_t qdesc = umma_desc_k_sw128(smem_u32(sQ + s * AT_Q_BYTES));
#pragma unroll
        for (int k = 0; k < AT_HD / 16; ++k)
          umma_bf16(tmem_base + s * 256, qdesc + 2 * k, kdesc + 2 * k, idesc_s, k != 0 ? 1u : 0u);
        umma_commit(&q_empty[s]);
        umma_commit(&s_full[s]);
      };
      uint32_t it = 0;
      int item = blockIdx.x;
      if (item < p.items) {
        mbar_wait(k_full, 0u);
        for (int s = 0; s < 2; ++s) {
          mbar_wait(&q_full[s], 0u);
          tc_fence_after();
          issue_s(s);
        }
        umma_commit(k_empty);
      }
      for (; item < p.items; item += gridDim.x, ++it) {
        const uint32_t ph = it & 1u;
        const bool next = item + static_cast<int>(gridDim.x) < p.items;
        for (int s = 0; s < 2; ++s) {
          mbar_wait(&p_full[s], ph);
          if (s == 0) mbar_wait(v_full, ph);
          tc_fence_after();
          const uint32_t pbase = smem_u32(sP + s * AT_P_BYTES);
          const uint32_t vbase = smem_u32(sV);
#pragma unroll 4
          for (int j = 0; j < L / 16; ++j) {
            const uint64_t pdesc = umma_desc_k_sw128(pbase + (j >> 2) * (AT_M * 128)) + 2 * (j & 3);
            const uint64_t vdesc = umma_desc_mn_sw128(vbase + j * 16 * 128);
            umma_bf16(tmem_base + s * 256, pdesc, vdesc, idesc_o, j != 0 ? 1u : 0u);
          }
          umma_commit(&o_full[s]);
          if (s == 1) umma_commit(v_empty);
          if (next) {
            if (s == 0) mbar_wait(k_full, ph ^ 1u);
            mbar_wait(&q_full[s], ph ^ 1u);
            mbar_wait(&s_empty[s], ph);   // the output warps have read O(i) and the row sums of this slot
            tc_fence_after();
            issue_s(s);
            if (s == 1) umma_commit(k_empty);
          }
        }
      }
    }
  } else if (warp < 8) {
    // ===================== softmax: 8 warps on one q-tile; thread = (query row, half of the keys) =====================
    const int q = warp & 3;
    const int half = warp >> 2;
    const int row = q * 32 + lane;
    const int dbg = p.debug;
    uint32_t it = 0;
    for (int item = blockIdx.x; item < p.items; item += gridDim.x, ++it) {
      const uint32_t ph = it & 1u;
#pragma unroll 1
      for (int slot = 0; slot < 2; ++slot) {
        const uint32_t la = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + slot * 256 + half * 128;
        uint8_t* sPs = sP + slot * AT_P_BYTES + (half * 2) * (AT_M * 128) + row * 128;
        float* mx = s_mx + slot * 2 * AT_M;
        mbar_wait(&s_full[slot], ph);
        tc_fence_after();
        uint32_t ra[32], rb[32];
        float m = -INFINITY;
        // pass 1: maximum of this thread's 128 scores (one TMEM load kept in flight)
        if (!(dbg & 1)) {
        tmem_ld_32x32(la, ra);
        tmem_ld_wait();
        tmem_ld_32x32(la + 32, rb);
#pragma unroll
        for (int j = 0; j < 32; ++j) m = fmaxf(m, __uint_as_float(ra[j]));
        tmem_ld_wait();
        tmem_ld_32x32(la + 64, ra);
#pragma unroll
        for (int j = 0; j < 32; ++j) m = fmaxf(m, __uint_as_float(rb[j]));
        tmem_ld_wait();
        tmem_ld_32x32(la + 96, rb);
#pragma unroll
        for (int j = 0; j < 32; ++j) m = fmaxf(m, __uint_as_float(ra[j]));
        tmem_ld_wait();
        tmem_ld_32x32(la, ra);  // chunk 0 again, for pass 2
#pragma unroll
        for (int j = 0; j < 32; ++j) m = fmaxf(m, __uint_as_float(rb[j]));
        } else {
          m = 30.f;
          tmem_ld_32x32(la, ra);
        }
        if (!(dbg & 16)) {
        mx[half * AT_M + row] = m;
        named_bar_sync(1, 256);
        m = fmaxf(m, mx[(half ^ 1) * AT_M + row]);
        }
        const float ms = m * p.scale_log2e;
        float sum = 0.f;
        // pass 2: p = 2^(s * scale - max * scale), partial row sum, bf16 P into the K-major SWIZZLE_128B operand layout
        auto emit = [&](const uint32_t (&r)[32], int c0) {
          uint32_t pk[16];
#pragma unroll
          for (int j = 0; j < 32; j += 2) {
            float e0 = fmaf(__uint_as_float(r[j]), p.scale_log2e, -ms);
            float e1 = fmaf(__uint_as_float(r[j + 1]), p.scale_log2e, -ms);
            if (!(dbg & 4)) {
              e0 = ex2_approx(e0);
              e1 = ex2_approx(e1);
            }
            if (!(dbg & 64)) sum += e0 + e1;
            pk[j >> 1] = pack_bf16x2(e0, e1);
          }
          uint8_t* sub = sPs + (c0 >> 6) * (AT_M * 128);
          const int chunk0 = (c0 & 63) >> 3;
          if (dbg & 2) {
            if (pk[0] == 0x12345678u && pk[7] == pk[15]) *reinterpret_cast<uint4*>(sub) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
          } else
#pragma unroll
          for (int q4 = 0; q4 < 4; ++q4) {
            const int chunk = (chunk0 + q4) ^ (row & 7);
            *reinterpret_cast<uint4*>(sub + chunk * 16) = make_uint4(pk[4 * q4], pk[4 * q4 + 1], pk[4 * q4 + 2], pk[4 * q4 + 3]);
          }
        };
        tmem_ld_wait();
        tmem_ld_32x32(la + 32, rb);
        emit(ra, 0);
        tmem_ld_wait();
        tmem_ld_32x32(la + 64, ra);
        emit(rb, 32);
        tmem_ld_wait();
        tmem_ld_32x32(la + 96, rb);
        emit(ra, 64);
        tmem_ld_wait();
        emit(rb, 96);
        // the partial row sum goes to a dead S column of this thread's own half (64 / 192: outside O's columns 0..63)
        if (!(dbg & 8)) {
          tmem_st_1(la + 64, __float_as_uint(sum));
          tmem_st_wait();
        }
        fence_proxy_async_smem();
        tc_fence_before();
        mbar_arrive(&p_full[slot]);
      }
    }
  } else {
    // ===================== output warps: O * (1 / rowsum) -> bf16 -> global; hand the accumulator back =====================
    const int q = warp & 3;
    const int row = q * 32 + lane;
    uint32_t it = 0;
    for (int item = blockIdx.x; item < p.items; item += gridDim.x, ++it) {
      const uint32_t ph = it & 1u;
      const int h = item % p.heads;
      const int row0 = (item / p.heads) * L;
#pragma unroll 1
      for (int slot = 0; slot < 2; ++slot) {
        const uint32_t la = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + slot * 256;
        mbar_wait(&o_full[slot], ph);
        tc_fence_after();
        uint32_t ro[2][32];
        tmem_ld_32x32(la, ro[0]);
        tmem_ld_32x32(la + 32, ro[1]);
        const uint32_t s0 = tmem_ld_1(la + 64);
        const uint32_t s1 = tmem_ld_1(la + 192);
        tmem_ld_wait();
        tc_fence_before();
        mbar_arrive(&s_empty[slot]);
        const float inv = 1.0f / (__uint_as_float(s0) + __uint_as_float(s1));
        const int grow = row0 + slot * AT_M + row;
        __nv_bfloat16* op = p.out + static_cast<size_t>(grow) * p.C + h * AT_HD;
        if (p.debug & 32) continue;
#pragma unroll
        for (int c0 = 0; c0 < AT_HD; c0 += 32) {
          const uint32_t (&r)[32] = ro[c0 >> 5];
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
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 13) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// ------------------------------------------------------------------------------------------------
// L = 256, third form: P NEVER LEAVES TENSOR MEMORY.  What the timing switches of the ping-pong kernel showed (run 16, 2048 images,
// tools/bench_attention.py): 9 600 clocks per (image, head) item; without the exponentials 9 500; with the softmax warps doing
// nothing at all 7 700 -- the kernel is not bound by the SFU or the tensor pipe but by its LOADS: 128 FLOP per HBM byte is half the
// chip's ridge (254), K and V were single-buffered (each load could only be issued ~4 000 clocks before its first use, about the
// HBM latency under load), and 128 of the 227 KB of shared memory held the two P tiles.  Here
//   * the softmax warps write P (bf16 pairs) back into the S columns they have just read (tcgen05.st) and O = P V is a
//     tcgen05.mma with the A operand IN TENSOR MEMORY -- no P in shared memory, no generic-proxy fence, 4 KB less shared-memory
//     traffic per MMA;
//   * the 128 KB that frees double-buffer Q, K and V: the producer runs a whole item ahead;
//   * the output warps stage their 32 x 64 block in shared memory and store it with TMA (full 128-byte rows) instead of 8
//     row-strided 16-byte stores per thread.
// TMEM columns of slot s (256 s + ...): S [0, 256) -> P keys 0..127 in [0, 64), O in [64, 128), P keys 128..255 in [128, 192).
// ------------------------------------------------------------------------------------------------
constexpr int TS_SMEM_Q = 2 * 2 * AT_Q_BYTES;          // [stage][slot]
constexpr int TS_SMEM_KV = 2 * AT_KV_BYTES;            // [stage], K and V each
constexpr int TS_SMEM_O = 4 * 4096;                    // one 32-row x 64-channel staging block per output warp
constexpr int TS_SMEM_F = 2 * 2 * 2 * AT_M * 4;        // row maxima + partial row sums: [slot][half][row] each
constexpr int TS_NEED = TS_SMEM_Q + 2 * TS_SMEM_KV + TS_SMEM_O + TS_SMEM_F + 256;
constexpr size_t TS_SMEM = TS_NEED + 1024;

__device__ __forceinline__ void tmem_st_16(uint32_t taddr, const uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(
          taddr),
      "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]), "r"(v[10]),
      "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
      : "memory");
}
// D[tmem] (+)= A[tmem] * B[smem desc]: the A operand (128 rows = lanes, 16 bf16 = 8 columns per K step) is read from tensor memory
__device__ __forceinline__ void umma_bf16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(tmem_d),
      "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* m, const void* src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(
                   reinterpret_cast<uint64_t>(m)),
               "r"(smem_u32(src)), "r"(c0), "r"(c1)
               : "memory");
}

__global__ void __launch_bounds__(PP_THREADS, 1)
attention_ts_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmKV,
                    const __grid_constant__ CUtensorMap tmO, const __grid_constant__ AttnParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_align1024(smem_raw);
  uint8_t* sQ = smem;                          // [stage][slot] 16 KB
  uint8_t* sK = sQ + TS_SMEM_Q;                // [stage] 32 KB
  uint8_t* sV = sK + TS_SMEM_KV;               // [stage] 32 KB
  uint8_t* sO = sV + TS_SMEM_KV;               // [output warp] 4 KB
  float* s_mx = reinterpret_cast<float*>(sO + TS_SMEM_O);   // [slot][half][row]
  float* s_sum = s_mx + 2 * 2 * AT_M;                       // [slot][half][row]
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_sum + 2 * 2 * AT_M);
  uint64_t* k_full = bars + 0;    // [2 stages]
  uint64_t* k_empty = bars + 2;   // [2]
  uint64_t* v_full = bars + 4;    // [2]
  uint64_t* v_empty = bars + 6;   // [2]
  uint64_t* q_full = bars + 8;    // [stage][slot]
  uint64_t* q_empty = bars + 12;  // [stage][slot]
  uint64_t* s_full = bars + 16;   // [slot]
  uint64_t* s_empty = bars + 18;  // [slot] 128 arrivals: the output warps have read O
  uint64_t* p_full = bars + 20;   // [slot] 256 arrivals: P is in tensor memory, the partial row sums in shared memory
  uint64_t* o_full = bars + 22;   // [slot]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 24);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 12 && lane == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmKV);
    tma_prefetch_desc(&tmO);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&k_full[i], 1);
      mbar_init(&k_empty[i], 1);
      mbar_init(&v_full[i], 1);
      mbar_init(&v_empty[i], 1);
      mbar_init(&s_full[i], 1);
      mbar_init(&s_empty[i], 128);
      mbar_init(&p_full[i], 256);
      mbar_init(&o_full[i], 1);
    }
    for (int i = 0; i < 4; ++i) {
      mbar_init(&q_full[i], 1);
      mbar_init(&q_empty[i], 1);
    }
    mbar_fence_init();
  }
  if (warp == 13) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  constexpr int L = AT_MAXKEYS;

  if (warp == 12) {
    // ===================== TMA producer: a whole item ahead of the MMAs =====================
    if (lane == 0) {
      uint32_t it = 0;
      for (int item = blockIdx.x; item < p.items; item += gridDim.x, ++it) {
        const int st = it & 1u;
        const uint32_t ph = (it >> 1) & 1u;
        const int h = item % p.heads;
        const int row0 = (item / p.heads) * L;
        mbar_wait(&k_empty[st], ph ^ 1u);
        mbar_expect_tx(&k_full[st], AT_KV_BYTES);
        tma_load_2d(sK + st * AT_KV_BYTES, &tmKV, &k_full[st], p.C + h * AT_HD, row0);
        for (int s = 0; s < 2; ++s) {
          mbar_wait(&q_empty[st * 2 + s], ph ^ 1u);
          mbar_expect_tx(&q_full[st * 2 + s], AT_Q_BYTES);
          tma_load_2d(sQ + (st * 2 + s) * AT_Q_BYTES, &tmQ, &q_full[st * 2 + s], h * AT_HD, row0 + s * AT_M);
        }
        mbar_wait(&v_empty[st], ph ^ 1u);
        mbar_expect_tx(&v_full[st], AT_KV_BYTES);
        tma_load_2d(sV + st * AT_KV_BYTES, &tmKV, &v_full[st], 2 * p.C + h * AT_HD, row0);
      }
    }
  } else if (warp == 13) {
    // ===================== MMA issuer:  S_A(0) S_B(0);  per item i:  PV_A(i)  S_A(i+1)  PV_B(i)  S_B(i+1) =====================
    if (lane == 0) {
      const uint32_t idesc_s = umma_idesc_bf16(AT_M, L);
      const uint32_t idesc_o = umma_idesc_bf16(AT_M, AT_HD, /*b_mn_major=*/1);
      auto issue_s = [&](int s, uint32_t jt) {  // S[s] = Q[s] K^T of the CTA's jt-th item
        const int st = jt & 1u;
        const uint32_t ph = (jt >> 1) & 1u;
        if (s == 0) mbar_wait(&k_full[st], ph);
        mbar_wait(&q_full[st * 2 + s], ph);
        tc_fence_after();
        const uint64_t qdesc = umma_desc_k_sw128(smem_u32(sQ + (st * 2 + s) * AT_Q_BYTES));
        const uint64_t kdesc = umma_desc_k_sw128(smem_u32(sK + st * AT_KV_BYTES));
#pragma unroll
        for (int k = 0; k < AT_HD / 16; ++k)
          umma_bf16(tmem_base + s * 256, qdesc + 2 * k, kdesc + 2 * k, idesc_s, k != 0 ? 1u : 0u);
        umma_commit(&q_empty[st * 2 + s]);
        umma_commit(&s_full[s]);
        if (s == 1) umma_commit(&k_empty[st]);
      };
      uint32_t it = 0;
      int item = blockIdx.x;
      if (item < p.items) {
        issue_s(0, 0u);
        issue_s(1, 0u);
      }
      for (; item < p.items; item += gridDim.x, ++it) {
        const int st = it & 1u;
        const uint32_t ph = it & 1u;            // per-slot barriers complete once per item
        const uint32_t phs = (it >> 1) & 1u;    // per-stage barriers complete once per two items
        const bool next = item + static_cast<int>(gridDim.x) < p.items;
        for (int s = 0; s < 2; ++s) {
          mbar_wait(&p_full[s], ph);
          if (s == 0) mbar_wait(&v_full[st], phs);
          tc_fence_after();
          const uint32_t vbase = smem_u32(sV + st * AT_KV_BYTES);
          const uint32_t tslot = tmem_base + s * 256;
#pragma unroll 4
          for (int j = 0; j < L / 16; ++j) {
            const uint64_t vdesc = umma_desc_mn_sw128(vbase + j * 16 * 128);
            const uint32_t pa = tslot + (j < 8 ? 8 * j : 128 + 8 * (j - 8));
            umma_bf16_ts(tslot + 64, pa, vdesc, idesc_o, j != 0 ? 1u : 0u);
          }
          umma_commit(&o_full[s]);
          if (s == 1) umma_commit(&v_empty[st]);
          if (next) {
            mbar_wait(&s_empty[s], ph);   // the output warps have read O(i) of this slot
            issue_s(s, it + 1);
          }
        }
      }
    }
  } else if (warp < 8) {
    // ===================== softmax: 8 warps on one q-tile; thread = (query row, half of the keys) =====================
    const int q = warp & 3;
    const int half = warp >> 2;
    const int row = q * 32 + lane;
    const int dbg = p.debug;   // DMC_ATTN_DEBUG timing switches (wrong results): 1 no row max, 2 no P stores, 4 no exponentials,
                               // 16 no maximum exchange, 32 no output stores
    uint32_t it = 0;
    for (int item = blockIdx.x; item < p.items; item += gridDim.x, ++it) {
      const uint32_t ph = it & 1u;
#pragma unroll 1
      for (int slot = 0; slot < 2; ++slot) {
        const uint32_t la = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + slot * 256 + half * 128;
        float* mx = s_mx + slot * 2 * AT_M;
        mbar_wait(&s_full[slot], ph);
        tc_fence_after();
        uint32_t ra[32], rb[32];
        float m = -INFINITY;
        if (!(dbg & 1)) {
        tmem_ld_32x32(la, ra);
        tmem_ld_wait();
        tmem_ld_32x32(la + 32, rb);
#pragma unroll
        for (int j = 0; j < 32; ++j) m = fmaxf(m, __uint_as_float(ra[j]));
        tmem_ld_wait();
        tmem_ld_32x32(la + 64, ra);
#pragma unroll
        for (int j = 0; j < 32; ++j) m = fmaxf(m, __uint_as_float(rb[j]));
        tmem_ld_wait();
        tmem_ld_32x32(la + 96, rb);
#pragma unroll
        for (int j = 0; j < 32; ++j) m = fmaxf(m, __uint_as_float(ra[j]));
        tmem_ld_wait();
        tmem_ld_32x32(la, ra);  // chunk 0 again, for pass 2
#pragma unroll
        for (int j = 0; j < 32; ++j) m = fmaxf(m, __uint_as_float(rb[j]));
        } else {
          m = 30.f;
          tmem_ld_32x32(la, ra);
        }
        if (!(dbg & 16)) {
        mx[half * AT_M + row] = m;
        named_bar_sync(1, 256);
        m = fmaxf(m, mx[(half ^ 1) * AT_M + row]);
        }
        const float ms = m * p.scale_log2e;
        float sum = 0.f;
        // pass 2: p = 2^(s * scale - max * scale) as bf16 pairs, written over the S columns this thread has already consumed
        auto emit = [&](const uint32_t (&r)[32], int c) {
          uint32_t pk[16];
#pragma unroll
          for (int j = 0; j < 32; j += 2) {
            float e0 = fmaf(__uint_as_float(r[j]), p.scale_log2e, -ms);
            float e1 = fmaf(__uint_as_float(r[j + 1]), p.scale_log2e, -ms);
            if (!(dbg & 4)) {
              e0 = ex2_approx(e0);
              e1 = ex2_approx(e1);
            }
            sum += e0 + e1;
            pk[j >> 1] = pack_bf16x2(e0, e1);
          }
          if (!(dbg & 2) || pk[3] == 0x12345678u) tmem_st_16(la + 16 * c, pk);
        };
        tmem_ld_wait();
        tmem_ld_32x32(la + 32, rb);
        emit(ra, 0);
        tmem_ld_wait();
        tmem_ld_32x32(la + 64, ra);
        emit(rb, 1);
        tmem_ld_wait();
        tmem_ld_32x32(la + 96, rb);
        emit(ra, 2);
        tmem_ld_wait();
        emit(rb, 3);
        s_sum[(slot * 2 + half) * AT_M + row] = sum;
        tmem_st_wait();
        tc_fence_before();
        mbar_arrive(&p_full[slot]);
      }
    }
  } else {
    // ===================== output warps: O * (1 / rowsum) -> bf16 -> shared memory -> TMA store =====================
    const int q = warp & 3;
    const int row = q * 32 + lane;
    uint8_t* stg = sO + q * 4096;
    const uint32_t xr = static_cast<uint32_t>(lane & 7) << 4;
    uint32_t it = 0;
    for (int item = blockIdx.x; item < p.items; item += gridDim.x, ++it) {
      const uint32_t ph = it & 1u;
      const int h = item % p.heads;
      const int row0 = (item / p.heads) * L;
#pragma unroll 1
      for (int slot = 0; slot < 2; ++slot) {
        const uint32_t la = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + slot * 256 + 64;
        mbar_wait(&p_full[slot], ph);   // (acquires the softmax warps' partial row sums)
        mbar_wait(&o_full[slot], ph);
        tc_fence_after();
        uint32_t r[32];
        uint32_t pk[32];
        tmem_ld_32x32(la, r);
        const float inv = 1.0f / (s_sum[(slot * 2) * AT_M + row] + s_sum[(slot * 2 + 1) * AT_M + row]);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 16; ++j) pk[j] = pack_bf16x2(__uint_as_float(r[2 * j]) * inv, __uint_as_float(r[2 * j + 1]) * inv);
        tmem_ld_32x32(la + 32, r);
        tmem_ld_wait();
        tc_fence_before();
        mbar_arrive(&s_empty[slot]);
#pragma unroll
        for (int j = 0; j < 16; ++j) pk[16 + j] = pack_bf16x2(__uint_as_float(r[2 * j]) * inv, __uint_as_float(r[2 * j + 1]) * inv);
        if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");  // the previous store has left the buffer
        __syncwarp();
        uint8_t* rowp = stg + lane * 128;
#pragma unroll
        for (int c = 0; c < 8; ++c)
          *reinterpret_cast<uint4*>(rowp + ((static_cast<uint32_t>(c) << 4) ^ xr)) =
              make_uint4(pk[4 * c], pk[4 * c + 1], pk[4 * c + 2], pk[4 * c + 3]);
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0 && !(p.debug & 32)) {
          tma_store_2d(&tmO, stg, h * AT_HD, row0 + slot * AT_M + q * 32);
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
      }
    }
    if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 13) {
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
  if (L >= AT_M) return L == 128 || L == 256 || (L > AT_MAXKEYS && L % AT_MAXKEYS == 0);
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
  p.keys = d.L < AT_M ? AT_M : (d.L > AT_MAXKEYS ? AT_MAXKEYS : d.L);
  p.total_rows = d.B * d.L;
  p.tiles = ((p.total_rows + AT_M - 1) / AT_M) * d.heads;
  p.items = (p.tiles + 1) / 2;
  p.scale_log2e = 1.4426950408889634f / sqrtf(static_cast<float>(AT_HD));
  p.out = reinterpret_cast<__nv_bfloat16*>(d.out);
  const uint64_t cols = 3ull * d.C, rows = static_cast<uint64_t>(p.total_rows);
  if (encode2d(&P->tmQ, d.qkv, cols, rows, AT_M) != 0 || encode2d(&P->tmKV, d.qkv, cols, rows, p.keys) != 0) {
    delete P;
    return -1;
  }
  const char* e = getenv("DMC_ATTN_PP");
  P->pp = d.L == AT_MAXKEYS ? ((e && (e[0] == '0' || e[0] == '1')) ? e[0] - '0' : 2) : 0;
  if (P->pp == 2 && encode2d(&P->tmO, d.out, static_cast<uint64_t>(d.C), rows, 32) != 0) {
    delete P;
    return -1;
  }
  if (P->pp) p.items = p.tiles / 2;   // one item = one (image, head): both q-tiles, shared K / V
  const char* dbg = getenv("DMC_ATTN_DEBUG");
  p.debug = dbg ? atoi(dbg) : 0;
  P->grid = std::min(p.items, num_sms());
  *out = P;
  return 0;
}

void attention_release(AttnPrepared* p) { delete p; }

int launch_attention_umma(const AttnPrepared* P, cudaStream_t st) {
  static DeviceOnce attr_set;
  int attr_set_dev = 0;
  if (attr_set.need(&attr_set_dev)) {
    DMC_CUDA_OK(cudaFuncSetAttribute(attention_umma_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     static_cast<int>(AT_SMEM)));
    DMC_CUDA_OK(cudaFuncSetAttribute(attention_umma_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     static_cast<int>(AT_SMEM)));
    DMC_CUDA_OK(cudaFuncSetAttribute(attention_long_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     static_cast<int>(AT_SMEM)));
    DMC_CUDA_OK(cudaFuncSetAttribute(attention_pp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     static_cast<int>(PP_SMEM)));
    DMC_CUDA_OK(cudaFuncSetAttribute(attention_ts_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     static_cast<int>(TS_SMEM)));
    attr_set.done(attr_set_dev);
  }
  if (P->pp == 2) {
    attention_ts_kernel<<<P->grid, PP_THREADS, TS_SMEM, st>>>(P->tmQ, P->tmKV, P->tmO, P->p);
    DMC_CUDA_OK(cudaGetLastError());
    return 0;
  }
  if (P->pp) {
    attention_pp_kernel<<<P->grid, PP_THREADS, PP_SMEM, st>>>(P->tmQ, P->tmKV, P->p);
    DMC_CUDA_OK(cudaGetLastError());
    return 0;
  }
  if (P->p.L > AT_MAXKEYS) {
    attention_long_kernel<<<P->grid, AT_THREADS, AT_SMEM, st>>>(P->tmQ, P->tmKV, P->p);
    DMC_CUDA_OK(cudaGetLastError());
    return 0;
  }
  if (P->p.L < AT_M) attention_umma_kernel<true><<<P->grid, AT_THREADS, AT_SMEM, st>>>(P->tmQ, P->tmKV, P->p);
  else attention_umma_kernel<false><<<P->grid, AT_THREADS, AT_SMEM, st>>>(P->tmQ, P->tmKV, P->p);
  DMC_CUDA_OK(cudaGetLastError());
  return 0;
}

}  // namespace dmc
