// Convolution as an implicit GEMM on the 5th-generation tensor cores (tcgen05.mma, accumulators in TMEM),
// operands staged by TMA.  Replaces nn.Conv2d 3x3 / 1x1 of /root/reference/models/unet.py:37,54,58,81,82,106,116,240
// (and the nn.Linear GEMMs of models/dit.py:94-109) with bias + conditioning + residual (+ fused 1x1 shortcut as
// extra K columns) in the epilogue.
//
// GEMM view:  D[M = B*Hout*Wout pixels, N = Cout] = A[M, K] * W[N, K]^T,  K = sum_src taps * C_src.
//   * A is never materialised (no im2col buffer): for K-block (source s, tap (dh,dw), 64-channel chunk c) the
//     128 x 64 bf16 A tile is ONE tiled 4-D TMA box {64 ch, BW, BH, BNIMG} over the NHWC tensor, shifted by the tap
//     offset; TMA zero-fills the out-of-image halo (= conv padding) and strided (stride-2) convs use the tensor
//     map's element strides.  The box lands in shared memory as 128 rows x 128 B with the 128-byte swizzle,
//     i.e. exactly the K-major SWIZZLE_128B operand layout tcgen05.mma consumes.
//   * W is a bf16 [Cout_pad, Ktot] matrix (K ordered like the K-blocks), loaded as {64, BN} boxes.
//   * One CTA per SM, persistent over output tiles (MT x 128 pixels x BN channels), warp-specialised:
//       warp 0 lane 0 : TMA producer          (smem ring of NST stages, full/empty mbarriers)
//       warp 1 lane 0 : tcgen05.mma issuer    (UMMA 128 x BN x 16, fp32 accumulate, 2 TMEM accumulator stages)
//       warp 2        : TMEM allocator
//       warps 4..11   : epilogue: tcgen05.ld -> +bias +cond +residual -> bf16 NHWC (or fp32 NCHW for the model
//                       head) -> GroupNorm partial sums, overlapped with the next tile's MMAs.
#include "common.cuh"
#include "kernels.h"

#include <stdlib.h>

#include <new>

namespace dmc {

constexpr int TILE_M = 128;
constexpr int KB = 64;  // K elements per block = 128 bytes of bf16 = one swizzle row
constexpr int A_STAGE_BYTES = TILE_M * KB * 2;

struct ConvKParams {
  int nseg;
  int seg_taps[3];
  int seg_chunks[3];
  int seg_kb_end[3];  // cumulative K-block count
  signed char dh[3][9];
  signed char dw[3][9];
  int stride;
  int BW, BH, BNIMG;
  int tiles_w, tiles_h;
  int num_m_tiles, num_n_tiles, num_kb;
  int B, Hout, Wout;     // iteration space (output pixels per image = Hout*Wout)
  int out_H, out_W;      // stored output tensor spatial dims
  int oscale, ooff_h, ooff_w;
  int Cout;
  const float* bias;
  const float* cond;
  int cond_stride;
  const __nv_bfloat16* residual;
  __nv_bfloat16* out;
  float* out_nchw;
  float* stats;
  int stats_slots, stats_slot_base;
  // transformer epilogues (DiT)
  int act;                    // 1: GELU
  const float* gate;          // per-image per-channel multiplier
  int gate_stride;
  const float* residual_f32;  // fp32 NHWC residual stream
  float* out_f32;             // fp32 NHWC output
  int unpatch_p;              // > 0: out_nchw columns are (pi, qi, c) patch entries
  // TMA-store epilogue: every epilogue warp stores its 32 rows x 64 channels as one box {64, qbw, qbh, qbn}
  int tma_store;
  int qbw, qbh;               // quarter box: qbw pixels x qbh rows x 32/(qbw*qbh) images
};

struct ConvPrepared {
  CUtensorMap tmA[3];
  CUtensorMap tmB;
  CUtensorMap tmOut;
  ConvKParams kp;
  int BN, MT, CG;
  int grid;
  size_t smem;
};

constexpr int CONV_THREADS = 384;   // warp 0 TMA, 1 MMA, 2 TMEM allocator, 3 idle, 4..11 epilogue
constexpr int EPI_THREADS = 256;

template <int BN, int MT, int CG>
struct ConvCfg {
  static constexpr int A_BYTES = MT * A_STAGE_BYTES;
  static constexpr int B_STAGE_BYTES = (BN / CG) * KB * 2;  // a CTA pair splits the weight tile: N/2 rows each
  static constexpr int STAGE_BYTES = A_BYTES + B_STAGE_BYTES;
  // epilogue staging for TMA stores: one 32-row x 64-channel bf16 box (4 KB, SWIZZLE_128B) per epilogue warp
  static constexpr bool TMA_STORE = BN >= 128;
  static constexpr int STORE_BYTES = TMA_STORE ? (EPI_THREADS / 32) * 4096 : 0;
  static constexpr int RING_BYTES = 216 * 1024 - STORE_BYTES;
  static constexpr int NST = RING_BYTES / STAGE_BYTES > 8 ? 8 : RING_BYTES / STAGE_BYTES;
  static constexpr int ACC_COLS = MT * BN;  // TMEM columns of one accumulator stage
  static constexpr int TMEM_COLS = (2 * ACC_COLS < 32) ? 32 : 2 * ACC_COLS;
  static constexpr size_t SMEM = static_cast<size_t>(NST) * STAGE_BYTES + STORE_BYTES + 1024 /*align slack*/ + 256 /*barriers*/;
  static_assert(TMEM_COLS <= 512 && (TMEM_COLS & (TMEM_COLS - 1)) == 0, "TMEM allocation must be a power of two <= 512");
};

// Sums each of 8 per-lane values over the 32 lanes of the warp (full) or over each 16-lane half, with 9 (8) shuffles
// instead of 40: every round halves the number of values a lane carries.  On return `r` is the total of value `idx`.
__device__ __forceinline__ void reduce8(const float (&v)[8], bool full, int lane, float& r, int& idx) {
  float w[4];
  int base;
  if (full) {
    const bool hi = (lane & 16) != 0;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float recv = __shfl_xor_sync(0xFFFFFFFFu, hi ? v[i] : v[4 + i], 16);
      w[i] = (hi ? v[4 + i] : v[i]) + recv;
    }
    base = hi ? 4 : 0;
    const bool h8 = (lane & 8) != 0;
    float u[2];
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const float recv = __shfl_xor_sync(0xFFFFFFFFu, h8 ? w[i] : w[2 + i], 8);
      u[i] = (h8 ? w[2 + i] : w[i]) + recv;
    }
    base += h8 ? 2 : 0;
    const bool h4 = (lane & 4) != 0;
    const float recv = __shfl_xor_sync(0xFFFFFFFFu, h4 ? u[0] : u[1], 4);
    float t = (h4 ? u[1] : u[0]) + recv;
    base += h4 ? 1 : 0;
    t += __shfl_xor_sync(0xFFFFFFFFu, t, 2);
    t += __shfl_xor_sync(0xFFFFFFFFu, t, 1);
    r = t;
    idx = base;
  } else {
    const bool h8 = (lane & 8) != 0;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float recv = __shfl_xor_sync(0xFFFFFFFFu, h8 ? v[i] : v[4 + i], 8);
      w[i] = (h8 ? v[4 + i] : v[i]) + recv;
    }
    base = h8 ? 4 : 0;
    const bool h4 = (lane & 4) != 0;
    float u[2];
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const float recv = __shfl_xor_sync(0xFFFFFFFFu, h4 ? w[i] : w[2 + i], 4);
      u[i] = (h4 ? w[2 + i] : w[i]) + recv;
    }
    base += h4 ? 2 : 0;
    const bool h2 = (lane & 2) != 0;
    const float recv = __shfl_xor_sync(0xFFFFFFFFu, h2 ? u[0] : u[1], 2);
    float t = (h2 ? u[1] : u[0]) + recv;
    base += h2 ? 1 : 0;
    t += __shfl_xor_sync(0xFFFFFFFFu, t, 1);
    r = t;
    idx = base;
  }
}

template <int BN, int MT, int CG>
__global__ void __launch_bounds__(CONV_THREADS, 1)
conv_umma_kernel(const __grid_constant__ CUtensorMap tmA0, const __grid_constant__ CUtensorMap tmA1,
                 const __grid_constant__ CUtensorMap tmA2, const __grid_constant__ CUtensorMap tmB,
                 const __grid_constant__ CUtensorMap tmOut, const __grid_constant__ ConvKParams p) {
  using Cfg = ConvCfg<BN, MT, CG>;
  constexpr int NST = Cfg::NST;
  constexpr int MTG = MT * CG;  // 128-pixel tiles per CTA-group tile
  const int rank = (CG == 2) ? static_cast<int>(cluster_ctarank()) : 0;
  const int group = blockIdx.x / CG, num_groups = gridDim.x / CG;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint8_t* store_stage = smem + NST * Cfg::STAGE_BYTES;  // 1024-aligned: STAGE_BYTES is a multiple of 1024
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(store_stage + Cfg::STORE_BYTES);
  uint64_t* empty_bar = full_bar + NST;
  uint64_t* tfull_bar = empty_bar + NST;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int num_ct = (p.num_m_tiles + MTG - 1) / MTG;  // group tiles along M (MTG consecutive 128-pixel tiles each)
  const int num_tiles = num_ct * p.num_n_tiles;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA0);
    if (p.nseg > 1) tma_prefetch_desc(&tmA1);
    if (p.nseg > 2) tma_prefetch_desc(&tmA2);
    tma_prefetch_desc(&tmB);
    if (Cfg::TMA_STORE && p.tma_store) tma_prefetch_desc(&tmOut);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < NST; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull_bar[i], 1);
      mbar_init(&tempty_bar[i], CG * (EPI_THREADS / 32));  // one arrival per epilogue warp (of both CTAs of a pair)
    }
    mbar_fence_init();
  }
  if (warp == 2) {
    if (CG == 2) {
      tmem_alloc_cg2(tmem_slot, Cfg::TMEM_COLS);
      tmem_relinquish_cg2();
    } else {
      tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
      tmem_relinquish();
    }
  }
  tc_fence_before();
  __syncthreads();
  if (CG == 2) cluster_sync_all();  // the peer's barriers are initialised before anything signals them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      uint32_t it = 0;
      for (int tile = group; tile < num_tiles; tile += num_groups) {
        const int n_tile = tile % p.num_n_tiles;
        const int ct = tile / p.num_n_tiles;
        int w0[MT], h0[MT], n0[MT];
#pragma unroll
        for (int mt = 0; mt < MT; ++mt) {
          const int m_tile = ct * MTG + rank * MT + mt;  // may be past the end: its box is fully out of bounds -> zero fill
          w0[mt] = (m_tile % p.tiles_w) * p.BW * p.stride;
          h0[mt] = ((m_tile / p.tiles_w) % p.tiles_h) * p.BH * p.stride;
          n0[mt] = (m_tile / (p.tiles_w * p.tiles_h)) * p.BNIMG;
        }
        int seg = 0, kb_in_seg = 0;
        for (int kb = 0; kb < p.num_kb; ++kb, ++it) {
          const int stage = it % NST;
          const uint32_t phase = (it / NST) & 1u;
          mbar_wait(&empty_bar[stage], phase ^ 1u);
          // the leader's barrier counts the bytes of BOTH CTAs of a pair (the MMAs it issues read both)
          if (rank == 0) mbar_expect_tx(&full_bar[stage], CG * Cfg::STAGE_BYTES);
          uint8_t* sa = smem + stage * Cfg::STAGE_BYTES;
          uint8_t* sb = sa + Cfg::A_BYTES;
          const int chunks = p.seg_chunks[seg];
          const int tap = kb_in_seg / chunks, chunk = kb_in_seg % chunks;
          const CUtensorMap* tm = seg == 0 ? &tmA0 : (seg == 1 ? &tmA1 : &tmA2);
          if (CG == 2) {
            const uint32_t lbar = mapa_shared(smem_u32(&full_bar[stage]), 0);
#pragma unroll
            for (int mt = 0; mt < MT; ++mt)
              tma_load_4d_cg2(sa + mt * A_STAGE_BYTES, tm, lbar, chunk * KB, w0[mt] + p.dw[seg][tap],
                              h0[mt] + p.dh[seg][tap], n0[mt]);
            tma_load_2d_cg2(sb, &tmB, lbar, kb * KB, n_tile * BN + rank * (BN / 2));
          } else {
#pragma unroll
            for (int mt = 0; mt < MT; ++mt)
              tma_load_4d(sa + mt * A_STAGE_BYTES, tm, &full_bar[stage], chunk * KB, w0[mt] + p.dw[seg][tap],
                          h0[mt] + p.dh[seg][tap], n0[mt]);
            tma_load_2d(sb, &tmB, &full_bar[stage], kb * KB, n_tile * BN);
          }
          if (++kb_in_seg == p.seg_taps[seg] * chunks) {
            ++seg;
            kb_in_seg = 0;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0 && rank == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(TILE_M * CG, BN);
      uint32_t it = 0, local = 0;
      for (int tile = group; tile < num_tiles; tile += num_groups, ++local) {
        const uint32_t acc = local & 1u;
        const uint32_t acc_phase = (local >> 1) & 1u;
        mbar_wait(&tempty_bar[acc], acc_phase ^ 1u);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * Cfg::ACC_COLS;
        for (int kb = 0; kb < p.num_kb; ++kb, ++it) {
          const int stage = it % NST;
          const uint32_t phase = (it / NST) & 1u;
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + stage * Cfg::STAGE_BYTES);
          const uint64_t bdesc = umma_desc_k_sw128(sa + Cfg::A_BYTES);
#pragma unroll
          for (int mt = 0; mt < MT; ++mt) {
            const uint64_t adesc = umma_desc_k_sw128(sa + mt * A_STAGE_BYTES);
#pragma unroll
            for (int k = 0; k < KB / 16; ++k) {
              // advance 16 bf16 = 32 bytes inside the 128-byte swizzle row: +2 in the (addr >> 4) field
              if (CG == 2) umma_bf16_cg2(d_tmem + mt * BN, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
              else umma_bf16(d_tmem + mt * BN, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
            }
          }
          // frees the smem stage (in both CTAs of a pair) once these MMAs have read it
          if (CG == 2) umma_commit_cg2(&empty_bar[stage]);
          else umma_commit(&empty_bar[stage]);
        }
        // accumulators complete -> epilogue (of both CTAs)
        if (CG == 2) umma_commit_cg2(&tfull_bar[acc]);
        else umma_commit(&tfull_bar[acc]);
      }
    }
  } else if (warp >= 4) {
    // ===================== epilogue (8 warps) =====================
    // warp -> TMEM lane quarter q (= warp % 4, the hardware rule) and group grp: with MT == 2 the group is the
    // 128-pixel sub-tile, with MT == 1 it is the half of the BN output channels this warp converts.
    const int q = warp & 3;
    const int grp = (warp - 4) >> 2;
    const int row = q * 32 + lane;      // tile row == TMEM lane
    const int ppi = p.BW * p.BH;        // pixels per image inside one 128-pixel tile
    constexpr int COLS = (MT == 2) ? BN : (BN >= 64 ? BN / 2 : BN);
    const int col0 = (MT == 2) ? 0 : (BN >= 64 ? grp * COLS : 0);
    const bool idle = (MT == 1 && BN < 64 && grp == 1);
    const int wi = row % p.BW, hi = (row / p.BW) % p.BH, ni = row / ppi;
    uint32_t local = 0;
    for (int tile = group; tile < num_tiles; tile += num_groups, ++local) {
      const uint32_t acc = local & 1u;
      const uint32_t acc_phase = (local >> 1) & 1u;
      const int n_tile = tile % p.num_n_tiles;
      const int m_tile = (tile / p.num_n_tiles) * MTG + rank * MT + (MT == 2 ? grp : 0);
      const int tw = m_tile % p.tiles_w;
      const int th = (m_tile / p.tiles_w) % p.tiles_h;
      const int ti = m_tile / (p.tiles_w * p.tiles_h);
      const int n = ti * p.BNIMG + ni;
      const int oh = (th * p.BH + hi) * p.oscale + p.ooff_h;
      const int ow = (tw * p.BW + wi) * p.oscale + p.ooff_w;
      const bool valid = n < p.B;
      const size_t pix = (static_cast<size_t>(n) * p.out_H + oh) * p.out_W + ow;
      // TMA store: coordinates (iteration space) of the first row of this warp's 32-row quarter
      const bool use_tma = Cfg::TMA_STORE && p.tma_store != 0;
      const int r0 = q * 32;
      const int sw0 = tw * p.BW + r0 % p.BW, sh0 = th * p.BH + (r0 / p.BW) % p.BH, sn0 = ti * p.BNIMG + r0 / ppi;
      uint8_t* my_stage = store_stage + (warp - 4) * 4096;

      mbar_wait(&tfull_bar[acc], acc_phase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * Cfg::ACC_COLS +
                             (MT == 2 ? grp * BN : 0) + col0;
      if (!idle) {
#pragma unroll 1
        for (int c0 = 0; c0 < COLS; c0 += 32) {
          uint32_t r[32];
          tmem_ld_32x32(taddr + c0, r);
          const int cg = n_tile * BN + col0 + c0;  // first global output channel of this chunk
          // per-channel addends (warp-uniform addresses -> L1 broadcast), fetched while the TMEM load is in flight
          float add[32];
          const bool real = cg < p.Cout;  // false only for padded weight rows (warp-uniform)
          if (p.out_nchw == nullptr && real) {
            if (p.bias != nullptr) {
              const float4* b4 = reinterpret_cast<const float4*>(p.bias + cg);
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                const float4 b = __ldg(b4 + j);
                add[4 * j] = b.x; add[4 * j + 1] = b.y; add[4 * j + 2] = b.z; add[4 * j + 3] = b.w;
              }
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j) add[j] = 0.f;
            }
            if (p.cond != nullptr && valid) {
              const float4* c4 = reinterpret_cast<const float4*>(p.cond + static_cast<size_t>(n) * p.cond_stride + cg);
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                const float4 b = __ldg(c4 + j);
                add[4 * j] += b.x; add[4 * j + 1] += b.y; add[4 * j + 2] += b.z; add[4 * j + 3] += b.w;
              }
            }
          }
          uint4 res[4];
          const bool has_res = p.residual != nullptr && valid && real && p.out_nchw == nullptr;
          if (has_res) {
            const uint4* r4 = reinterpret_cast<const uint4*>(p.residual + pix * p.Cout + cg);
#pragma unroll
            for (int j = 0; j < 4; ++j) res[j] = __ldg(r4 + j);
          }
          tmem_ld_wait();
          if (!real) continue;
          float v[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
          if (p.out_nchw != nullptr) {
            // model head: few real channels, fp32 NCHW, coalesced along W across the warp
            if (valid) {
              const int up = p.unpatch_p;
              const int oc = up > 0 ? p.Cout / (up * up) : p.Cout;  // image channels
#pragma unroll
              for (int j = 0; j < 32; ++j) {
                const int c = cg + j;
                if (c < p.Cout) {
                  float o = v[j] + (p.bias ? __ldg(p.bias + c) : 0.f);
                  if (up > 0) {  // DiT.unpatchify: column = (pi * up + qi) * oc + ch
                    const int ch = c % oc, pq = c / oc;
                    const int yy = oh * up + pq / up, xx = ow * up + pq % up;
                    p.out_nchw[((static_cast<size_t>(n) * oc + ch) * (p.out_H * up) + yy) * (p.out_W * up) + xx] = o;
                  } else {
                    p.out_nchw[((static_cast<size_t>(n) * p.Cout + c) * p.out_H + oh) * p.out_W + ow] = o;
                  }
                }
              }
            }
            continue;
          }
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] += add[j];
          if (p.act == 1) {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = gelu_f(v[j]);
          }
          if (p.gate != nullptr && valid) {
            const float4* g4 = reinterpret_cast<const float4*>(p.gate + static_cast<size_t>(n) * p.gate_stride + cg);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float4 gv = __ldg(g4 + j);
              v[4 * j] *= gv.x; v[4 * j + 1] *= gv.y; v[4 * j + 2] *= gv.z; v[4 * j + 3] *= gv.w;
            }
          }
          if (p.residual_f32 != nullptr && valid) {
            const float4* r4 = reinterpret_cast<const float4*>(p.residual_f32 + pix * p.Cout + cg);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float4 rv = r4[j];  // plain load: out_f32 may alias the residual stream
              v[4 * j] += rv.x; v[4 * j + 1] += rv.y; v[4 * j + 2] += rv.z; v[4 * j + 3] += rv.w;
            }
          }
          if (p.out_f32 != nullptr && valid) {
            float4* o4 = reinterpret_cast<float4*>(p.out_f32 + pix * p.Cout + cg);
#pragma unroll
            for (int j = 0; j < 8; ++j) o4[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
          }
          if (has_res) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              uint32_t w[4] = {res[j].x, res[j].y, res[j].z, res[j].w};
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                float2 f = unpack_bf16x2(w[k]);
                v[8 * j + 2 * k] += f.x;
                v[8 * j + 2 * k + 1] += f.y;
              }
            }
          }
          if (use_tma) {
            // stage this 32-column half of a 64-channel box in shared memory (row = lane, 128-byte rows, 16-byte chunks
            // XOR-swizzled with the row index: conflict-free writes and the layout SWIZZLE_128B tensor maps expect)
            const int half = (c0 >> 5) & 1;
            if (half == 0) {  // the previous box of this warp must have been read out of the staging buffer
              if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
              __syncwarp();
            }
            uint8_t* rowp = my_stage + lane * 128;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              uint4 u;
              u.x = pack_bf16x2(v[8 * j], v[8 * j + 1]);
              u.y = pack_bf16x2(v[8 * j + 2], v[8 * j + 3]);
              u.z = pack_bf16x2(v[8 * j + 4], v[8 * j + 5]);
              u.w = pack_bf16x2(v[8 * j + 6], v[8 * j + 7]);
              *reinterpret_cast<uint4*>(rowp + (((half * 4 + j) ^ (lane & 7)) << 4)) = u;
            }
            if (half == 1) {
              fence_proxy_async_smem();
              __syncwarp();
              if (lane == 0) {
                tma_store_4d(&tmOut, my_stage, cg - 32, sw0, sh0, sn0);  // clipped at the tensor bounds (n >= B)
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
              }
            }
          } else if (valid && p.out != nullptr) {
            uint4* o4 = reinterpret_cast<uint4*>(p.out + pix * p.Cout + cg);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              uint4 u;
              u.x = pack_bf16x2(v[8 * j], v[8 * j + 1]);
              u.y = pack_bf16x2(v[8 * j + 2], v[8 * j + 3]);
              u.z = pack_bf16x2(v[8 * j + 4], v[8 * j + 5]);
              u.w = pack_bf16x2(v[8 * j + 6], v[8 * j + 7]);
              o4[j] = u;
            }
          }
          if (p.stats != nullptr) {
            // GroupNorm partial sums of the OUTPUT per (image, 8-channel block) over the rows of this warp that belong
            // to one image (all 32 when ppi >= 32, else each 16-lane half), stored in this warp's own slot: plain
            // stores, no atomics -> deterministic and batch-invariant; the consumer adds the slots in index order.
            float sv[8];
#pragma unroll
            for (int b = 0; b < 4; ++b) {
              float s = 0.f, ss = 0.f;
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                s += v[8 * b + j];
                ss = fmaf(v[8 * b + j], v[8 * b + j], ss);
              }
              sv[b] = valid ? s : 0.f;
              sv[4 + b] = valid ? ss : 0.f;
            }
            const bool full = ppi >= 32;
            float tot;
            int idx;
            reduce8(sv, full, lane, tot, idx);
            const bool writer = full ? ((lane & 3) == 0) : ((lane & 1) == 0);
            if (writer && valid) {
              const int wpi = ppi >> 5;  // epilogue warps per image inside one tile (0: an image is a half-warp)
              const int slot = p.stats_slot_base + (full ? ((th * p.tiles_w + tw) * wpi + (q % wpi)) : 0);
              float* dst = p.stats + ((static_cast<size_t>(n) * p.stats_slots + slot) * (p.Cout >> 3) + ((cg >> 3) + (idx & 3))) * 2;
              dst[idx >> 2] = tot;
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {  // one arrival per warp, on the leader's barrier (its MMA warp reuses the accumulator stage)
        if (CG == 2) mbar_arrive_cluster(mapa_shared(smem_u32(&tempty_bar[acc]), 0));
        else mbar_arrive(&tempty_bar[acc]);
      }
    }
    // all TMA stores of this warp have left shared memory and are complete before the CTA exits
    if (lane == 0 && Cfg::TMA_STORE && p.tma_store != 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }

  tc_fence_before();
  __syncthreads();
  if (CG == 2) cluster_sync_all();  // no CTA leaves (or frees TMEM) while its peer may still signal / read it
  if (warp == 2) {
    tc_fence_after();
    if (CG == 2) tmem_dealloc_cg2(tmem_base, Cfg::TMEM_COLS);
    else tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
  }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
static int encode_map(CUtensorMap* m, const void* base, int rank, const cuuint64_t* dims, const cuuint64_t* strides,
                      const cuuint32_t* box, const cuuint32_t* estr) {
  EncodeTiledFn fn = encode_tiled_fn();
  DMC_REQUIRE(fn != nullptr, "conv: cuTensorMapEncodeTiled unavailable -- call dmc_init() on a CUDA 12+ driver");
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, static_cast<cuuint32_t>(rank), const_cast<void*>(base), dims,
                  strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  DMC_REQUIRE(r == CUDA_SUCCESS, "conv: cuTensorMapEncodeTiled failed with CUresult %d", static_cast<int>(r));
  return 0;
}

// Tile configuration (BN output channels x MT 128-pixel sub-tiles per CTA tile).  Bytes staged per MAC fall with
// MT * BN (the A tile is reused by BN channels, the B tile by MT * 128 pixels): prefer 256 TMEM columns per
// accumulator when that still yields at least one tile per SM.
struct TileCfg { int bn, mt, cg; };
static TileCfg pick_cfg(int cout_pad, int m_tiles) {
  const int sms = num_sms();
  // DMC_CONV_CG (debug / tests): "1" never pairs SMs, "2" pairs them whenever the channel count allows
  const char* e = getenv("DMC_CONV_CG");
  const bool allow_pairs = !(e && e[0] == '1');
  const bool force_pairs = e && e[0] == '2';
  const TileCfg cands[7] = {{256, 1, 2}, {128, 2, 2}, {256, 1, 1}, {128, 2, 1}, {128, 1, 1}, {64, 1, 1}, {32, 1, 1}};
  if (force_pairs)
    for (int i = 0; i < 2; ++i)
      if (cout_pad % cands[i].bn == 0) return cands[i];
  for (const TileCfg& c : cands) {
    if (cout_pad % c.bn != 0 || (c.cg == 2 && !allow_pairs)) continue;
    const long long ctas = static_cast<long long>((m_tiles + c.mt * c.cg - 1) / (c.mt * c.cg)) * (cout_pad / c.bn) * c.cg;
    if (ctas >= sms) return c;
  }
  for (int i = 6; i >= 2; --i)
    if (cout_pad % cands[i].bn == 0) return cands[i];
  return {32, 1, 1};
}

int conv_prepare(const dmc_conv_desc& d, ConvPrepared** out) {
  DMC_REQUIRE(d.nsrc >= 1 && d.nsrc <= 3, "conv: nsrc=%d", d.nsrc);
  DMC_REQUIRE(d.stride == 1 || d.stride == 2, "conv: stride=%d", d.stride);
  DMC_REQUIRE(d.up_phase >= -1 && d.up_phase <= 3, "conv: up_phase=%d", d.up_phase);
  DMC_REQUIRE(d.B > 0 && d.Hin > 0 && d.Win > 0, "conv: empty input");
  DMC_REQUIRE(d.Hin % d.stride == 0 && d.Win % d.stride == 0, "conv: odd spatial size with stride 2");
  DMC_REQUIRE(d.weight && (d.out_bf16 || d.out_f32_nchw || d.out_f32_nhwc), "conv: null weight/output");
  DMC_REQUIRE(d.Cout_pad % 32 == 0 && d.Cout <= d.Cout_pad, "conv: Cout_pad=%d must be a multiple of 32", d.Cout_pad);
  if (d.out_bf16 || d.out_f32_nhwc)
    DMC_REQUIRE(d.Cout % 32 == 0, "conv: NHWC output needs Cout %% 32 == 0 (got %d)", d.Cout);
  DMC_REQUIRE(d.act == 0 || d.act == 1, "conv: act=%d", d.act);
  DMC_REQUIRE(!(d.residual && d.residual_f32), "conv: bf16 and fp32 residuals are exclusive");
  DMC_REQUIRE(d.unpatch_p >= 0 && (d.unpatch_p == 0 || (d.out_f32_nchw && d.up_phase < 0 && d.stride == 1 &&
                                                        d.Cout % (d.unpatch_p * d.unpatch_p) == 0)),
              "conv: unpatch_p=%d needs an fp32 NCHW output and Cout = p*p*channels", d.unpatch_p);
  if (d.out_f32_nchw) DMC_REQUIRE(!d.gate && !d.residual_f32 && !d.out_f32_nhwc && d.act == 0,
                                  "conv: the fp32 NCHW head takes bias only");
  if (d.stats) DMC_REQUIRE(d.out_bf16 != nullptr, "conv: stats need a bf16 output");

  ConvPrepared* P = new (std::nothrow) ConvPrepared();
  DMC_REQUIRE(P != nullptr, "conv: out of host memory");
  ConvKParams& kp = P->kp;
  memset(&kp, 0, sizeof(kp));
  kp.nseg = d.nsrc;
  kp.stride = d.stride;
  const int Hout = d.Hin / d.stride, Wout = d.Win / d.stride;
  kp.B = d.B; kp.Hout = Hout; kp.Wout = Wout;
  if (d.up_phase >= 0) {
    DMC_REQUIRE(d.stride == 1, "conv: upsample phases need stride 1");
    kp.oscale = 2; kp.ooff_h = d.up_phase >> 1; kp.ooff_w = d.up_phase & 1;
    kp.out_H = 2 * Hout; kp.out_W = 2 * Wout;
  } else {
    kp.oscale = 1; kp.ooff_h = kp.ooff_w = 0; kp.out_H = Hout; kp.out_W = Wout;
  }
  // tile box over output pixels: 128 = BW * BH * BNIMG
  int BW = std::min(Wout, TILE_M);
  DMC_REQUIRE(TILE_M % BW == 0 && Wout % BW == 0, "conv: Wout=%d unsupported", Wout);
  int BH = std::min(Hout, TILE_M / BW);
  DMC_REQUIRE((TILE_M / BW) % BH == 0 && Hout % BH == 0, "conv: Hout=%d unsupported", Hout);
  int BNIMG = TILE_M / (BW * BH);
  kp.BW = BW; kp.BH = BH; kp.BNIMG = BNIMG;
  kp.tiles_w = Wout / BW; kp.tiles_h = Hout / BH;
  const int img_tiles = (d.B + BNIMG - 1) / BNIMG;
  kp.num_m_tiles = img_tiles * kp.tiles_w * kp.tiles_h;

  int kb = 0;
  for (int s = 0; s < d.nsrc; ++s) {
    const int taps = d.src_taps[s], C = d.src_c[s];
    DMC_REQUIRE(d.src[s] != nullptr, "conv: source %d is null", s);
    DMC_REQUIRE(C > 0 && C % KB == 0, "conv: source %d has %d channels (need a multiple of 64)", s, C);
    DMC_REQUIRE(taps == 9 || taps == 1 || (taps == 4 && d.up_phase >= 0), "conv: source %d taps=%d", s, taps);
    kp.seg_taps[s] = taps;
    kp.seg_chunks[s] = C / KB;
    for (int t = 0; t < taps; ++t) {
      int dh = 0, dw = 0;
      if (taps == 9) { dh = t / 3 - 1; dw = t % 3 - 1; }
      else if (taps == 4) { dh = (d.up_phase >> 1) - 1 + t / 2; dw = (d.up_phase & 1) - 1 + t % 2; }
      kp.dh[s][t] = static_cast<signed char>(dh);
      kp.dw[s][t] = static_cast<signed char>(dw);
    }
    kb += taps * (C / KB);
    kp.seg_kb_end[s] = kb;
    cuuint64_t dims[4] = {static_cast<cuuint64_t>(C), static_cast<cuuint64_t>(d.Win), static_cast<cuuint64_t>(d.Hin),
                          static_cast<cuuint64_t>(d.B)};
    cuuint64_t strides[3] = {static_cast<cuuint64_t>(C) * 2, static_cast<cuuint64_t>(d.Win) * C * 2,
                             static_cast<cuuint64_t>(d.Hin) * d.Win * C * 2};
    cuuint32_t box[4] = {KB, static_cast<cuuint32_t>(BW * d.stride), static_cast<cuuint32_t>(BH * d.stride),
                         static_cast<cuuint32_t>(BNIMG)};
    cuuint32_t estr[4] = {1, static_cast<cuuint32_t>(d.stride), static_cast<cuuint32_t>(d.stride), 1};
    if (encode_map(&P->tmA[s], d.src[s], 4, dims, strides, box, estr) != 0) { delete P; return -1; }
  }
  for (int s = d.nsrc; s < 3; ++s) P->tmA[s] = P->tmA[0];
  kp.num_kb = kb;
  if (kb * KB != d.Ktot) {
    delete P;
    DMC_REQUIRE(false, "conv: Ktot=%d does not match the sources (%d)", d.Ktot, kb * KB);
  }

  const TileCfg tc = pick_cfg(d.Cout_pad, kp.num_m_tiles);
  const int BN = tc.bn;
  P->BN = BN;
  P->MT = tc.mt;
  P->CG = tc.cg;
  kp.num_n_tiles = d.Cout_pad / BN;
  {
    cuuint64_t dims[2] = {static_cast<cuuint64_t>(d.Ktot), static_cast<cuuint64_t>(d.Cout_pad)};
    cuuint64_t strides[1] = {static_cast<cuuint64_t>(d.Ktot) * 2};
    cuuint32_t box[2] = {KB, static_cast<cuuint32_t>(BN / tc.cg)};
    cuuint32_t estr[2] = {1, 1};
    if (encode_map(&P->tmB, d.weight, 2, dims, strides, box, estr) != 0) { delete P; return -1; }
  }
  P->tmOut = P->tmB;
  kp.tma_store = 0;
  {
    const char* e = getenv("DMC_CONV_TMA_STORE");  // "0": per-thread stores everywhere (debug / A-B measurements)
    const bool allow = !(e && e[0] == '0');
    if (allow && d.out_bf16 != nullptr && d.Cout % 64 == 0 && BN >= 128) {
      const int qbw = std::min(BW, 32), qbh = std::min(BH, 32 / qbw), qbn = 32 / (qbw * qbh);
      const int os = kp.oscale;
      cuuint64_t dims[4] = {static_cast<cuuint64_t>(d.Cout), static_cast<cuuint64_t>(Wout), static_cast<cuuint64_t>(Hout),
                            static_cast<cuuint64_t>(d.B)};
      cuuint64_t strides[3] = {static_cast<cuuint64_t>(os) * d.Cout * 2, static_cast<cuuint64_t>(os) * kp.out_W * d.Cout * 2,
                               static_cast<cuuint64_t>(kp.out_H) * kp.out_W * d.Cout * 2};
      cuuint32_t box[4] = {KB, static_cast<cuuint32_t>(qbw), static_cast<cuuint32_t>(qbh), static_cast<cuuint32_t>(qbn)};
      cuuint32_t estr[4] = {1, 1, 1, 1};
      const char* base = reinterpret_cast<const char*>(d.out_bf16) +
                         (static_cast<size_t>(kp.ooff_h) * kp.out_W + kp.ooff_w) * d.Cout * 2;
      if (encode_map(&P->tmOut, base, 4, dims, strides, box, estr) != 0) { delete P; return -1; }
      kp.tma_store = 1;
      kp.qbw = qbw;
      kp.qbh = qbh;
    }
  }
  kp.Cout = d.Cout;
  kp.bias = d.bias; kp.cond = d.cond; kp.cond_stride = d.cond_stride;
  kp.residual = reinterpret_cast<const __nv_bfloat16*>(d.residual);
  kp.out = reinterpret_cast<__nv_bfloat16*>(d.out_bf16);
  kp.out_nchw = d.out_f32_nchw;
  kp.stats = d.stats;
  kp.act = d.act;
  kp.gate = d.gate; kp.gate_stride = d.gate_stride;
  kp.residual_f32 = d.residual_f32; kp.out_f32 = d.out_f32_nhwc;
  kp.unpatch_p = d.unpatch_p;
  if (d.stats) {
    const int ppi_img = Hout * Wout;  // iteration pixels per image
    const int base = ppi_img >= 32 ? ppi_img / 32 : 1;
    const int want = base * (d.up_phase >= 0 ? 4 : 1);
    if (ppi_img < 16 || (ppi_img % 32 != 0 && ppi_img != 16) || d.stats_slots != want) {
      delete P;
      DMC_REQUIRE(false, "conv: stats_slots=%d but this geometry (%d pixels/image, up_phase %d) writes %d slots",
                  d.stats_slots, ppi_img, d.up_phase, want);
    }
    kp.stats_slots = want;
    kp.stats_slot_base = d.up_phase >= 0 ? d.up_phase * base : 0;
  }
  {
    const int mtg = tc.mt * tc.cg;
    const int group_tiles = ((kp.num_m_tiles + mtg - 1) / mtg) * kp.num_n_tiles;
    P->grid = tc.cg * std::min(group_tiles, num_sms() / tc.cg);
  }
  P->smem = 0;
  *out = P;
  return 0;
}

void conv_release(ConvPrepared* p) { delete p; }

template <int BN, int MT, int CG>
static int launch_cfg(const ConvPrepared* P, const ConvKParams& kp, cudaStream_t st) {
  static bool attr_set = false;
  if (!attr_set) {
    DMC_CUDA_OK(cudaFuncSetAttribute(conv_umma_kernel<BN, MT, CG>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     static_cast<int>(ConvCfg<BN, MT, CG>::SMEM)));
    attr_set = true;
  }
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(P->grid);
  cfg.blockDim = dim3(CONV_THREADS);
  cfg.dynamicSmemBytes = ConvCfg<BN, MT, CG>::SMEM;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CG;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  DMC_CUDA_OK(cudaLaunchKernelEx(&cfg, conv_umma_kernel<BN, MT, CG>, P->tmA[0], P->tmA[1], P->tmA[2], P->tmB, P->tmOut, kp));
  return 0;
}

int launch_conv(const dmc_conv_desc& d, const ConvPrepared* P, cudaStream_t st) {
  ConvKParams kp = P->kp;
  kp.out_nchw = d.out_f32_nchw;  // the only re-bindable pointer (dmc_plan_rebind which=2)
  if (P->CG == 2) return P->BN == 256 ? launch_cfg<256, 1, 2>(P, kp, st) : launch_cfg<128, 2, 2>(P, kp, st);
  if (P->MT == 2) return launch_cfg<128, 2, 1>(P, kp, st);
  switch (P->BN) {
    case 256: return launch_cfg<256, 1, 1>(P, kp, st);
    case 128: return launch_cfg<128, 1, 1>(P, kp, st);
    case 64: return launch_cfg<64, 1, 1>(P, kp, st);
    default: return launch_cfg<32, 1, 1>(P, kp, st);
  }
}

}  // namespace dmc
