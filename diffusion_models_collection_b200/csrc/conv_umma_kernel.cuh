// tcgen05 implicit-GEMM convolution kernel (see conv_umma.cu for the description): parameter block, tile configuration,
// the kernel template and its launcher.  Included by conv_umma.cu (host side) and by the conv_inst_*.cu translation
// units, each of which instantiates the kernel variants of ONE tile configuration (parallel, bounded compile times).
#pragma once

#include "common.cuh"
#include "kernels.h"

#include <stdlib.h>
#include <string.h>

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
  int prefetch_cond;          // prefetch the tile's conditioning row into L1 before waiting for the accumulator
  // A-operand affine transform (GroupNorm without SiLU applied to the input of a 1x1 convolution, models/unet.py:80-81): the
  // otherwise idle warps 2 and 3 rewrite every landed A tile in shared memory as x * scale[n, c] + shift[n, c] before the MMAs
  const float2* a_affine;     // [B, C_src] (scale, shift) of source 0, or nullptr
  int aff_C;                  // channels of source 0
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
  const __nv_bfloat16* residual_lo;  // split-bf16 mode: low parts of the residual and of the output
  __nv_bfloat16* out_lo;
  // TMA-store epilogue: every epilogue warp stores its 32 rows x 64 channels as one box {64, qbw, qbh, qbn}
  int tma_store;
  int qbw, qbh;               // quarter box: qbw pixels x qbh rows x 32/(qbw*qbh) images
  int store_bufs;             // staging buffers per epilogue warp (1 or 2)
  int res_tma;                // the residual (bf16 or fp32 stream) is fetched as TMA boxes into the staging buffer
  int epi_fast;               // bf16 TMA-store epilogue with bias only (+ GELU, + TMA residual): the straight-line box loop
  // shared-memory plan (host computed): [resident weights][ring: nst stages][store staging][barriers]
  int nst, stage_bytes, a_bytes, b_region_bytes;
  int bres;                   // weights of the whole K extent stay resident (short-K GEMMs: 1x1 convs, transformer linears)
  int slab, slab_bytes;       // 3x3 segment 0 is loaded as row slabs shared by the three vertical taps
  int slab_nv, slab_nh;       // vertical / horizontal taps of the slab segment: 3 x 3, or 2 x 2 for an Upsample phase convolution
  int slab_dh0, slab_dw0;     // offset of its first tap (-1 for 3x3; -1 or 0 per axis for a phase)
  // fused GroupNorm(+SiLU) of the output (VAR_GN): see dmc_conv_desc.gn_*
  int gn_nver;                // normalised versions written by the epilogue (1 or 2)
  int gn_P;                   // output pixels per image
  int gn_imgs;                // images (or the one partial image) in this CTA's table: P <= 128 * MT ? 128 * MT / P : 1
  int gn_ctas_per_img;        // CTAs holding pieces of one image (1: the shared-memory barrier is the only hand-shake)
  int gn_tab_groups;          // table columns per version (BN / smallest group size)
  int gn_sc_smem;             // per-channel scale / shift of the tile's image precomputed once per tile in shared memory
  int* gn_counters;           // [image * num_n_tiles + n_tile][arrived, done]
  int gn_debug;               // DMC_GN_DEBUG (timing experiments only, results are WRONG): 1 no cross-CTA wait, 2 no pass 2,
                              // 4 no table, 8 no park-in-TMEM
  float gn_eps;
  __nv_bfloat16* gn_out[2];
  int gn_pitch[2], gn_coff[2], gn_gsize[2], gn_silu[2];
  const float* gn_gamma[2];
  const float* gn_beta[2];
};

constexpr int MAX_NST = 8;
constexpr int SMEM_LIMIT = 227 * 1024;  // opt-in dynamic shared memory per CTA on sm_100

struct ConvPrepared {
  CUtensorMap tmA[3];
  CUtensorMap tmS;
  CUtensorMap tmB;
  CUtensorMap tmOut;
  CUtensorMap tmRes;
  CUtensorMap tmV[2];  // fused-GroupNorm versions (TMA-store epilogue)
  ConvKParams kp;
  int BN, MT, CG;
  int var;  // kernel variant (VAR_* bits)
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
  static constexpr int STORE_BYTES = (EPI_THREADS / 32) * 4096;
  static constexpr int ACC_COLS = MT * BN;  // TMEM columns of one accumulator stage
  static constexpr int TMEM_COLS = (2 * ACC_COLS < 32) ? 32 : 2 * ACC_COLS;
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

// VAR (compile-time kernel variant, so that each instantiation carries only the code it runs -- the all-in-one kernel was
// 107 KB of SASS and lost 8 % to instruction fetch):
//   bit 0 SLAB   3x3 segment 0 loaded as row slabs shared by the three vertical taps
//   bit 1 BRES   weights of the whole K extent resident in shared memory (short-K GEMMs)
//   bit 2 TS     TMA-store epilogue (BN >= 128 only)
//   bits 3.. EPI 0: UNet convolution (bias, conditioning, bf16 residual, bf16 NHWC out, GroupNorm partial sums)
//                1: transformer linear, bf16 NHWC out (bias, GELU)
//                2: model head (bias, fp32 NCHW out, optional unpatchify)
//                3: transformer linear on the fp32 residual stream (bias, gate, fp32 residual, fp32 NHWC out)
//                4: UNet convolution in split-bf16 mode: like 0, the residual and the output are (hi, lo) bf16 pairs
//   bit 6 GN     EPI 0 only: GroupNorm(+SiLU) of the output fused into the epilogue (two passes over the tile in tensor memory)
constexpr int VAR_SLAB = 1, VAR_BRES = 2, VAR_TS = 4, VAR_EPI_SHIFT = 3, VAR_GN = 64;
constexpr int GN_TAB_OFFSET = 512;  // the mean / rstd table follows the barrier block

template <int BN, int MT, int CG, int VAR>
__global__ void __launch_bounds__(CONV_THREADS, 1)
conv_umma_kernel(const __grid_constant__ CUtensorMap tmA0, const __grid_constant__ CUtensorMap tmA1,
                 const __grid_constant__ CUtensorMap tmA2, const __grid_constant__ CUtensorMap tmS,
                 const __grid_constant__ CUtensorMap tmB, const __grid_constant__ CUtensorMap tmOut,
                 const __grid_constant__ CUtensorMap tmRes, const __grid_constant__ CUtensorMap tmV0,
                 const __grid_constant__ CUtensorMap tmV1, const __grid_constant__ ConvKParams p) {
  using Cfg = ConvCfg<BN, MT, CG>;
  constexpr bool SLAB = (VAR & VAR_SLAB) != 0;
  constexpr bool BRES = (VAR & VAR_BRES) != 0;
  constexpr bool TS = (VAR & VAR_TS) != 0 && Cfg::TMA_STORE;
  constexpr int EPI = (VAR >> VAR_EPI_SHIFT) & 7;
  constexpr bool GN = (VAR & VAR_GN) != 0 && EPI == 0;
  constexpr bool UNET = EPI == 0 || EPI == 4;  // bias + conditioning + bf16 residual + GroupNorm partial sums
  constexpr int MTG = MT * CG;  // 128-pixel tiles per CTA-group tile
  constexpr int B_TILE = Cfg::B_STAGE_BYTES;
  const int NST = p.nst;
  const int rank = (CG == 2) ? static_cast<int>(cluster_ctarank()) : 0;
  const int group = blockIdx.x / CG, num_groups = gridDim.x / CG;
  extern __shared__ uint8_t smem_raw[];
  // [resident weights (bres)] [ring: nst x stage_bytes] [TMA-store staging] [barriers]
  uint8_t* smem = smem_align1024(smem_raw);
  uint8_t* ring = smem + p.b_region_bytes;
  uint8_t* store_stage = ring + NST * p.stage_bytes;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(store_stage + (TS ? Cfg::STORE_BYTES * p.store_bufs : 0));
  uint64_t* empty_bar = full_bar + MAX_NST;
  uint64_t* tfull_bar = empty_bar + MAX_NST;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint64_t* bfull_bar = tempty_bar + 2;   // resident weights loaded
  uint64_t* rbar = bfull_bar + 1;         // residual boxes: [epilogue warp][staging buffer]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(rbar + 2 * (EPI_THREADS / 32));
  uint64_t* afull_bar = rbar + 2 * (EPI_THREADS / 32) + 1;  // A-affine mode: this CTA's A tile has landed (per stage)
  uint64_t* xf_bar = afull_bar + MAX_NST;                   // ... and has been transformed in BOTH CTAs of a pair (on the leader)

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int num_ct = (p.num_m_tiles + MTG - 1) / MTG;  // group tiles along M (MTG consecutive 128-pixel tiles each)
  const int num_tiles = num_ct * p.num_n_tiles;
  // Tile schedule: strided over the CTA groups, n fastest -- groups running side by side work on the same pixels, so
  // the A tile comes from HBM once and from L2 for the other n tiles.  With resident weights a group keeps ONE n tile
  // for its whole life (n_tile = group % num_n_tiles, loaded once) and strides over the M tiles with the groups that
  // share its n; the (num_groups % num_n_tiles) left-over groups stay idle.
  constexpr bool bres = BRES;
  const int lanes = num_groups / p.num_n_tiles;  // groups per n tile (BRES)
  const bool bres_idle = bres && group >= lanes * p.num_n_tiles;
  const int t_begin = bres ? (bres_idle ? num_ct : group / p.num_n_tiles) : group;
  const int t_end = bres ? num_ct : num_tiles;
  const int t_step = bres ? lanes : num_groups;
#define DMC_DECODE_TILE(tile, n_tile, ct)                                   \
  const int n_tile = bres ? group % p.num_n_tiles : (tile) % p.num_n_tiles; \
  const int ct = bres ? (tile) : (tile) / p.num_n_tiles;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA0);
    if (p.nseg > 1) tma_prefetch_desc(&tmA1);
    if (p.nseg > 2) tma_prefetch_desc(&tmA2);
    if (SLAB) tma_prefetch_desc(&tmS);
    tma_prefetch_desc(&tmB);
    if (TS) tma_prefetch_desc(&tmOut);
    if (TS && p.res_tma) tma_prefetch_desc(&tmRes);
    if (GN && TS) {
      tma_prefetch_desc(&tmV0);
      if (p.gn_nver > 1) tma_prefetch_desc(&tmV1);
    }
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
    mbar_init(bfull_bar, 1);
    if (p.a_affine != nullptr) {
      for (int i = 0; i < NST; ++i) {
        mbar_init(&afull_bar[i], 1);
        mbar_init(&xf_bar[i], CG * 2);  // warps 2 and 3 of every CTA of the group
      }
    }
    for (int i = 0; i < 2 * (EPI_THREADS / 32); ++i) mbar_init(&rbar[i], 1);
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

  // K loop of one tile as "steps" (one ring stage each):
  //   slab step (3x3 segment 0 when p.slab): ONE box of (MT*BH + 2) image rows x BW pixels x 64 channels shifted by dw
  //     serves the three vertical taps dh = -1, 0, +1 of MT vertically adjacent sub-tiles (sub-tile mt, tap dh reads rows
  //     [mt*BH + dh + 1, +BH) of the box: a 1024-byte aligned offset) -> the activations cross L2->SM 3x instead of 9x;
  //     the 2 x 2 windows of the Upsample phase convolutions run the same steps with slab_nv = slab_nh = 2 (2x instead of 4x);
  //   regular step: MT boxes of 128 pixels x 64 channels for one (tap, chunk) K block.
  const int slab_nv = p.slab_nv, slab_nh = p.slab_nh;
  const int slab_steps = SLAB ? slab_nh * p.seg_chunks[0] : 0;
  const int reg_kb0 = SLAB ? p.seg_kb_end[0] : 0;            // first K block handled by regular steps
  const int num_steps = slab_steps + (p.num_kb - reg_kb0);

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      uint32_t phase = 0;
      int cur_n = -1, stage = 0;
      const uint32_t bfull_cl = (CG == 2) ? mapa_shared(smem_u32(bfull_bar), 0) : 0;
      for (int tile = t_begin; tile < t_end; tile += t_step) {
        DMC_DECODE_TILE(tile, n_tile, ct)
        if (bres && cur_n < 0) {  // the weight tile of this group: every K block, once
          if (rank == 0) mbar_expect_tx(bfull_bar, static_cast<uint32_t>(CG) * p.num_kb * B_TILE);
          for (int kb = 0; kb < p.num_kb; ++kb) {
            if (CG == 2) tma_load_2d_cg2(smem + kb * B_TILE, &tmB, bfull_cl, kb * KB, n_tile * BN + rank * (BN / 2));
            else tma_load_2d(smem + kb * B_TILE, &tmB, bfull_bar, kb * KB, n_tile * BN);
          }
          cur_n = n_tile;
        }
        int w0[MT], h0[MT], n0[MT];
#pragma unroll
        for (int mt = 0; mt < MT; ++mt) {
          const int m_tile = ct * MTG + rank * MT + mt;  // may be past the end: its box is fully out of bounds -> zero fill
          w0[mt] = (m_tile % p.tiles_w) * p.BW * p.stride;
          h0[mt] = ((m_tile / p.tiles_w) % p.tiles_h) * p.BH * p.stride;
          n0[mt] = (m_tile / (p.tiles_w * p.tiles_h)) * p.BNIMG;
        }
        int seg = SLAB ? 1 : 0, kb_in_seg = 0;
        for (int step = 0; step < num_steps; ++step) {
          mbar_wait(&empty_bar[stage], phase ^ 1u);
          uint8_t* sa = ring + stage * p.stage_bytes;
          uint8_t* sb = sa + p.a_bytes;
          const uint32_t lbar = (CG == 2) ? mapa_shared(smem_u32(&full_bar[stage]), 0) : 0;
          if (SLAB && step < slab_steps) {
            const int chunk = step / slab_nh, dwi = step % slab_nh;
            // the leader's barrier counts the bytes of BOTH CTAs of a pair (the MMAs it issues read both)
            if (rank == 0) mbar_expect_tx(&full_bar[stage], static_cast<uint32_t>(CG) * (p.slab_bytes + (bres ? 0 : slab_nv * B_TILE)));
            if (CG == 2) tma_load_4d_cg2(sa, &tmS, lbar, chunk * KB, w0[0] + dwi + p.slab_dw0, h0[0] + p.slab_dh0, n0[0]);
            else tma_load_4d(sa, &tmS, &full_bar[stage], chunk * KB, w0[0] + dwi + p.slab_dw0, h0[0] + p.slab_dh0, n0[0]);
            if (!bres) {
#pragma unroll
              for (int dhi = 0; dhi < 3; ++dhi) {
                if (dhi >= slab_nv) break;
                const int kb = (dhi * slab_nh + dwi) * p.seg_chunks[0] + chunk;
                if (CG == 2) tma_load_2d_cg2(sb + dhi * B_TILE, &tmB, lbar, kb * KB, n_tile * BN + rank * (BN / 2));
                else tma_load_2d(sb + dhi * B_TILE, &tmB, &full_bar[stage], kb * KB, n_tile * BN);
              }
            }
          } else {
            const int kb = reg_kb0 + (step - slab_steps);
            const int chunks = p.seg_chunks[seg];
            const int tap = kb_in_seg / chunks, chunk = kb_in_seg % chunks;
            const CUtensorMap* tm = seg == 0 ? &tmA0 : (seg == 1 ? &tmA1 : &tmA2);
            if (bres && p.a_affine != nullptr) {
              // A-affine mode (resident weights: the ring holds A tiles only): every CTA counts ITS OWN A tile on its own
              // barrier -- its transform warps wait there, and hand the tile to the MMA warp through xf_bar
              mbar_expect_tx(&afull_bar[stage], Cfg::A_BYTES);
#pragma unroll
              for (int mt = 0; mt < MT; ++mt)
                tma_load_4d(sa + mt * A_STAGE_BYTES, tm, &afull_bar[stage], chunk * KB, w0[mt] + p.dw[seg][tap],
                            h0[mt] + p.dh[seg][tap], n0[mt]);
            } else {
            if (rank == 0) mbar_expect_tx(&full_bar[stage], static_cast<uint32_t>(CG) * (Cfg::A_BYTES + (bres ? 0 : B_TILE)));
            if (CG == 2) {
#pragma unroll
              for (int mt = 0; mt < MT; ++mt)
                tma_load_4d_cg2(sa + mt * A_STAGE_BYTES, tm, lbar, chunk * KB, w0[mt] + p.dw[seg][tap],
                                h0[mt] + p.dh[seg][tap], n0[mt]);
              if (!bres) tma_load_2d_cg2(sb, &tmB, lbar, kb * KB, n_tile * BN + rank * (BN / 2));
            } else {
#pragma unroll
              for (int mt = 0; mt < MT; ++mt)
                tma_load_4d(sa + mt * A_STAGE_BYTES, tm, &full_bar[stage], chunk * KB, w0[mt] + p.dw[seg][tap],
                            h0[mt] + p.dh[seg][tap], n0[mt]);
              if (!bres) tma_load_2d(sb, &tmB, &full_bar[stage], kb * KB, n_tile * BN);
            }
            }
            if (++kb_in_seg == p.seg_taps[seg] * chunks) {
              ++seg;
              kb_in_seg = 0;
            }
          }
          if (++stage == NST) {
            stage = 0;
            phase ^= 1u;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0 && rank == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(TILE_M * CG, BN);
      uint32_t local = 0, phase = 0;
      int cur_n = -1, stage = 0;
      const uint32_t bres_addr = smem_u32(smem);
      for (int tile = t_begin; tile < t_end; tile += t_step, ++local) {
        DMC_DECODE_TILE(tile, n_tile, ct)
        (void)ct;
        const uint32_t acc = local & 1u;
        const uint32_t acc_phase = (local >> 1) & 1u;
        mbar_wait(&tempty_bar[acc], acc_phase ^ 1u);
        if (bres && cur_n < 0) {
          mbar_wait(bfull_bar, 0u);
          cur_n = n_tile;
        }
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * Cfg::ACC_COLS;
        uint32_t accum = 0;
        for (int step = 0; step < num_steps; ++step) {
          mbar_wait((bres && p.a_affine != nullptr) ? &xf_bar[stage] : &full_bar[stage], phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(ring + stage * p.stage_bytes);
          const uint32_t sb = sa + p.a_bytes;
          if (SLAB && step < slab_steps) {
            const int chunk = step / slab_nh, dwi = step % slab_nh;
#pragma unroll
            for (int dhi = 0; dhi < 3; ++dhi) {
              if (dhi >= slab_nv) break;
              const int kb = (dhi * slab_nh + dwi) * p.seg_chunks[0] + chunk;
              const uint64_t bdesc = umma_desc_k_sw128(bres ? bres_addr + kb * B_TILE : sb + dhi * B_TILE);
#pragma unroll
              for (int mt = 0; mt < MT; ++mt) {
                const uint64_t adesc = umma_desc_k_sw128(sa + ((mt * p.BH + dhi) * p.BW) * 128);
#pragma unroll
                for (int k = 0; k < KB / 16; ++k) {
                  if (CG == 2) umma_bf16_cg2(d_tmem + mt * BN, adesc + 2 * k, bdesc + 2 * k, idesc, (accum | dhi | k) != 0 ? 1u : 0u);
                  else umma_bf16(d_tmem + mt * BN, adesc + 2 * k, bdesc + 2 * k, idesc, (accum | dhi | k) != 0 ? 1u : 0u);
                }
              }
            }
          } else {
            const int kb = reg_kb0 + (step - slab_steps);
            const uint64_t bdesc = umma_desc_k_sw128(bres ? bres_addr + kb * B_TILE : sb);
#pragma unroll
            for (int mt = 0; mt < MT; ++mt) {
              const uint64_t adesc = umma_desc_k_sw128(sa + mt * A_STAGE_BYTES);
#pragma unroll
              for (int k = 0; k < KB / 16; ++k) {
                // advance 16 bf16 = 32 bytes inside the 128-byte swizzle row: +2 in the (addr >> 4) field
                if (CG == 2) umma_bf16_cg2(d_tmem + mt * BN, adesc + 2 * k, bdesc + 2 * k, idesc, (accum | k) != 0 ? 1u : 0u);
                else umma_bf16(d_tmem + mt * BN, adesc + 2 * k, bdesc + 2 * k, idesc, (accum | k) != 0 ? 1u : 0u);
              }
            }
          }
          accum = 1;
          // frees the smem stage (in both CTAs of a pair) once these MMAs have read it
          if (CG == 2) umma_commit_cg2(&empty_bar[stage]);
          else umma_commit(&empty_bar[stage]);
          if (++stage == NST) {
            stage = 0;
            phase ^= 1u;
          }
        }
        // accumulators complete -> epilogue (of both CTAs)
        if (CG == 2) umma_commit_cg2(&tfull_bar[acc]);
        else umma_commit(&tfull_bar[acc]);
      }
    }
  } else if (warp < 4) {
    // ===================== A-operand affine transform (warps 2 and 3, 1x1 convolutions with resident weights) ==========
    // thread = (16-byte chunk cl of the 128-byte row, row group rg): it owns 8 fixed channels of the K block -- its
    // (scale, shift) pairs are loaded once per image -- and rewrites rows rg, rg + 8, ... in place; the physical chunk is
    // cl ^ (row & 7) = cl ^ rg (128-byte swizzle).  y = fma(x, scale, shift) rounded to bf16: bit-identical to gn_apply_kernel.
    if (bres && p.a_affine != nullptr) {
      const int tw = (warp - 2) * 32 + lane;
      const int cl = tw & 7, rg = tw >> 3;
      const int ppi = p.BW * p.BH;
      uint32_t phase = 0;
      int stage = 0;
      for (int tile = t_begin; tile < t_end; tile += t_step) {
        DMC_DECODE_TILE(tile, n_tile, ct)
        (void)n_tile;
        for (int step = 0; step < num_steps; ++step) {
          mbar_wait(&afull_bar[stage], phase);
          const int ch0 = step * KB + cl * 8;  // (one source, one tap: K block = 64-channel chunk `step`)
#pragma unroll 1
          for (int mt = 0; mt < MT; ++mt) {
            const int m_tile = ct * MTG + rank * MT + mt;
            const int n0 = (m_tile / (p.tiles_w * p.tiles_h)) * p.BNIMG;
            uint8_t* base = ring + stage * p.stage_bytes + mt * A_STAGE_BYTES + ((cl ^ rg) << 4);
            int cur = -1;
            float sc[8], sh[8];
#pragma unroll 4
            for (int i = 0; i < TILE_M / 8; ++i) {
              const int r = rg + 8 * i;
              const int img = min(n0 + r / ppi, p.B - 1);  // (rows past the batch are zero-filled and never stored)
              if (img != cur) {
                cur = img;
                const float4* c4 = reinterpret_cast<const float4*>(p.a_affine + static_cast<size_t>(img) * p.aff_C + ch0);
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                  const float4 t = __ldg(c4 + j);
                  sc[2 * j] = t.x; sh[2 * j] = t.y; sc[2 * j + 1] = t.z; sh[2 * j + 1] = t.w;
                }
              }
              uint4* ptr = reinterpret_cast<uint4*>(base + r * 128);
              const uint4 v = *ptr;
              const uint32_t w[4] = {v.x, v.y, v.z, v.w};
              uint32_t o[4];
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const float2 f = unpack_bf16x2(w[j]);
                o[j] = pack_bf16x2(fmaf(f.x, sc[2 * j], sh[2 * j]), fmaf(f.y, sc[2 * j + 1], sh[2 * j + 1]));
              }
              *ptr = make_uint4(o[0], o[1], o[2], o[3]);
            }
          }
          fence_proxy_async_smem();  // generic-proxy writes -> visible to the tensor core's shared-memory reads
          __syncwarp();
          if (lane == 0) {
            if (CG == 2) mbar_arrive_cluster(mapa_shared(smem_u32(&xf_bar[stage]), 0));
            else mbar_arrive(&xf_bar[stage]);
          }
          if (++stage == NST) {
            stage = 0;
            phase ^= 1u;
          }
        }
      }
    }
  } else {
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
    constexpr int BOXC = (EPI == 3) ? 32 : 64;  // channels per TMA box: 128-byte rows of fp32 / bf16
    uint32_t local = 0, boxi = 0, rphase = 0;
    for (int tile = t_begin; tile < t_end; tile += t_step, ++local) {
      const uint32_t acc = local & 1u;
      const uint32_t acc_phase = (local >> 1) & 1u;
      DMC_DECODE_TILE(tile, n_tile, ct)
      const int m_tile = ct * MTG + rank * MT + (MT == 2 ? grp : 0);
      const int tw = m_tile % p.tiles_w;
      const int th = (m_tile / p.tiles_w) % p.tiles_h;
      const int ti = m_tile / (p.tiles_w * p.tiles_h);
      const int n = ti * p.BNIMG + ni;
      const int oh = (th * p.BH + hi) * p.oscale + p.ooff_h;
      const int ow = (tw * p.BW + wi) * p.oscale + p.ooff_w;
      const bool valid = n < p.B;
      const size_t pix = (static_cast<size_t>(n) * p.out_H + oh) * p.out_W + ow;
      // TMA store: coordinates (iteration space) of the first row of this warp's 32-row quarter
      const int r0 = q * 32;
      const int sw0 = tw * p.BW + r0 % p.BW, sh0 = th * p.BH + (r0 / p.BW) % p.BH, sn0 = ti * p.BNIMG + r0 / ppi;
      uint8_t* my_stage = store_stage + (warp - 4) * 4096 * p.store_bufs;

      if (UNET && p.cond != nullptr && p.prefetch_cond && valid) {
        // the conditioning row of this tile's image is a first touch (one row per image per layer): start pulling its lines
        // into L1 while the tile's MMAs are still running instead of paying the L2 latency inside the chunk loop
        const float* cp = p.cond + static_cast<size_t>(n) * p.cond_stride + n_tile * BN + col0;
        asm volatile("prefetch.global.L1 [%0];" ::"l"(cp + (lane % (COLS / 32)) * 32));
      }
      mbar_wait(&tfull_bar[acc], acc_phase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * Cfg::ACC_COLS +
                             (MT == 2 ? grp * BN : 0) + col0;
      if constexpr (GN) {
        // ======================= fused GroupNorm(+SiLU) of the output: two passes over the tile =======================
        // pass 1  finish the tile (bias, conditioning, residual), write the GroupNorm partial sums of this warp's rows and
        //         park the finished fp32 values back in tensor memory -- NO bulk stores yet: the hand-shake below must
        //         not wait behind a tile's worth of output traffic;
        // sync    every warp holding a piece of the same image has written its partial sums: a named barrier inside the
        //         CTA; a release-add / acquire-poll on a self-resetting global counter when the image spans CTAs (all
        //         CTAs of a persistent grid are co-resident; the tile schedule puts the pieces of one image in the same
        //         iteration);
        // table   the 8 epilogue warps reduce the partial sums (slot order: bit-reproducible, batch-invariant, the same
        //         order as gn_apply_kernel) to mean / rstd per (image, group, version) in shared memory;
        // pass 2  re-read the tile from tensor memory: the raw output (when something reads it) and each normalised
        //         version (a channel slice of its tensor).
        const int C8 = p.Cout >> 3;
        const bool full = ppi >= 32;
        const bool raw = p.out != nullptr;
        const bool res_tma = TS && p.res_tma != 0;
#pragma unroll 1
        for (int c0 = 0; c0 < COLS; c0 += 32) {
          const int cg = n_tile * BN + col0 + c0;
          // with the TMA-store epilogue the raw output leaves in pass 1 (bulk-async stores do not delay the hand-shake below);
          // per-thread stores would, so without it the raw output is written in pass 2
          const bool raw1 = TS && raw;
          const bool box_start = (res_tma || raw1) && (c0 % BOXC) == 0;
          const bool box_end = (res_tma || raw1) && ((c0 + 32) % BOXC) == 0;
          const uint32_t bi = (TS && p.store_bufs == 2) ? (boxi & 1u) : 0u;
          uint8_t* stg = my_stage + bi * 4096u;
          uint64_t* rb = &rbar[(warp - 4) * 2 + bi];
          if (box_start) {  // the staging buffer is free once the TMA store issued from it has read it out of shared memory
            if (lane == 0) {
              if (p.store_bufs == 2) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
              else asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
              if (res_tma) {  // residual box: coalesced, asynchronous
                mbar_expect_tx(rb, 4096);
                tma_load_4d(stg, &tmRes, rb, cg, sw0, sh0, sn0);
              }
            }
            __syncwarp();
          }
          uint32_t r[32];
          tmem_ld_32x32(taddr + c0, r);
          float add[32];
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
          uint4 res[4];
          const bool has_res = p.residual != nullptr && valid;
          if (has_res && !res_tma) {
            const uint4* r4 = reinterpret_cast<const uint4*>(p.residual + pix * p.Cout + cg);
#pragma unroll
            for (int j = 0; j < 4; ++j) res[j] = __ldg(r4 + j);
          }
          tmem_ld_wait();
          float v[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]) + add[j];
          if (res_tma) {
            if (box_start) {
              mbar_wait(rb, (rphase >> bi) & 1u);
              rphase ^= 1u << bi;
            }
            const uint8_t* rowp = stg + lane * 128;
            const int half = (c0 >> 5) & 1;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const uint4 rv = *reinterpret_cast<const uint4*>(rowp + (((half * 4 + j) ^ (lane & 7)) << 4));
              const uint32_t w[4] = {rv.x, rv.y, rv.z, rv.w};
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                const float2 f = unpack_bf16x2(w[k]);
                v[8 * j + 2 * k] += f.x;
                v[8 * j + 2 * k + 1] += f.y;
              }
            }
          } else if (has_res) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const uint32_t w[4] = {res[j].x, res[j].y, res[j].z, res[j].w};
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                const float2 f = unpack_bf16x2(w[k]);
                v[8 * j + 2 * k] += f.x;
                v[8 * j + 2 * k + 1] += f.y;
              }
            }
          }
          if (raw1) {  // ---- raw output through the staging buffer (the residual box, if any, has been consumed above) ----
            uint8_t* rowp = stg + lane * 128;
            const int half = (c0 >> 5) & 1;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              uint4 u;
              u.x = pack_bf16x2(v[8 * j], v[8 * j + 1]);
              u.y = pack_bf16x2(v[8 * j + 2], v[8 * j + 3]);
              u.z = pack_bf16x2(v[8 * j + 4], v[8 * j + 5]);
              u.w = pack_bf16x2(v[8 * j + 6], v[8 * j + 7]);
              *reinterpret_cast<uint4*>(rowp + (((half * 4 + j) ^ (lane & 7)) << 4)) = u;
            }
          }
          if (box_end) {
            if (raw1) {
              fence_proxy_async_smem();
              __syncwarp();
              if (lane == 0) {
                tma_store_4d(&tmOut, stg, cg + 32 - BOXC, sw0, sh0, sn0);
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
              }
            }
            ++boxi;
          }
          // ---- GroupNorm partial sums of this warp's rows (same slots as the stand-alone consumers read) ----
          {
            float sv[8];
#pragma unroll
            for (int b = 0; b < 4; ++b) {
              float s_ = 0.f, ss_ = 0.f;
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                s_ += v[8 * b + j];
                ss_ = fmaf(v[8 * b + j], v[8 * b + j], ss_);
              }
              sv[b] = valid ? s_ : 0.f;
              sv[4 + b] = valid ? ss_ : 0.f;
            }
            float tot;
            int idx;
            reduce8(sv, full, lane, tot, idx);
            const bool writer = full ? ((lane & 3) == 0) : ((lane & 1) == 0);
            if (writer && valid) {
              const int wpi = ppi >> 5;
              const int slot = p.stats_slot_base + (full ? ((th * p.tiles_w + tw) * wpi + (q % wpi)) : 0);
              float* dst = p.stats + ((static_cast<size_t>(n) * p.stats_slots + slot) * C8 + ((cg >> 3) + (idx & 3))) * 2;
              __stcg(dst + (idx >> 2), tot);
            }
          }
          // ---- park the finished values in tensor memory for pass 2 ----
#pragma unroll
          for (int j = 0; j < 32; ++j) r[j] = __float_as_uint(v[j]);
          if (!(p.gn_debug & 8)) tmem_st_32x32(taddr + c0, r);
        }
        tmem_st_wait();
        // ---- sync: every partial sum of the image(s) of this tile is in global memory ----
        if (p.gn_ctas_per_img > 1 && !(p.gn_debug & 1)) {
          __syncwarp();
          if (lane == 0 && valid) {  // (BNIMG == 1 here: the whole warp belongs to image n)
            int* cnt = p.gn_counters + 2 * (static_cast<size_t>(n) * p.num_n_tiles + n_tile);
            const int expected = (EPI_THREADS / 32) * p.gn_ctas_per_img;
            // release: the partial sums of all lanes (ordered before this by __syncwarp) are visible to whoever sees the count
            asm volatile("red.release.gpu.global.add.s32 [%0], 1;" ::"l"(cnt) : "memory");
            uint32_t spins = 0;
            while (ld_acquire_gpu(cnt) < expected) {
              __nanosleep(20);
              if (++spins > (1u << 22)) {
                printf("dmc: GroupNorm image counter timed out (block %d warp %d image %d)\n", (int)blockIdx.x, warp, n);
                __trap();
              }
            }
            // self-resetting: the last warp to get past the wait clears both words for the next launch
            if (atomicAdd(cnt + 1, 1) == expected - 1) {
              cnt[0] = 0;
              cnt[1] = 0;
            }
          }
          __syncwarp();
        }
        named_bar_sync(1, EPI_THREADS);
        // ---- table: mean / rstd per (version, image of this tile, group) ----
        // (one buffer is enough: a warp reaches this tile's barrier only after it has finished pass 2 of the previous tile)
        float2* tab = reinterpret_cast<float2*>(reinterpret_cast<uint8_t*>(full_bar) + GN_TAB_OFFSET);
        float2* scsh = tab + 2 * p.gn_imgs * p.gn_tab_groups;  // [version][BN] (scale, shift), when p.gn_sc_smem
        if (!(p.gn_debug & 4)) {
          const int n_first = ((ct * MTG + rank * MT) * TILE_M) / p.gn_P;  // first image of this CTA's rows
          for (int ver = 0; ver < p.gn_nver; ++ver) {
            const int gsz = p.gn_gsize[ver], ng = BN / gsz, nb = gsz >> 3;
            const float fold = p.gn_silu[ver] ? 0.5f : 1.0f;
            for (int e = warp - 4; e < p.gn_imgs * ng; e += EPI_THREADS / 32) {
              const int il = e / ng, g = e % ng;
              const int n_e = n_first + il;
              float s_ = 0.f, ss_ = 0.f;
              if (n_e < p.B) {
                const float2* base = reinterpret_cast<const float2*>(p.stats) +
                                     static_cast<size_t>(n_e) * p.stats_slots * C8 + ((n_tile * BN + g * gsz) >> 3);
                for (int i2 = lane; i2 < nb * p.stats_slots; i2 += 32) {
                  const float2 t2 = __ldcg(base + static_cast<size_t>(i2 / nb) * C8 + i2 % nb);
                  s_ += t2.x;
                  ss_ += t2.y;
                }
              }
#pragma unroll
              for (int o = 16; o; o >>= 1) {  // butterfly: every lane ends with the totals
                s_ += __shfl_xor_sync(0xFFFFFFFFu, s_, o);
                ss_ += __shfl_xor_sync(0xFFFFFFFFu, ss_, o);
              }
              const float inv_cnt = 1.0f / (static_cast<float>(gsz) * static_cast<float>(p.gn_P));
              const float mean = s_ * inv_cnt;
              const float var = fmaxf(ss_ * inv_cnt - mean * mean, 0.f);
              const float rstd = rsqrtf(var + p.gn_eps);
              if (lane == 0) tab[(ver * p.gn_imgs + il) * p.gn_tab_groups + g] = make_float2(mean, rstd);
              if (p.gn_sc_smem) {  // (one image per tile) the channels of this group: scale / shift with the SiLU 1/2 folded in
                for (int c = lane; c < gsz; c += 32) {
                  const int cl = g * gsz + c;
                  const float a = rstd * __ldg(p.gn_gamma[ver] + n_tile * BN + cl);
                  scsh[ver * BN + cl] = make_float2(fold * a, fold * (__ldg(p.gn_beta[ver] + n_tile * BN + cl) - mean * a));
                }
              }
            }
          }
        }
        named_bar_sync(1, EPI_THREADS);
        // ---- pass 2: the raw output (ver == -1) and the normalised versions ----
        const int il_row = p.gn_imgs > 1 ? (((MT == 2 ? grp : 0) * TILE_M + row) / p.gn_P) : 0;
#pragma unroll 1
        for (int ver = (raw && !TS) ? -1 : 0; ver < ((p.gn_debug & 2) ? 0 : p.gn_nver); ++ver) {
          const int vi = ver < 0 ? 0 : ver;
          const int gsz = p.gn_gsize[vi];
          const bool norm = ver >= 0;
          const bool silu = norm && p.gn_silu[vi] != 0;
          const float fold = silu ? 0.5f : 1.0f;  // SiLU(y) = h + h tanh(h), h = y / 2: folded into scale / shift
          const float2* trow = tab + (vi * p.gn_imgs + il_row) * p.gn_tab_groups;
          const float* gam = p.gn_gamma[vi];
          const float* bet = p.gn_beta[vi];
          __nv_bfloat16* vout = norm ? p.gn_out[vi] : p.out;
          const int vpitch = norm ? p.gn_pitch[vi] : p.Cout, vcoff = norm ? p.gn_coff[vi] : 0;
          const CUtensorMap* tmv = !norm ? &tmOut : (ver == 0 ? &tmV0 : &tmV1);
#pragma unroll 1
          for (int c0 = 0; c0 < COLS; c0 += 32) {
            const int cl = col0 + c0;  // first channel of this chunk inside the n tile
            const int cg = n_tile * BN + cl;
            const bool box_start = TS && (c0 % BOXC) == 0;
            const bool box_end = TS && ((c0 + 32) % BOXC) == 0;
            const uint32_t bi = (TS && p.store_bufs == 2) ? (boxi & 1u) : 0u;
            uint8_t* stg = my_stage + bi * 4096u;
            if (box_start) {
              if (lane == 0) {
                if (p.store_bufs == 2) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
                else asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
              }
              __syncwarp();
            }
            uint32_t r[32];
            tmem_ld_32x32(taddr + c0, r);
            uint32_t o[16];
            if (norm && p.gn_sc_smem) {
              const float4* t4 = reinterpret_cast<const float4*>(scsh + vi * BN + cl);  // (sc, sh) pairs: broadcast reads
              tmem_ld_wait();
#pragma unroll
              for (int j = 0; j < 16; ++j) {
                const float4 t = t4[j];
                float y0 = fmaf(__uint_as_float(r[2 * j]), t.x, t.y);
                float y1 = fmaf(__uint_as_float(r[2 * j + 1]), t.z, t.w);
                if (silu) {
                  y0 = silu_from_half(y0);
                  y1 = silu_from_half(y1);
                }
                o[j] = pack_bf16x2(y0, y1);
              }
            } else if (norm) {
              const float2 m0 = trow[cl / gsz], m1 = trow[(cl + 16) / gsz];
              float sc[32], sh[32];
              const float4* g4 = reinterpret_cast<const float4*>(gam + cg);
              const float4* b4 = reinterpret_cast<const float4*>(bet + cg);
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                const float4 gv = __ldg(g4 + j), bv = __ldg(b4 + j);
                const float2 m = j < 4 ? m0 : m1;
                const float a0 = m.y * gv.x, a1 = m.y * gv.y, a2 = m.y * gv.z, a3 = m.y * gv.w;
                sc[4 * j] = fold * a0; sc[4 * j + 1] = fold * a1; sc[4 * j + 2] = fold * a2; sc[4 * j + 3] = fold * a3;
                sh[4 * j] = fold * (bv.x - m.x * a0); sh[4 * j + 1] = fold * (bv.y - m.x * a1);
                sh[4 * j + 2] = fold * (bv.z - m.x * a2); sh[4 * j + 3] = fold * (bv.w - m.x * a3);
              }
              tmem_ld_wait();
#pragma unroll
              for (int j = 0; j < 16; ++j) {
                float y0 = fmaf(__uint_as_float(r[2 * j]), sc[2 * j], sh[2 * j]);
                float y1 = fmaf(__uint_as_float(r[2 * j + 1]), sc[2 * j + 1], sh[2 * j + 1]);
                if (silu) {
                  y0 = silu_from_half(y0);
                  y1 = silu_from_half(y1);
                }
                o[j] = pack_bf16x2(y0, y1);
              }
            } else {
              tmem_ld_wait();
#pragma unroll
              for (int j = 0; j < 16; ++j) o[j] = pack_bf16x2(__uint_as_float(r[2 * j]), __uint_as_float(r[2 * j + 1]));
            }
            if (TS) {
              uint8_t* rowp = stg + lane * 128;
              const int half = (c0 >> 5) & 1;
#pragma unroll
              for (int j = 0; j < 4; ++j)
                *reinterpret_cast<uint4*>(rowp + (((half * 4 + j) ^ (lane & 7)) << 4)) =
                    make_uint4(o[4 * j], o[4 * j + 1], o[4 * j + 2], o[4 * j + 3]);
              if (box_end) {
                fence_proxy_async_smem();
                __syncwarp();
                if (lane == 0) {
                  tma_store_4d(tmv, stg, vcoff + cg + 32 - BOXC, sw0, sh0, sn0);
                  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                }
                ++boxi;
              }
            } else if (valid) {
              uint4* o4 = reinterpret_cast<uint4*>(vout + pix * vpitch + vcoff + cg);
#pragma unroll
              for (int j = 0; j < 4; ++j) o4[j] = make_uint4(o[4 * j], o[4 * j + 1], o[4 * j + 2], o[4 * j + 3]);
            }
          }
        }
      } else
      if (!idle) {
        bool fast_done = false;
        if constexpr (TS && (EPI == 0 || EPI == 1) && (COLS % 64) == 0) {
          // ---- straight-line box loop (short-K GEMMs: qkv / proj 1x1 convolutions, transformer qkv / fc1) ----
          // The generic chunk loop below executes ~210 instructions per 32 columns (run 14: 21 ISETP, 17 BRA, 13 BSSY / BSYNC,
          // 16 CS2R ... around 32 FADD + 16 F2FP + 8 LDG + 4 STS of real work) and the 8 epilogue warps -- two per scheduler, in
          // order -- are latency-bound on it (36 % of their samples on fixed-latency dependencies, 9 % resolving branches, 8 %
          // waiting for instruction fetch): 5 600 clocks per 128 x 256 tile against 2 048 clocks of MMAs in the qkv GEMM.
          // When the epilogue is bias only (+ GELU, + a residual fetched by TMA) this loop handles one 64-channel box per
          // iteration with both tcgen05.ld in flight and no per-chunk decisions; rows past the batch are computed on zeros
          // and clipped by the TMA store.
          if (p.epi_fast && n_tile * BN + col0 + COLS <= p.Cout) {
            fast_done = true;
            const bool res_tma = EPI == 0 && p.res_tma != 0;
            const uint32_t xr = static_cast<uint32_t>(lane & 7) << 4;
#pragma unroll 1
            for (int c0 = 0; c0 < COLS; c0 += 64) {
              const int cg = n_tile * BN + col0 + c0;
              const uint32_t bi = (p.store_bufs == 2) ? (boxi & 1u) : 0u;
              uint8_t* stg = my_stage + bi * 4096u;
              uint64_t* rb = &rbar[(warp - 4) * 2 + bi];
              if (lane == 0) {
                if (p.store_bufs == 2) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
                else asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
                if (res_tma) {
                  mbar_expect_tx(rb, 4096);
                  tma_load_4d(stg, &tmRes, rb, cg, sw0, sh0, sn0);
                }
              }
              __syncwarp();
              uint32_t r0[32], r1[32];
              tmem_ld_32x32(taddr + c0, r0);
              tmem_ld_32x32(taddr + c0 + 32, r1);
              float v[64];
              tmem_ld_wait();
              if (p.bias != nullptr) {
                const float4* b4 = reinterpret_cast<const float4*>(p.bias + cg);
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                  const float4 b = __ldg(b4 + j);
                  v[4 * j] = __uint_as_float(r0[4 * j]) + b.x;
                  v[4 * j + 1] = __uint_as_float(r0[4 * j + 1]) + b.y;
                  v[4 * j + 2] = __uint_as_float(r0[4 * j + 2]) + b.z;
                  v[4 * j + 3] = __uint_as_float(r0[4 * j + 3]) + b.w;
                }
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                  const float4 b = __ldg(b4 + 8 + j);
                  v[32 + 4 * j] = __uint_as_float(r1[4 * j]) + b.x;
                  v[32 + 4 * j + 1] = __uint_as_float(r1[4 * j + 1]) + b.y;
                  v[32 + 4 * j + 2] = __uint_as_float(r1[4 * j + 2]) + b.z;
                  v[32 + 4 * j + 3] = __uint_as_float(r1[4 * j + 3]) + b.w;
                }
              } else {
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                  v[j] = __uint_as_float(r0[j]);
                  v[32 + j] = __uint_as_float(r1[j]);
                }
              }
              if (EPI == 1 && p.act == 1) {
#pragma unroll
                for (int j = 0; j < 64; ++j) v[j] = gelu_fast(v[j]);
              }
              uint8_t* rowp = stg + lane * 128;
              if (res_tma) {
                mbar_wait(rb, (rphase >> bi) & 1u);
                rphase ^= 1u << bi;
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                  const uint4 rv = *reinterpret_cast<const uint4*>(rowp + ((static_cast<uint32_t>(j) << 4) ^ xr));
                  const uint32_t w[4] = {rv.x, rv.y, rv.z, rv.w};
#pragma unroll
                  for (int k = 0; k < 4; ++k) {
                    const float2 f = unpack_bf16x2(w[k]);
                    v[8 * j + 2 * k] += f.x;
                    v[8 * j + 2 * k + 1] += f.y;
                  }
                }
              }
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                uint4 u;
                u.x = pack_bf16x2(v[8 * j], v[8 * j + 1]);
                u.y = pack_bf16x2(v[8 * j + 2], v[8 * j + 3]);
                u.z = pack_bf16x2(v[8 * j + 4], v[8 * j + 5]);
                u.w = pack_bf16x2(v[8 * j + 6], v[8 * j + 7]);
                *reinterpret_cast<uint4*>(rowp + ((static_cast<uint32_t>(j) << 4) ^ xr)) = u;
              }
              fence_proxy_async_smem();
              __syncwarp();
              if (lane == 0) {
                tma_store_4d(&tmOut, stg, cg, sw0, sh0, sn0);  // clipped at the tensor bounds (n >= B)
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
              }
              ++boxi;
              if (UNET && p.stats != nullptr) {
                // GroupNorm partial sums of the output, exactly as the generic loop writes them (same slots, same order)
                const bool full = ppi >= 32;
#pragma unroll
                for (int hf = 0; hf < 2; ++hf) {
                  float sv[8];
#pragma unroll
                  for (int b = 0; b < 4; ++b) {
                    float s_ = 0.f, ss_ = 0.f;
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                      s_ += v[32 * hf + 8 * b + j];
                      ss_ = fmaf(v[32 * hf + 8 * b + j], v[32 * hf + 8 * b + j], ss_);
                    }
                    sv[b] = valid ? s_ : 0.f;
                    sv[4 + b] = valid ? ss_ : 0.f;
                  }
                  float tot;
                  int idx;
                  reduce8(sv, full, lane, tot, idx);
                  const bool writer = full ? ((lane & 3) == 0) : ((lane & 1) == 0);
                  if (writer && valid) {
                    const int wpi = ppi >> 5;
                    const int slot = p.stats_slot_base + (full ? ((th * p.tiles_w + tw) * wpi + (q % wpi)) : 0);
                    float* dst = p.stats + ((static_cast<size_t>(n) * p.stats_slots + slot) * (p.Cout >> 3) +
                                            (((cg + 32 * hf) >> 3) + (idx & 3))) * 2;
                    dst[idx >> 2] = tot;
                  }
                }
              }
            }
          }
        }
        if constexpr (TS && EPI == 3) {
          // ---- the same straight-line loop for the transformer residual GEMMs (out_proj, fc2): out = res + gate * (acc + bias),
          // fp32 boxes of 32 channels, residual fetched and result stored by TMA ----
          if (p.epi_fast && n_tile * BN + col0 + COLS <= p.Cout) {
            fast_done = true;
            const uint32_t xr = static_cast<uint32_t>(lane & 7) << 4;
            const float* gate_row = p.gate != nullptr ? p.gate + static_cast<size_t>(min(n, p.B - 1)) * p.gate_stride : nullptr;
#pragma unroll 1
            for (int c0 = 0; c0 < COLS; c0 += 32) {
              const int cg = n_tile * BN + col0 + c0;
              const uint32_t bi = (p.store_bufs == 2) ? (boxi & 1u) : 0u;
              uint8_t* stg = my_stage + bi * 4096u;
              uint64_t* rb = &rbar[(warp - 4) * 2 + bi];
              if (lane == 0) {
                if (p.store_bufs == 2) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
                else asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
                mbar_expect_tx(rb, 4096);
                tma_load_4d(stg, &tmRes, rb, cg, sw0, sh0, sn0);
              }
              __syncwarp();
              uint32_t r[32];
              tmem_ld_32x32(taddr + c0, r);
              float v[32];
              tmem_ld_wait();
              if (p.bias != nullptr) {
                const float4* b4 = reinterpret_cast<const float4*>(p.bias + cg);
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                  const float4 b = __ldg(b4 + j);
                  v[4 * j] = __uint_as_float(r[4 * j]) + b.x;
                  v[4 * j + 1] = __uint_as_float(r[4 * j + 1]) + b.y;
                  v[4 * j + 2] = __uint_as_float(r[4 * j + 2]) + b.z;
                  v[4 * j + 3] = __uint_as_float(r[4 * j + 3]) + b.w;
                }
              } else {
#pragma unroll
                for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
              }
              if (gate_row != nullptr) {
                const float4* g4 = reinterpret_cast<const float4*>(gate_row + cg);
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                  const float4 gv = __ldg(g4 + j);
                  v[4 * j] *= gv.x; v[4 * j + 1] *= gv.y; v[4 * j + 2] *= gv.z; v[4 * j + 3] *= gv.w;
                }
              }
              uint8_t* rowp = stg + lane * 128;
              mbar_wait(rb, (rphase >> bi) & 1u);
              rphase ^= 1u << bi;
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                float4* cp = reinterpret_cast<float4*>(rowp + ((static_cast<uint32_t>(j) << 4) ^ xr));
                const float4 rv = *cp;
                *cp = make_float4(v[4 * j] + rv.x, v[4 * j + 1] + rv.y, v[4 * j + 2] + rv.z, v[4 * j + 3] + rv.w);
              }
              fence_proxy_async_smem();
              __syncwarp();
              if (lane == 0) {
                tma_store_4d(&tmOut, stg, cg, sw0, sh0, sn0);
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
              }
              ++boxi;
            }
          }
        }
        if (!fast_done)
#pragma unroll 1
        for (int c0 = 0; c0 < COLS; c0 += 32) {
          const int cg = n_tile * BN + col0 + c0;  // first global output channel of this chunk
          const bool real = cg < p.Cout;           // false only for padded weight rows (warp-uniform)
          // ---- TMA epilogue I/O: a "box" is this warp's 32 rows x BOXC channels (128-byte rows in shared memory) ----
          const bool box_start = TS && (c0 % BOXC) == 0;
          const bool box_end = TS && ((c0 + 32) % BOXC) == 0;
          const uint32_t bi = (TS && p.store_bufs == 2) ? (boxi & 1u) : 0u;  // staging buffer of this box
          uint8_t* stg = my_stage + bi * 4096u;
          uint64_t* rb = &rbar[(warp - 4) * 2 + bi];
          const bool res_tma = TS && (EPI == 0 || EPI == 3) && p.res_tma != 0;
          if (box_start && real) {
            if (lane == 0) {
              // the staging buffer is free once the TMA store issued from it has read it out of shared memory
              if (p.store_bufs == 2) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
              else asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
              if (res_tma) {  // residual box: coalesced, asynchronous, lands in the staging buffer
                mbar_expect_tx(rb, 4096);
                tma_load_4d(stg, &tmRes, rb, cg, sw0, sh0, sn0);
              }
            }
            __syncwarp();
          }
          uint32_t r[32];
          tmem_ld_32x32(taddr + c0, r);
          // per-channel addends (warp-uniform addresses -> L1 broadcast), fetched while the TMEM load is in flight
          float add[32];
          if (EPI != 2 && real) {
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
            if (UNET && p.cond != nullptr && valid) {
              const float4* c4 = reinterpret_cast<const float4*>(p.cond + static_cast<size_t>(n) * p.cond_stride + cg);
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                const float4 b = __ldg(c4 + j);
                add[4 * j] += b.x; add[4 * j + 1] += b.y; add[4 * j + 2] += b.z; add[4 * j + 3] += b.w;
              }
            }
          }
          uint4 res[4];
          const bool has_res = UNET && p.residual != nullptr && valid && real;
          if (has_res && !res_tma) {
            const uint4* r4 = reinterpret_cast<const uint4*>(p.residual + pix * p.Cout + cg);
#pragma unroll
            for (int j = 0; j < 4; ++j) res[j] = __ldg(r4 + j);
          }
          tmem_ld_wait();
          if (!real) continue;
          float v[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
          if (EPI == 2) {
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
          if (EPI == 1 && p.act == 1) {
            if (p.out_lo != nullptr) {  // accuracy mode: the two-SFU-op form (max abs error 2.5e-5)
#pragma unroll
              for (int j = 0; j < 32; ++j) v[j] = gelu_f(v[j]);
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j) v[j] = gelu_fast(v[j]);
            }
          }
          if (EPI == 3 && p.gate != nullptr && valid) {
            const float4* g4 = reinterpret_cast<const float4*>(p.gate + static_cast<size_t>(n) * p.gate_stride + cg);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float4 gv = __ldg(g4 + j);
              v[4 * j] *= gv.x; v[4 * j + 1] *= gv.y; v[4 * j + 2] *= gv.z; v[4 * j + 3] *= gv.w;
            }
          }
          // ---- residual ----
          if (res_tma) {
            if (box_start) {  // (uniform per warp) the residual box of this staging buffer has landed
              mbar_wait(rb, (rphase >> bi) & 1u);
              rphase ^= 1u << bi;
            }
            const uint8_t* rowp = stg + lane * 128;
            if (EPI == 3) {  // fp32 row: 8 chunks of 4 floats
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                const float4 rv = *reinterpret_cast<const float4*>(rowp + ((j ^ (lane & 7)) << 4));
                v[4 * j] += rv.x; v[4 * j + 1] += rv.y; v[4 * j + 2] += rv.z; v[4 * j + 3] += rv.w;
              }
            } else {  // bf16 row: this 32-channel half = 4 chunks of 8
              const int half = (c0 >> 5) & 1;
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const uint4 rv = *reinterpret_cast<const uint4*>(rowp + (((half * 4 + j) ^ (lane & 7)) << 4));
                const uint32_t w[4] = {rv.x, rv.y, rv.z, rv.w};
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                  const float2 f = unpack_bf16x2(w[k]);
                  v[8 * j + 2 * k] += f.x;
                  v[8 * j + 2 * k + 1] += f.y;
                }
              }
            }
          } else {
            if (EPI == 3 && p.residual_f32 != nullptr && valid) {
              const float4* r4 = reinterpret_cast<const float4*>(p.residual_f32 + pix * p.Cout + cg);
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                const float4 rv = r4[j];  // plain load: out_f32 may alias the residual stream
                v[4 * j] += rv.x; v[4 * j + 1] += rv.y; v[4 * j + 2] += rv.z; v[4 * j + 3] += rv.w;
              }
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
              if (EPI == 4 && p.residual_lo != nullptr) {  // low part of the split-bf16 residual
                const uint4* r4 = reinterpret_cast<const uint4*>(p.residual_lo + pix * p.Cout + cg);
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                  const uint4 rv = r4[j];  // plain load: the output pair may alias the residual pair
                  const uint32_t w[4] = {rv.x, rv.y, rv.z, rv.w};
#pragma unroll
                  for (int k = 0; k < 4; ++k) {
                    const float2 f = unpack_bf16x2(w[k]);
                    v[8 * j + 2 * k] += f.x;
                    v[8 * j + 2 * k + 1] += f.y;
                  }
                }
              }
            }
          }
          // ---- output ----
          if (TS) {
            // row = lane, 128-byte rows, 16-byte chunks XOR-swizzled with the row index: conflict-free shared-memory
            // accesses and the layout SWIZZLE_128B tensor maps expect
            uint8_t* rowp = stg + lane * 128;
            if (EPI == 3) {
#pragma unroll
              for (int j = 0; j < 8; ++j)
                *reinterpret_cast<float4*>(rowp + ((j ^ (lane & 7)) << 4)) =
                    make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
            } else {
              const int half = (c0 >> 5) & 1;
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                uint4 u;
                u.x = pack_bf16x2(v[8 * j], v[8 * j + 1]);
                u.y = pack_bf16x2(v[8 * j + 2], v[8 * j + 3]);
                u.z = pack_bf16x2(v[8 * j + 4], v[8 * j + 5]);
                u.w = pack_bf16x2(v[8 * j + 6], v[8 * j + 7]);
                *reinterpret_cast<uint4*>(rowp + (((half * 4 + j) ^ (lane & 7)) << 4)) = u;
              }
            }
            if (box_end) {
              fence_proxy_async_smem();
              __syncwarp();
              if (lane == 0) {
                tma_store_4d(&tmOut, stg, cg + 32 - BOXC, sw0, sh0, sn0);  // clipped at the tensor bounds (n >= B)
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
              }
              ++boxi;
            }
          } else if (EPI == 3) {
            if (p.out_f32 != nullptr && valid) {
              float4* o4 = reinterpret_cast<float4*>(p.out_f32 + pix * p.Cout + cg);
#pragma unroll
              for (int j = 0; j < 8; ++j) o4[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
            }
          } else if (valid && p.out != nullptr) {
            uint4* o4 = reinterpret_cast<uint4*>(p.out + pix * p.Cout + cg);
            uint4* l4 = reinterpret_cast<uint4*>(p.out_lo + pix * p.Cout + cg);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              uint4 u;
              u.x = pack_bf16x2(v[8 * j], v[8 * j + 1]);
              u.y = pack_bf16x2(v[8 * j + 2], v[8 * j + 3]);
              u.z = pack_bf16x2(v[8 * j + 4], v[8 * j + 5]);
              u.w = pack_bf16x2(v[8 * j + 6], v[8 * j + 7]);
              o4[j] = u;
              if ((EPI == 4 || EPI == 1) && p.out_lo != nullptr) {  // lo = bf16(v - hi): 16 mantissa bits in the pair
                const uint32_t w[4] = {u.x, u.y, u.z, u.w};
                uint32_t lo[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                  const float2 f = unpack_bf16x2(w[k]);
                  lo[k] = pack_bf16x2(v[8 * j + 2 * k] - f.x, v[8 * j + 2 * k + 1] - f.y);
                }
                l4[j] = make_uint4(lo[0], lo[1], lo[2], lo[3]);
              }
            }
          }
          if (UNET && p.stats != nullptr) {
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
        if (CG == 2) mbar_arrive_cluster_relaxed(mapa_shared(smem_u32(&tempty_bar[acc]), 0));
        else mbar_arrive(&tempty_bar[acc]);
      }
    }
    // all TMA stores of this warp have left shared memory and are complete before the CTA exits
    if (lane == 0 && TS) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }

  tc_fence_before();
  __syncthreads();
  if (CG == 2) cluster_sync_all();  // no CTA leaves (or frees TMEM) while its peer may still signal / read it
  if (warp == 2) {
    tc_fence_after();
    if (CG == 2) tmem_dealloc_cg2(tmem_base, Cfg::TMEM_COLS);
    else tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
  }
#undef DMC_DECODE_TILE
}

// ------------------------------------------------------------------------------------------------
// launch: one translation unit per tile configuration (conv_inst_*.cu) instantiates the variants it can run
// ------------------------------------------------------------------------------------------------
template <int BN, int MT, int CG, int VAR>
static int launch_variant(const ConvPrepared* P, const ConvKParams& kp, cudaStream_t st) {
  static DeviceOnce attr_set;  // the opt-in shared-memory limit is a per-device function attribute
  int attr_dev = 0;
  if (attr_set.need(&attr_dev)) {
    DMC_CUDA_OK(cudaFuncSetAttribute(conv_umma_kernel<BN, MT, CG, VAR>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_LIMIT));
    attr_set.done(attr_dev);
  }
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(P->grid);
  cfg.blockDim = dim3(CONV_THREADS);
  cfg.dynamicSmemBytes = P->smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CG;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  DMC_CUDA_OK(cudaLaunchKernelEx(&cfg, conv_umma_kernel<BN, MT, CG, VAR>, P->tmA[0], P->tmA[1], P->tmA[2], P->tmS, P->tmB,
                                 P->tmOut, P->tmRes, P->tmV[0], P->tmV[1], kp));
  return 0;
}

// The variants conv_prepare() can select (see conv_variant_supported): the TMA store needs BN >= 128; resident weights and
// slabs are exclusive (slabs are for 3x3 K extents, resident weights for at most 8 K blocks); transformer linears are
// never 3x3; the head writes fp32 NCHW with per-thread stores.
template <int BN, int MT, int CG>
static int launch_tile_cfg(const ConvPrepared* P, const ConvKParams& kp, cudaStream_t st) {
  constexpr bool TSOK = ConvCfg<BN, MT, CG>::TMA_STORE;
  switch (P->var) {
#define DMC_V(v) case (v): return launch_variant<BN, MT, CG, (v)>(P, kp, st)
    DMC_V(0);
    DMC_V(VAR_SLAB);
    DMC_V(VAR_BRES);
    DMC_V(1 << VAR_EPI_SHIFT);
    DMC_V((1 << VAR_EPI_SHIFT) | VAR_BRES);
    DMC_V(2 << VAR_EPI_SHIFT);
    DMC_V((2 << VAR_EPI_SHIFT) | VAR_SLAB);
    DMC_V(3 << VAR_EPI_SHIFT);
    DMC_V((3 << VAR_EPI_SHIFT) | VAR_BRES);
    DMC_V(4 << VAR_EPI_SHIFT);
    DMC_V((4 << VAR_EPI_SHIFT) | VAR_SLAB);
#undef DMC_V
    default: break;
  }
  if constexpr (BN >= 64) {  // fused GroupNorm epilogue (plain UNet convolutions only)
    switch (P->var) {
#define DMC_V(v) case (v): return launch_variant<BN, MT, CG, (v)>(P, kp, st)
      DMC_V(VAR_GN);
      DMC_V(VAR_GN | VAR_SLAB);
      DMC_V(VAR_GN | VAR_BRES);
#undef DMC_V
      default: break;
    }
  }
  if constexpr (TSOK) {
    switch (P->var) {
#define DMC_V(v) case (v): return launch_variant<BN, MT, CG, (v)>(P, kp, st)
      DMC_V(VAR_GN | VAR_TS);
      DMC_V(VAR_GN | VAR_TS | VAR_SLAB);
      DMC_V(VAR_GN | VAR_TS | VAR_BRES);
      DMC_V(VAR_TS);
      DMC_V(VAR_TS | VAR_SLAB);
      DMC_V(VAR_TS | VAR_BRES);
      DMC_V((1 << VAR_EPI_SHIFT) | VAR_TS);
      DMC_V((1 << VAR_EPI_SHIFT) | VAR_TS | VAR_BRES);
      DMC_V((3 << VAR_EPI_SHIFT) | VAR_TS);
      DMC_V((3 << VAR_EPI_SHIFT) | VAR_TS | VAR_BRES);
#undef DMC_V
      default: break;
    }
  }
  set_error("conv: kernel variant %d is not built for tile configuration (%d, %d, %d)", P->var, BN, MT, CG);
  return -1;
}

}  // namespace dmc
