// Weight gradient of a 3x3 / 1x1 convolution on the 5th-generation tensor cores (training step, SURVEY.md section 8 f2;
// reference: autograd of nn.Conv2d in /root/reference/models/unet.py:37,54,58,81,82,106,116).
//
//   dW[co, ci, r, s] = sum over (n, h, w) of  dY[n, h, w, co] * X[n, h*stride + r - pad, w*stride + s - pad, ci]
//
// GEMM view per tap (r, s):  D[M = co, N = ci] += A[M, K] * B[N, K]  with K = output pixels.  BOTH operands are MN-major:
// a TMA box {64 channels, pixel box} of the NHWC tensor lands as K rows (pixels) of 128 bytes (64 channels) with the
// 128-byte swizzle, which is exactly the MN-major SWIZZLE_128B canonical layout ((T,8,m),(8,k)):((1,T,LBO),(8T,SBO)) --
// so dY and the tap-shifted X tiles feed tcgen05.mma straight from the tensors the forward pass left in HBM (TMA zero
// fill = the convolution's padding), nothing is transposed or im2col-ed.
//
// One CTA = one work item (128 output channels, 64 input channels, a group of up to 5 taps, one slice of the pixels):
//   warp 0 lane 0 : TMA producer (ring of stages: dY tile 2 x 8 KB + one 8 KB X tile per tap, 64 pixels per stage)
//   warp 1 lane 0 : tcgen05.mma issuer (128 x 64 x 16 per tap and 16-pixel K step, fp32 accumulators in TMEM,
//                   64 columns per tap)
//   warp 2        : TMEM allocator
//   warps 4..7    : epilogue: TMEM -> fp32 partial sums [split][co][tap][ci]
// A second kernel adds the pixel slices in index order (deterministic, no atomics) into dW in the reference's
// [Cout, Cin, kh, kw] layout.
#include "common.cuh"
#include "kernels.h"

#include <stdlib.h>
#include <string.h>

#include <algorithm>

namespace dmc {

constexpr int WG_PIX = 64;                      // pixels (K) per stage
constexpr int WG_TILE_BYTES = WG_PIX * 128;     // one 64-channel MN block of 64 pixels
constexpr int WG_MAX_TAPS = 5;                  // taps per work item (TMEM: 5 x 64 columns)
constexpr int WG_THREADS = 256;
constexpr int WG_SMEM_LIMIT = 227 * 1024;

struct WgradParams {
  int B, Ho, Wo;            // output (dY) spatial size
  int stride;
  int BW, BH, BNIMG;        // pixel box of one 64-pixel tile
  int tiles_w, tiles_h, num_tiles;
  int Cin, Cout, taps;      // taps: 9 (3x3, pad 1) or 1 (1x1)
  int tap_groups, splits;
  int nst, stage_bytes;
  float* partial;           // [splits][Cout][taps][Cin]
};

// MN-major SWIZZLE_128B descriptor: 64-element MN blocks `lbo` bytes apart, groups of 8 K rows 1024 bytes apart
__device__ __forceinline__ uint64_t umma_desc_mn_sw128_lbo(uint32_t smem_addr, uint32_t lbo) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFF) >> 4);
  d |= static_cast<uint64_t>((lbo >> 4) & 0x3FFF) << 16;
  d |= static_cast<uint64_t>(1024 >> 4) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(2) << 61;
  return d;
}

__global__ void __launch_bounds__(WG_THREADS, 1)
conv_wgrad_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmDY,
                  const __grid_constant__ WgradParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_align1024(smem_raw);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + p.nst * p.stage_bytes);
  uint64_t* empty_bar = full_bar + 8;
  uint64_t* done_bar = empty_bar + 8;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(done_bar + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  // work item: blockIdx.x -> (split, tap group, ci chunk, co tile)
  int item = blockIdx.x;
  const int co_tiles = p.Cout / 128, ci_chunks = p.Cin / 64;
  const int co_tile = item % co_tiles; item /= co_tiles;
  const int ci_chunk = item % ci_chunks; item /= ci_chunks;
  const int tg = item % p.tap_groups; item /= p.tap_groups;
  const int split = item;
  const int tap0 = tg * WG_MAX_TAPS;
  const int ntap = min(WG_MAX_TAPS, p.taps - tap0);
  const int t_begin = static_cast<int>(static_cast<long long>(split) * p.num_tiles / p.splits);
  const int t_end = static_cast<int>(static_cast<long long>(split + 1) * p.num_tiles / p.splits);

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmX);
    tma_prefetch_desc(&tmDY);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < p.nst; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    mbar_init(done_bar, 1);
    mbar_fence_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int t = t_begin; t < t_end; ++t) {
        const int w0 = (t % p.tiles_w) * p.BW;
        const int h0 = ((t / p.tiles_w) % p.tiles_h) * p.BH;
        const int n0 = (t / (p.tiles_w * p.tiles_h)) * p.BNIMG;
        mbar_wait(&empty_bar[stage], phase ^ 1u);
        mbar_expect_tx(&full_bar[stage], static_cast<uint32_t>((2 + ntap) * WG_TILE_BYTES));
        uint8_t* sa = smem + stage * p.stage_bytes;
        tma_load_4d(sa, &tmDY, &full_bar[stage], co_tile * 128, w0, h0, n0);
        tma_load_4d(sa + WG_TILE_BYTES, &tmDY, &full_bar[stage], co_tile * 128 + 64, w0, h0, n0);
        for (int j = 0; j < ntap; ++j) {
          const int tap = tap0 + j;
          const int dh = p.taps == 9 ? tap / 3 - 1 : 0, dw = p.taps == 9 ? tap % 3 - 1 : 0;
          tma_load_4d(sa + (2 + j) * WG_TILE_BYTES, &tmX, &full_bar[stage], ci_chunk * 64, w0 * p.stride + dw,
                      h0 * p.stride + dh, n0);
        }
        if (++stage == p.nst) {
          stage = 0;
          phase ^= 1u;
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // M = 128 (two MN blocks of dY), N = 64 (one MN block of X), both MN-major: bits 15 and 16
      constexpr uint32_t idesc = umma_idesc_bf16(128, 64, 1) | (1u << 15);
      int stage = 0;
      uint32_t phase = 0, accum = 0;
      for (int t = t_begin; t < t_end; ++t) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        const uint32_t sa = smem_u32(smem + stage * p.stage_bytes);
#pragma unroll
        for (int kk = 0; kk < WG_PIX / 16; ++kk) {
          const uint64_t adesc = umma_desc_mn_sw128_lbo(sa + kk * 2048, WG_TILE_BYTES);
          for (int j = 0; j < ntap; ++j) {
            const uint64_t bdesc = umma_desc_mn_sw128_lbo(sa + (2 + j) * WG_TILE_BYTES + kk * 2048, WG_TILE_BYTES);
            umma_bf16(tmem_base + j * 64, adesc, bdesc, idesc, (accum | kk) != 0 ? 1u : 0u);
          }
        }
        accum = 1;
        umma_commit(&empty_bar[stage]);
        if (++stage == p.nst) {
          stage = 0;
          phase ^= 1u;
        }
      }
      umma_commit(done_bar);
    }
  } else if (warp >= 4) {
    const int q = warp & 3;
    const int co = co_tile * 128 + q * 32 + lane;
    if (t_end > t_begin) {
      mbar_wait(done_bar, 0u);
      tc_fence_after();
    }
    for (int j = 0; j < ntap; ++j) {
      float* dst = p.partial + ((static_cast<size_t>(split) * p.Cout + co) * p.taps + (tap0 + j)) * p.Cin + ci_chunk * 64;
#pragma unroll
      for (int c0 = 0; c0 < 64; c0 += 32) {
        uint32_t r[32];
        if (t_end > t_begin) {
          tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + j * 64 + c0, r);
          tmem_ld_wait();
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i) r[i] = 0u;  // an empty pixel slice contributes zeros
        }
#pragma unroll
        for (int i = 0; i < 8; ++i)
          reinterpret_cast<float4*>(dst + c0)[i] = make_float4(__uint_as_float(r[4 * i]), __uint_as_float(r[4 * i + 1]),
                                                                __uint_as_float(r[4 * i + 2]), __uint_as_float(r[4 * i + 3]));
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// Row-slab variant (stride 1, image width 8..64, Cin % 128 == 0): one CTA = 128 output channels x 128 input channels x the
// three VERTICAL taps of one column shift dw (or the single tap of a 1x1 convolution) x one slice of the pixels.
// A stage is BH full image rows (64 pixels).  The X operand of the three taps is ONE TMA box of BH + 2 rows: the tap (dh, dw)
// is the same shared-memory slab read BW pixels (a whole number of 1024-byte swizzle atoms) further down, so X moves
// L2 -> SM once instead of three times, and every MMA is 128 x 128 x 16 (half the shared-memory operand bytes per MAC of the
// 128 x 64 form above).
// ---------------------------------------------------------------------------------------------------------------------
struct WgradSlabParams {
  int BW, BH, tiles_h, num_tiles;
  int Cin, Cout, taps, ndw, splits;
  int nst, stage_bytes, slab_block_bytes, halo;
  float* partial;
};

__global__ void __launch_bounds__(WG_THREADS, 1)
conv_wgrad_slab_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmDY,
                       const __grid_constant__ WgradSlabParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_align1024(smem_raw);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + p.nst * p.stage_bytes);
  uint64_t* empty_bar = full_bar + 8;
  uint64_t* done_bar = empty_bar + 8;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(done_bar + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  int item = blockIdx.x;
  const int co_tiles = p.Cout / 128, ci_tiles = p.Cin / 128;
  const int co_tile = item % co_tiles; item /= co_tiles;
  const int ci_tile = item % ci_tiles; item /= ci_tiles;
  const int dwi = item % p.ndw; item /= p.ndw;
  const int split = item;
  const int nv = p.taps == 9 ? 3 : 1;      // vertical taps sharing the slab
  const int dw = p.taps == 9 ? dwi - 1 : 0;
  const int t_begin = static_cast<int>(static_cast<long long>(split) * p.num_tiles / p.splits);
  const int t_end = static_cast<int>(static_cast<long long>(split + 1) * p.num_tiles / p.splits);

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmX);
    tma_prefetch_desc(&tmDY);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < p.nst; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    mbar_init(done_bar, 1);
    mbar_fence_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int t = t_begin; t < t_end; ++t) {
        const int h0 = (t % p.tiles_h) * p.BH;
        const int n0 = t / p.tiles_h;
        mbar_wait(&empty_bar[stage], phase ^ 1u);
        mbar_expect_tx(&full_bar[stage], static_cast<uint32_t>(2 * WG_TILE_BYTES + 2 * p.slab_block_bytes));
        uint8_t* sa = smem + stage * p.stage_bytes;
        tma_load_4d(sa, &tmDY, &full_bar[stage], co_tile * 128, 0, h0, n0);
        tma_load_4d(sa + WG_TILE_BYTES, &tmDY, &full_bar[stage], co_tile * 128 + 64, 0, h0, n0);
        tma_load_4d(sa + 2 * WG_TILE_BYTES, &tmX, &full_bar[stage], ci_tile * 128, dw, h0 - p.halo, n0);
        tma_load_4d(sa + 2 * WG_TILE_BYTES + p.slab_block_bytes, &tmX, &full_bar[stage], ci_tile * 128 + 64, dw, h0 - p.halo, n0);
        if (++stage == p.nst) {
          stage = 0;
          phase ^= 1u;
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(128, 128, 1) | (1u << 15);  // both operands MN-major
      int stage = 0;
      uint32_t phase = 0, accum = 0;
      const uint32_t tap_bytes = static_cast<uint32_t>(p.BW) * 128u;  // one image row further down the slab
      for (int t = t_begin; t < t_end; ++t) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        const uint32_t sa = smem_u32(smem + stage * p.stage_bytes);
#pragma unroll
        for (int kk = 0; kk < WG_PIX / 16; ++kk) {
          const uint64_t adesc = umma_desc_mn_sw128_lbo(sa + kk * 2048, WG_TILE_BYTES);
          for (int v = 0; v < nv; ++v) {
            const uint64_t bdesc = umma_desc_mn_sw128_lbo(sa + 2 * WG_TILE_BYTES + v * tap_bytes + kk * 2048,
                                                          static_cast<uint32_t>(p.slab_block_bytes));
            umma_bf16(tmem_base + v * 128, adesc, bdesc, idesc, (accum | kk) != 0 ? 1u : 0u);
          }
        }
        accum = 1;
        umma_commit(&empty_bar[stage]);
        if (++stage == p.nst) {
          stage = 0;
          phase ^= 1u;
        }
      }
      umma_commit(done_bar);
    }
  } else if (warp >= 4) {
    const int q = warp & 3;
    const int co = co_tile * 128 + q * 32 + lane;
    if (t_end > t_begin) {
      mbar_wait(done_bar, 0u);
      tc_fence_after();
    }
    for (int v = 0; v < nv; ++v) {
      const int tap = p.taps == 9 ? v * 3 + dwi : 0;
      float* dst = p.partial + ((static_cast<size_t>(split) * p.Cout + co) * p.taps + tap) * p.Cin + ci_tile * 128;
#pragma unroll
      for (int c0 = 0; c0 < 128; c0 += 32) {
        uint32_t r[32];
        if (t_end > t_begin) {
          tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + v * 128 + c0, r);
          tmem_ld_wait();
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i) r[i] = 0u;
        }
#pragma unroll
        for (int i = 0; i < 8; ++i)
          reinterpret_cast<float4*>(dst + c0)[i] = make_float4(__uint_as_float(r[4 * i]), __uint_as_float(r[4 * i + 1]),
                                                                __uint_as_float(r[4 * i + 2]), __uint_as_float(r[4 * i + 3]));
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// dW[co][ci][tap] (+)= sum_split partial[split][co][tap][ci], slices added in index order
// dw may be a window of a larger parameter: rows co < dw_cout, columns ci0 .. ci0 + dw_cin of [dw_cout, dw_cin_total, kh, kw]
// (real channels of a zero-padded operand; one source of a fused 1x1 shortcut over concatenated inputs)
__global__ void __launch_bounds__(256) conv_wgrad_reduce_kernel(const float* __restrict__ partial, float* __restrict__ dw,
                                                                int splits, int Cout, int taps, int Cin, int accumulate,
                                                                int dw_cout, int dw_cin, int dw_cin_total, int dw_ci0) {
  const size_t total = static_cast<size_t>(Cout) * taps * Cin;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int ci = static_cast<int>(i % Cin);
    const int tap = static_cast<int>((i / Cin) % taps);
    const int co = static_cast<int>(i / (static_cast<size_t>(Cin) * taps));
    if (co >= dw_cout || ci >= dw_cin) continue;  // zero-padded rows / columns of the GEMM: not part of the parameter
    float s = 0.f;
    int k = 0;
    for (; k + 4 <= splits; k += 4) {  // four loads in flight, added in index order
      const float v0 = partial[static_cast<size_t>(k) * total + i], v1 = partial[static_cast<size_t>(k + 1) * total + i];
      const float v2 = partial[static_cast<size_t>(k + 2) * total + i], v3 = partial[static_cast<size_t>(k + 3) * total + i];
      s += v0; s += v1; s += v2; s += v3;
    }
    for (; k < splits; ++k) s += partial[static_cast<size_t>(k) * total + i];
    float* o = dw + (static_cast<size_t>(co) * dw_cin_total + dw_ci0 + ci) * taps + tap;  // [Cout, Cin, kh, kw]
    *o = accumulate ? *o + s : s;
  }
}

static int encode4d(CUtensorMap* m, const void* base, int C, int W, int H, int B, int bw, int bh, int bn, int stride) {
  EncodeTiledFn fn = encode_tiled_fn();
  DMC_REQUIRE(fn != nullptr, "wgrad: cuTensorMapEncodeTiled unavailable -- call dmc_init()");
  cuuint64_t dims[4] = {static_cast<cuuint64_t>(C), static_cast<cuuint64_t>(W), static_cast<cuuint64_t>(H),
                        static_cast<cuuint64_t>(B)};
  cuuint64_t strides[3] = {static_cast<cuuint64_t>(C) * 2, static_cast<cuuint64_t>(W) * C * 2,
                           static_cast<cuuint64_t>(H) * W * C * 2};
  cuuint32_t box[4] = {64, static_cast<cuuint32_t>(bw * stride), static_cast<cuuint32_t>(bh * stride),
                       static_cast<cuuint32_t>(bn)};
  cuuint32_t estr[4] = {1, static_cast<cuuint32_t>(stride), static_cast<cuuint32_t>(stride), 1};
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  DMC_REQUIRE(r == CUDA_SUCCESS, "wgrad: cuTensorMapEncodeTiled failed with CUresult %d", static_cast<int>(r));
  return 0;
}

// the row-slab kernel covers: stride 1, rows of 8..64 pixels, at least 64 pixels per image, 128-channel input tiles
static bool wgrad_use_slab(const dmc_wgrad_desc& d) {
  const char* e = getenv("DMC_WGRAD_SLAB");
  if (e && e[0] == '0') return false;
  const int W = d.Win, H = d.Hin;
  return d.stride == 1 && d.Cin % 128 == 0 && (W == 8 || W == 16 || W == 32 || W == 64) && H * W >= WG_PIX && H % (WG_PIX / W) == 0;
}

int conv_wgrad_splits(const dmc_wgrad_desc& d) {
  const int Ho = d.Hin / d.stride, Wo = d.Win / d.stride;
  const long long tiles = (static_cast<long long>(d.B) * Ho * Wo + WG_PIX - 1) / WG_PIX;
  const int items = wgrad_use_slab(d) ? (d.Cout / 128) * (d.Cin / 128) * (d.taps == 9 ? 3 : 1)
                                      : (d.Cout / 128) * (d.Cin / 64) * ((d.taps + WG_MAX_TAPS - 1) / WG_MAX_TAPS);
  // whole waves: one CTA per SM (most of the shared memory each), so the grid is the largest multiple of `items` that fits in
  // the wave budget (a grid of k * SMs + a few CTAs would run one more, almost empty wave)
  const int waves = wgrad_use_slab(d) ? 1 : 2;  // slab kernel: one wave (half the fp32 partial-sum traffic, half the epilogues)
  int splits = std::max(1, (waves * num_sms()) / std::max(items, 1));
  // at least 4 pixel tiles per CTA: below that the fp32 partial sums (splits x the weight tensor, written and read back) cost
  // more than the idle SMs
  return static_cast<int>(std::min<long long>(splits, std::max<long long>(tiles / 4, 1)));
}

static int launch_conv_wgrad_slab(const dmc_wgrad_desc& d, cudaStream_t st) {
  WgradSlabParams p;
  memset(&p, 0, sizeof(p));
  const int W = d.Win, H = d.Hin;
  p.BW = W; p.BH = WG_PIX / W;
  p.tiles_h = H / p.BH;
  p.num_tiles = d.B * p.tiles_h;
  p.Cin = d.Cin; p.Cout = d.Cout; p.taps = d.taps;
  p.ndw = d.taps == 9 ? 3 : 1;
  p.halo = d.taps == 9 ? 1 : 0;
  p.splits = d.splits;
  p.slab_block_bytes = (p.BH + 2 * p.halo) * W * 128;
  p.stage_bytes = 2 * WG_TILE_BYTES + 2 * p.slab_block_bytes;
  p.nst = std::min(8, (WG_SMEM_LIMIT - 1024 - 256) / p.stage_bytes);
  p.partial = d.partial;
  CUtensorMap tmX, tmDY;
  if (encode4d(&tmX, d.x, d.Cin, W, H, d.B, W, p.BH + 2 * p.halo, 1, 1) != 0) return -1;
  if (encode4d(&tmDY, d.dy, d.Cout, W, H, d.B, W, p.BH, 1, 1) != 0) return -1;
  static DeviceOnce attr_set;
  int attr_set_dev = 0;
  if (attr_set.need(&attr_set_dev)) {
    DMC_CUDA_OK(cudaFuncSetAttribute(conv_wgrad_slab_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, WG_SMEM_LIMIT));
    attr_set.done(attr_set_dev);
  }
  const int grid = (d.Cout / 128) * (d.Cin / 128) * p.ndw * p.splits;
  const size_t smem = static_cast<size_t>(p.nst) * p.stage_bytes + 1024 + 256;
  conv_wgrad_slab_kernel<<<grid, WG_THREADS, smem, st>>>(tmX, tmDY, p);
  DMC_CUDA_OK(cudaGetLastError());
  return 0;
}

int launch_conv_wgrad(const dmc_wgrad_desc& d, cudaStream_t st) {
  DMC_REQUIRE(d.x && d.dy && d.dw && d.partial, "wgrad: null pointer argument");
  DMC_REQUIRE(d.taps == 9 || d.taps == 1, "wgrad: taps=%d (3x3 with padding 1, or 1x1)", d.taps);
  DMC_REQUIRE(d.stride == 1 || d.stride == 2, "wgrad: stride=%d", d.stride);
  DMC_REQUIRE(d.Cin % 64 == 0 && d.Cout % 128 == 0, "wgrad: needs Cin %% 64 == 0 and Cout %% 128 == 0 (got %d, %d)", d.Cin, d.Cout);
  DMC_REQUIRE(d.B > 0 && d.Hin % d.stride == 0 && d.Win % d.stride == 0, "wgrad: bad geometry");
  const int Ho = d.Hin / d.stride, Wo = d.Win / d.stride;
  WgradParams p;
  memset(&p, 0, sizeof(p));
  p.B = d.B; p.Ho = Ho; p.Wo = Wo; p.stride = d.stride;
  int BW = std::min(Wo, WG_PIX);
  DMC_REQUIRE(WG_PIX % BW == 0 && Wo % BW == 0, "wgrad: Wout=%d unsupported", Wo);
  int BH = std::min(Ho, WG_PIX / BW);
  DMC_REQUIRE((WG_PIX / BW) % BH == 0 && Ho % BH == 0, "wgrad: Hout=%d unsupported", Ho);
  const int BNIMG = WG_PIX / (BW * BH);
  p.BW = BW; p.BH = BH; p.BNIMG = BNIMG;
  p.tiles_w = Wo / BW; p.tiles_h = Ho / BH;
  p.num_tiles = ((d.B + BNIMG - 1) / BNIMG) * p.tiles_w * p.tiles_h;
  p.Cin = d.Cin; p.Cout = d.Cout; p.taps = d.taps;
  p.tap_groups = (d.taps + WG_MAX_TAPS - 1) / WG_MAX_TAPS;
  p.splits = d.splits;
  DMC_REQUIRE(d.splits >= 1 && d.splits == conv_wgrad_splits(d), "wgrad: splits=%d, expected dmc_conv_wgrad_splits() = %d", d.splits,
              conv_wgrad_splits(d));
  const int w_cout = d.dw_cout > 0 ? d.dw_cout : d.Cout, w_cin = d.dw_cin > 0 ? d.dw_cin : d.Cin;
  const int w_cin_total = d.dw_cin_total > 0 ? d.dw_cin_total : w_cin;
  DMC_REQUIRE(w_cout <= d.Cout && w_cin <= d.Cin && d.dw_ci0 >= 0 && d.dw_ci0 + w_cin <= w_cin_total,
              "wgrad: bad dw window (%d x %d at column %d of %d)", w_cout, w_cin, d.dw_ci0, w_cin_total);
  const size_t total = static_cast<size_t>(d.Cout) * d.taps * d.Cin;
  const int rblocks = static_cast<int>(std::min<size_t>((total + 255) / 256, static_cast<size_t>(num_sms()) * 8));
  if (wgrad_use_slab(d)) {
    if (launch_conv_wgrad_slab(d, st) != 0) return -1;
    conv_wgrad_reduce_kernel<<<rblocks, 256, 0, st>>>(d.partial, d.dw, d.splits, d.Cout, d.taps, d.Cin, d.accumulate, w_cout, w_cin,
                                                      w_cin_total, d.dw_ci0);
    DMC_CUDA_OK(cudaGetLastError());
    return 0;
  }
  p.stage_bytes = (2 + std::min(d.taps, WG_MAX_TAPS)) * WG_TILE_BYTES;
  p.nst = std::min(8, (WG_SMEM_LIMIT - 1024 - 256) / p.stage_bytes);
  p.partial = d.partial;
  CUtensorMap tmX, tmDY;
  if (encode4d(&tmX, d.x, d.Cin, d.Win, d.Hin, d.B, BW, BH, BNIMG, d.stride) != 0) return -1;
  if (encode4d(&tmDY, d.dy, d.Cout, Wo, Ho, d.B, BW, BH, BNIMG, 1) != 0) return -1;
  static DeviceOnce attr_set;
  int attr_set_dev = 0;
  if (attr_set.need(&attr_set_dev)) {
    DMC_CUDA_OK(cudaFuncSetAttribute(conv_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, WG_SMEM_LIMIT));
    attr_set.done(attr_set_dev);
  }
  const int grid = (d.Cout / 128) * (d.Cin / 64) * p.tap_groups * p.splits;
  const size_t smem = static_cast<size_t>(p.nst) * p.stage_bytes + 1024 + 256;
  conv_wgrad_kernel<<<grid, WG_THREADS, smem, st>>>(tmX, tmDY, p);
  DMC_CUDA_OK(cudaGetLastError());
  conv_wgrad_reduce_kernel<<<rblocks, 256, 0, st>>>(d.partial, d.dw, p.splits, d.Cout, d.taps, d.Cin, d.accumulate, w_cout, w_cin,
                                                    w_cin_total, d.dw_ci0);
  DMC_CUDA_OK(cudaGetLastError());
  return 0;
}

}  // namespace dmc
