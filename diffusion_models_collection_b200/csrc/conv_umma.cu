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
#include "conv_umma_kernel.cuh"

#include <new>

namespace dmc {

// one translation unit per tile configuration (conv_inst_*.cu)
int launch_conv_256_1_2(const ConvPrepared* P, const ConvKParams& kp, cudaStream_t st);
int launch_conv_128_2_2(const ConvPrepared* P, const ConvKParams& kp, cudaStream_t st);
int launch_conv_256_1_1(const ConvPrepared* P, const ConvKParams& kp, cudaStream_t st);
int launch_conv_128_2_1(const ConvPrepared* P, const ConvKParams& kp, cudaStream_t st);
int launch_conv_128_1_1(const ConvPrepared* P, const ConvKParams& kp, cudaStream_t st);
int launch_conv_64_1_1(const ConvPrepared* P, const ConvKParams& kp, cudaStream_t st);
int launch_conv_32_1_1(const ConvPrepared* P, const ConvKParams& kp, cudaStream_t st);

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
static int encode_map(CUtensorMap* m, const void* base, int rank, const cuuint64_t* dims, const cuuint64_t* strides,
                      const cuuint32_t* box, const cuuint32_t* estr,
                      CUtensorMapDataType dtype = CU_TENSOR_MAP_DATA_TYPE_BFLOAT16) {
  EncodeTiledFn fn = encode_tiled_fn();
  DMC_REQUIRE(fn != nullptr, "conv: cuTensorMapEncodeTiled unavailable -- call dmc_init() on a CUDA 12+ driver");
  CUresult r = fn(m, dtype, static_cast<cuuint32_t>(rank), const_cast<void*>(base), dims,
                  strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  DMC_REQUIRE(r == CUDA_SUCCESS, "conv: cuTensorMapEncodeTiled failed with CUresult %d", static_cast<int>(r));
  return 0;
}

// Tile configuration (BN output channels x MT 128-pixel sub-tiles per CTA tile).  Bytes staged per MAC fall with
// MT * BN (the A tile is reused by BN channels, the B tile by MT * 128 pixels): prefer 256 TMEM columns per
// accumulator when that still yields at least one tile per SM.
struct TileCfg { int bn, mt, cg; };

// Fused GroupNorm epilogue: can tile configuration `c` run it for an output of P pixels per image, `cout` channels and
// statistics groups of at most `max_gsz` channels?  Every group must lie inside one n tile; an image spanning several CTAs
// needs whole CTA groups per image and all of them in one wave of the persistent grid.
static bool gn_cfg_ok(const TileCfg& c, int cout, int P, int max_gsz) {
  if (c.bn < 64 || cout % c.bn != 0 || c.bn % max_gsz != 0) return false;
  const int rows = TILE_M * c.mt;
  if (P <= rows) return rows % P == 0 && (P == 16 || P % 32 == 0);
  if (P % rows != 0) return false;
  const int ctas = P / rows;
  if (ctas % c.cg != 0) return false;
  return (ctas / c.cg) * (cout / c.bn) <= num_sms() / c.cg;
}

static TileCfg pick_cfg(int cout_pad, int m_tiles, int gn_P = 0, int gn_max_gsz = 0) {
  const int sms = num_sms();
  // DMC_CONV_CG (debug / tests): "1" never pairs SMs, "2" pairs them whenever the channel count allows
  const char* e = getenv("DMC_CONV_CG");
  const bool allow_pairs = !(e && e[0] == '1');
  const bool force_pairs = e && e[0] == '2';
  const TileCfg cands[7] = {{256, 1, 2}, {128, 2, 2}, {256, 1, 1}, {128, 2, 1}, {128, 1, 1}, {64, 1, 1}, {32, 1, 1}};
  auto ok = [&](const TileCfg& c) { return gn_P == 0 || gn_cfg_ok(c, cout_pad, gn_P, gn_max_gsz); };
  if (force_pairs)
    for (int i = 0; i < 2; ++i)
      if (cout_pad % cands[i].bn == 0 && ok(cands[i])) return cands[i];
  for (const TileCfg& c : cands) {
    if (cout_pad % c.bn != 0 || (c.cg == 2 && !allow_pairs) || !ok(c)) continue;
    const long long ctas = static_cast<long long>((m_tiles + c.mt * c.cg - 1) / (c.mt * c.cg)) * (cout_pad / c.bn) * c.cg;
    if (ctas >= sms) return c;
  }
  for (int i = 6; i >= 2; --i)
    if (cout_pad % cands[i].bn == 0 && ok(cands[i])) return cands[i];
  if (gn_P != 0) return {0, 0, 0};  // no configuration can fuse the GroupNorm of this output
  return {32, 1, 1};
}

// 128-pixel M tiles of a [B, Hout, Wout] output (box BW x BH x BNIMG, see conv_prepare); 0 when the shape does not tile
static int conv_m_tiles(int B, int Hout, int Wout, int* tiles_w_out) {
  const int BW = std::min(Wout, TILE_M);
  if (BW <= 0 || TILE_M % BW != 0 || Wout % BW != 0) return 0;
  const int BH = std::min(Hout, TILE_M / BW);
  if (BH <= 0 || (TILE_M / BW) % BH != 0 || Hout % BH != 0) return 0;
  const int BNIMG = TILE_M / (BW * BH);
  if (tiles_w_out) *tiles_w_out = Wout / BW;
  return ((B + BNIMG - 1) / BNIMG) * (Wout / BW) * (Hout / BH);
}

// resident weights (the whole K extent of one weight tile stays in shared memory): short-K GEMMs only
static bool bres_ok(const TileCfg& tc, int num_kb, int num_n_tiles) {
  const int b_tile = (tc.bn / tc.cg) * KB * 2;
  return num_kb <= 8 && num_kb * b_tile <= 112 * 1024 && num_n_tiles <= 12 && num_n_tiles <= (num_sms() / tc.cg);
}

bool conv_affine_supported(int B, int H, int W, int Cin, int Cout) {
  if (B <= 0 || H <= 0 || W <= 0 || Cin % KB != 0 || Cout % 32 != 0) return false;
  const int m_tiles = conv_m_tiles(B, H, W, nullptr);
  if (m_tiles == 0) return false;
  const TileCfg tc = pick_cfg(Cout, m_tiles);
  const char* e = getenv("DMC_CONV_BRES");
  if (e && e[0] == '0') return false;
  return bres_ok(tc, Cin / KB, Cout / tc.bn);
}

bool conv_gn_supported(int B, int Hout, int Wout, int Cout, int max_gsz) {
  if (B <= 0 || Hout <= 0 || Wout <= 0 || Cout % 32 != 0 || !(max_gsz == 16 || max_gsz == 32 || max_gsz == 64)) return false;
  int tiles_w = 0;
  const int m_tiles = conv_m_tiles(B, Hout, Wout, &tiles_w);
  if (m_tiles == 0 || tiles_w != 1) return false;
  const int P = Hout * Wout;
  if (!(P == 16 || (P >= 32 && P % 32 == 0))) return false;
  return pick_cfg(Cout, m_tiles, P, max_gsz).bn != 0;
}

int conv_prepare(const dmc_conv_desc& d, ConvPrepared** out) {
  DMC_REQUIRE(d.nsrc >= 1 && d.nsrc <= 3, "conv: nsrc=%d", d.nsrc);
  DMC_REQUIRE(d.stride == 1 || d.stride == 2, "conv: stride=%d", d.stride);
  DMC_REQUIRE(d.up_phase >= -1 && d.up_phase <= 3, "conv: up_phase=%d", d.up_phase);
  DMC_REQUIRE(d.B > 0 && d.Hin > 0 && d.Win > 0, "conv: empty input");
  DMC_REQUIRE(d.Hin % d.stride == 0 && d.Win % d.stride == 0, "conv: odd spatial size with stride 2");
  DMC_REQUIRE(d.weight && (d.out_bf16 || d.out_f32_nchw || d.out_f32_nhwc || d.gn_nver > 0), "conv: null weight/output");
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
  if (d.stats) DMC_REQUIRE(d.out_bf16 != nullptr || d.gn_nver > 0, "conv: stats need a bf16 output");

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

  // ---- fused GroupNorm request (dmc_conv_desc.gn_*) ----
  const bool gn = d.gn_nver != 0;
  const int gn_P = Hout * Wout;
  int gn_max_gsz = 0, gn_min_gsz = 64;
  if (gn) {
    bool okv = d.gn_nver >= 1 && d.gn_nver <= 2 && d.stats != nullptr && d.out_f32_nchw == nullptr && !d.act && !d.gate &&
               !d.residual_f32 && !d.out_f32_nhwc && !d.out_lo && !d.residual_lo && d.up_phase < 0 && d.Cout == d.Cout_pad &&
               kp.tiles_w == 1 && d.unpatch_p == 0;
    for (int v = 0; okv && v < d.gn_nver; ++v) {
      const int gs = d.gn_gsize[v];
      okv = d.gn_out[v] && d.gn_gamma[v] && d.gn_beta[v] && (gs == 16 || gs == 32 || gs == 64) && d.Cout % gs == 0 &&
            d.gn_coff[v] >= 0 && d.gn_coff[v] % gs == 0 && d.gn_pitch[v] >= d.gn_coff[v] + d.Cout && d.gn_pitch[v] % 8 == 0;
      gn_max_gsz = std::max(gn_max_gsz, gs);
      gn_min_gsz = std::min(gn_min_gsz, gs);
    }
    if (!okv) {
      delete P;
      DMC_REQUIRE(false, "conv: bad fused-GroupNorm request (needs a plain UNet convolution with stats, 1-2 versions, group "
                         "sizes 16/32/64 that divide Cout and the slice offset)");
    }
  }
  const TileCfg tc = pick_cfg(d.Cout_pad, kp.num_m_tiles, gn ? gn_P : 0, gn_max_gsz);
  if (tc.bn == 0) {
    delete P;
    DMC_REQUIRE(false, "conv: no tile configuration fuses the GroupNorm of a %d-pixel, %d-channel output (group size %d)", gn_P,
                d.Cout, gn_max_gsz);
  }
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
  kp.gn_nver = 0;
  int gn_tab_bytes = 0;
  if (gn) {
    const int rows = TILE_M * tc.mt;
    kp.gn_nver = d.gn_nver;
    kp.gn_P = gn_P;
    kp.gn_imgs = gn_P <= rows ? rows / gn_P : 1;
    kp.gn_ctas_per_img = gn_P <= rows ? 1 : gn_P / rows;
    kp.gn_tab_groups = BN / gn_min_gsz;
    kp.gn_eps = d.gn_eps;
    kp.gn_counters = d.gn_counters;
    {
      const char* e = getenv("DMC_GN_DEBUG");
      kp.gn_debug = (e && e[0]) ? atoi(e) : 0;
    }
    for (int v = 0; v < d.gn_nver; ++v) {
      kp.gn_out[v] = reinterpret_cast<__nv_bfloat16*>(d.gn_out[v]);
      kp.gn_pitch[v] = d.gn_pitch[v]; kp.gn_coff[v] = d.gn_coff[v]; kp.gn_gsize[v] = d.gn_gsize[v]; kp.gn_silu[v] = d.gn_silu[v];
      kp.gn_gamma[v] = d.gn_gamma[v]; kp.gn_beta[v] = d.gn_beta[v];
    }
    gn_tab_bytes = 2 /*versions*/ * kp.gn_imgs * kp.gn_tab_groups * 8;
    kp.gn_sc_smem = 0;
    if (kp.gn_ctas_per_img > 1 && d.gn_counters == nullptr) {
      delete P;
      DMC_REQUIRE(false, "conv: fused GroupNorm of a %d-pixel image spans %d CTAs: gn_counters is required", gn_P, kp.gn_ctas_per_img);
    }
  }
  // ---- kernel variant selection (every switch has an environment override for A/B measurements and tests) ----
  const bool split = d.out_lo != nullptr || d.residual_lo != nullptr;  // split-bf16 (hi, lo) output / residual pairs
  const int epi = d.out_f32_nchw ? 2 : ((d.gate || d.residual_f32 || d.out_f32_nhwc) ? 3 : (d.act ? 1 : (split ? 4 : 0)));
  if (split && !((epi == 4 || (epi == 1 && !d.residual_lo)) && d.out_bf16 && (!d.residual_lo || d.residual))) {
    delete P;
    DMC_REQUIRE(false, "conv: out_lo / residual_lo go with a bf16 NHWC output (and residual) of a plain convolution or linear");
  }
  if ((epi == 1 || epi == 3) && (d.cond || d.residual || d.stats)) {
    delete P;
    DMC_REQUIRE(false, "conv: the transformer epilogues (act / gate / fp32 stream) exclude cond, bf16 residual and stats");
  }
  if (epi == 3 && (d.act || !d.out_f32_nhwc || d.out_bf16)) {
    delete P;
    DMC_REQUIRE(false, "conv: the fp32-stream epilogue (gate / fp32 residual) writes out_f32_nhwc only and takes no activation");
  }
  auto env_flag = [](const char* name, int dflt) {
    const char* e = getenv(name);
    return (e && e[0]) ? atoi(e) : dflt;
  };
  const int b_tile = (BN / tc.cg) * KB * 2;
  const int a_tiles_bytes = tc.mt * A_STAGE_BYTES;
  // (1) resident weights: short-K GEMMs (1x1 convs, transformer linears) re-fetch the weight tile for every 128 pixels and
  //     are L2->SM bound; with the whole K extent of one weight tile resident only activations stream.
  kp.bres = 0;
  kp.b_region_bytes = 0;
  if (env_flag("DMC_CONV_BRES", 1) && epi != 2 && epi != 4 && bres_ok(tc, kp.num_kb, kp.num_n_tiles)) {
    kp.bres = 1;
    kp.b_region_bytes = kp.num_kb * b_tile;
  }
  if (gn && kp.bres && kp.gn_ctas_per_img > 1 && kp.num_n_tiles > 1) {  // the resident-weight schedule splits an image's n tiles
    kp.bres = 0;
    kp.b_region_bytes = 0;
  }
  // (2) row slabs for the 3x3 segment 0: one box serves the three vertical taps
  kp.slab = 0;
  kp.slab_bytes = 0;
  P->tmS = P->tmA[0];
  // the 2 x 2 windows of the Upsample phase convolutions (models/unet.py:118-120 as four stride-1 convolutions over the low-resolution
  // map) take the same route: one box of BH + 1 rows serves both vertical taps (DMC_CONV_SLAB_PHASE=0: regular steps, for A/B runs)
  const bool phase_slab = d.up_phase >= 0 && d.src_taps[0] == 4 && d.nsrc == 1 && env_flag("DMC_CONV_SLAB_PHASE", 1);
  kp.slab_nv = kp.slab_nh = 3;
  kp.slab_dh0 = kp.slab_dw0 = -1;
  if (phase_slab) {
    kp.slab_nv = kp.slab_nh = 2;
    kp.slab_dh0 = kp.dh[0][0];
    kp.slab_dw0 = kp.dw[0][0];
  }
  if (env_flag("DMC_CONV_SLAB", 1) && !kp.bres && (epi == 0 || epi == 2 || epi == 4) && (d.src_taps[0] == 9 ? d.up_phase < 0 : phase_slab) &&
      d.stride == 1 && BNIMG == 1 && kp.tiles_w == 1 && BW >= 8 && kp.tiles_h % tc.mt == 0) {
    const int rows = tc.mt * BH + kp.slab_nv - 1;
    const int C = d.src_c[0];
    cuuint64_t dims[4] = {static_cast<cuuint64_t>(C), static_cast<cuuint64_t>(d.Win), static_cast<cuuint64_t>(d.Hin),
                          static_cast<cuuint64_t>(d.B)};
    cuuint64_t strides[3] = {static_cast<cuuint64_t>(C) * 2, static_cast<cuuint64_t>(d.Win) * C * 2,
                             static_cast<cuuint64_t>(d.Hin) * d.Win * C * 2};
    cuuint32_t box[4] = {KB, static_cast<cuuint32_t>(BW), static_cast<cuuint32_t>(rows), 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    if (rows <= 256 && rows * BW * 128 <= 96 * 1024) {
      if (encode_map(&P->tmS, d.src[0], 4, dims, strides, box, estr) != 0) { delete P; return -1; }
      kp.slab = 1;
      kp.slab_bytes = rows * BW * 128;
    }
  }
  kp.a_bytes = kp.slab ? std::max(kp.slab_bytes, a_tiles_bytes) : a_tiles_bytes;
  kp.stage_bytes = kp.a_bytes + (kp.bres ? 0 : (kp.slab ? kp.slab_nv : 1) * b_tile);
  // (3) TMA-store epilogue: coalesced 32-row x 64-channel boxes instead of per-thread 16-byte stores.  Pays for itself
  //     when the K loop is short (the epilogue is then the bottleneck: measured -10 % .. -22 % up to 18-27 K blocks);
  //     longer K loops lose more from the 32 KB it takes out of the operand ring (+4 % .. +19 % at 36-48 K blocks).
  P->tmOut = P->tmB;
  P->tmRes = P->tmB;
  P->tmV[0] = P->tmV[1] = P->tmB;
  kp.tma_store = 0;
  kp.res_tma = 0;
  kp.store_bufs = 1;
  {
    const int ts = env_flag("DMC_CONV_TMA_STORE", 1);  // 0 never, 1 short-K (and every transformer linear), 2 always
    const bool short_k = kp.num_kb <= env_flag("DMC_CONV_TMA_STORE_MAX_KB", 30);
    const bool want = ts == 2 || (ts == 1 && (short_k || epi == 1 || epi == 3));
    const bool f32 = epi == 3;
    const int boxc = f32 ? 32 : 64;
    void* optr = f32 ? static_cast<void*>(d.out_f32_nhwc) : d.out_bf16;
    if (gn && optr == nullptr) optr = d.gn_out[0];  // no raw output: tmOut is a placeholder, never stored through
    bool gn_ts_ok = true;
    for (int v = 0; v < d.gn_nver; ++v) gn_ts_ok = gn_ts_ok && d.gn_pitch[v] % 8 == 0 && d.gn_coff[v] % 8 == 0;
    if (want && epi != 2 && epi != 4 && !split && optr != nullptr && d.Cout % boxc == 0 && BN >= 128 && gn_ts_ok) {
      const int qbw = std::min(BW, 32), qbh = std::min(BH, 32 / qbw), qbn = 32 / (qbw * qbh);
      const int os = kp.oscale;
      const size_t es = f32 ? 4 : 2;
      const CUtensorMapDataType dt = f32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
      cuuint64_t dims[4] = {static_cast<cuuint64_t>(d.Cout), static_cast<cuuint64_t>(Wout), static_cast<cuuint64_t>(Hout),
                            static_cast<cuuint64_t>(d.B)};
      cuuint64_t strides[3] = {static_cast<cuuint64_t>(os) * d.Cout * es, static_cast<cuuint64_t>(os) * kp.out_W * d.Cout * es,
                               static_cast<cuuint64_t>(kp.out_H) * kp.out_W * d.Cout * es};
      cuuint32_t box[4] = {static_cast<cuuint32_t>(boxc), static_cast<cuuint32_t>(qbw), static_cast<cuuint32_t>(qbh),
                           static_cast<cuuint32_t>(qbn)};
      cuuint32_t estr[4] = {1, 1, 1, 1};
      const char* base = reinterpret_cast<const char*>(optr) +
                         (static_cast<size_t>(kp.ooff_h) * kp.out_W + kp.ooff_w) * d.Cout * es;
      if (d.out_bf16 != nullptr || !gn) {
        if (encode_map(&P->tmOut, base, 4, dims, strides, box, estr, dt) != 0) { delete P; return -1; }
      }
      for (int v = 0; v < d.gn_nver; ++v) {  // the normalised versions: channel slices of [B, Hout, Wout, pitch] tensors
        const cuuint64_t pc = static_cast<cuuint64_t>(d.gn_pitch[v]);
        cuuint64_t vdims[4] = {pc, static_cast<cuuint64_t>(Wout), static_cast<cuuint64_t>(Hout), static_cast<cuuint64_t>(d.B)};
        cuuint64_t vstr[3] = {pc * 2, static_cast<cuuint64_t>(Wout) * pc * 2, static_cast<cuuint64_t>(Hout) * Wout * pc * 2};
        if (encode_map(&P->tmV[v], d.gn_out[v], 4, vdims, vstr, box, estr) != 0) { delete P; return -1; }
      }
      if (gn && d.out_bf16 == nullptr) P->tmOut = P->tmV[0];
      kp.tma_store = 1;
      kp.qbw = qbw;
      kp.qbh = qbh;
      // the residual (same shape as the output) comes in as boxes too: coalesced and asynchronous
      const void* rptr = f32 ? static_cast<const void*>(d.residual_f32) : d.residual;
      if (rptr != nullptr && os == 1 && env_flag("DMC_CONV_RES_TMA", 1) && (!gn || d.out_bf16 != nullptr)) {
        if (rptr == optr) P->tmRes = P->tmOut;
        else if (encode_map(&P->tmRes, rptr, 4, dims, strides, box, estr, dt) != 0) { delete P; return -1; }
        kp.res_tma = 1;
      }
    }
  }
  {
    auto plan = [&](int bufs, int* nst) {
      const int store_bytes = kp.tma_store ? bufs * (EPI_THREADS / 32) * 4096 : 0;
      const int fixed = 1024 /*alignment slack*/ + 512 /*barriers*/ + gn_tab_bytes + store_bytes + kp.b_region_bytes;
      *nst = std::min(std::min(MAX_NST, env_flag("DMC_CONV_NST", MAX_NST)), (SMEM_LIMIT - fixed) / kp.stage_bytes);
      return fixed;
    };
    int nst1 = 0, nst2 = 0;
    int fixed = plan(1, &nst1);
    if (gn && kp.gn_imgs == 1 && env_flag("DMC_GN_SC_SMEM", 1)) {
      // per-channel (scale, shift) of the tile's image in shared memory, computed once per tile instead of once per thread --
      // when it costs no ring stage
      const int sc_bytes = 2 * BN * 8;
      gn_tab_bytes += sc_bytes;
      int nst_sc = 0;
      const int fixed_sc = plan(1, &nst_sc);
      if (nst_sc == nst1) {
        kp.gn_sc_smem = 1;
        fixed = fixed_sc;
      } else {
        gn_tab_bytes -= sc_bytes;
      }
    }
    kp.nst = nst1;
    if (kp.tma_store && env_flag("DMC_CONV_STORE_BUFS", 2) == 2) {  // a second staging buffer per warp, if the ring can spare it
      const int fixed2 = plan(2, &nst2);
      if (nst2 >= 4 || nst2 == nst1) {
        kp.store_bufs = 2;
        kp.nst = nst2;
        fixed = fixed2;
      }
    }
    if (kp.nst < 2) {
      delete P;
      DMC_REQUIRE(false, "conv: shared-memory plan does not fit (stage %d B, resident %d B)", kp.stage_bytes, kp.b_region_bytes);
    }
    P->smem = static_cast<size_t>(fixed) + static_cast<size_t>(kp.nst) * kp.stage_bytes;
  }
  P->var = (kp.slab ? VAR_SLAB : 0) | (kp.bres ? VAR_BRES : 0) | (kp.tma_store ? VAR_TS : 0) | (epi << VAR_EPI_SHIFT) |
           (gn ? VAR_GN : 0);
  kp.Cout = d.Cout;
  kp.bias = d.bias; kp.cond = d.cond; kp.cond_stride = d.cond_stride;
  kp.prefetch_cond = env_flag("DMC_CONV_PREFETCH_COND", 1);
  kp.a_affine = reinterpret_cast<const float2*>(d.a_affine);
  kp.aff_C = d.src_c[0];
  if (d.a_affine != nullptr && !(kp.bres && !kp.slab && epi == 0 && d.nsrc == 1 && d.src_taps[0] == 1 && d.stride == 1 && !gn)) {
    delete P;
    DMC_REQUIRE(false, "conv: a_affine needs a plain 1x1 convolution (one source, stride 1) that runs with resident weights");
  }
  kp.residual = reinterpret_cast<const __nv_bfloat16*>(d.residual);
  kp.out = reinterpret_cast<__nv_bfloat16*>(d.out_bf16);
  kp.out_nchw = d.out_f32_nchw;
  kp.stats = d.stats;
  kp.act = d.act;
  kp.gate = d.gate; kp.gate_stride = d.gate_stride;
  kp.residual_f32 = d.residual_f32; kp.out_f32 = d.out_f32_nhwc;
  kp.unpatch_p = d.unpatch_p;
  kp.residual_lo = reinterpret_cast<const __nv_bfloat16*>(d.residual_lo);
  kp.out_lo = reinterpret_cast<__nv_bfloat16*>(d.out_lo);
  // the straight-line box loop of the epilogue: bf16 TMA stores, bias only (+ GELU), residual only as TMA boxes, no conditioning
  // rows, no split-bf16 low parts (DMC_CONV_EPI_FAST=0: the generic chunk loop, for A/B runs)
  kp.epi_fast = (env_flag("DMC_CONV_EPI_FAST", 1) && kp.tma_store && !gn && (epi == 0 || epi == 1) && d.cond == nullptr &&
                 d.out_lo == nullptr && d.residual_lo == nullptr && d.out_bf16 != nullptr &&
                 (d.residual == nullptr || kp.res_tma)) ? 1 : 0;
  // transformer residual GEMMs (out_proj, fc2): out = res + gate * (acc + bias) in fp32, residual and output as TMA boxes
  if (env_flag("DMC_CONV_EPI_FAST", 1) && kp.tma_store && epi == 3 && kp.res_tma && d.residual_f32 != nullptr &&
      d.out_f32_nhwc != nullptr && d.out_bf16 == nullptr)
    kp.epi_fast = 1;
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
    int groups = std::min(group_tiles, num_sms() / tc.cg);
    if (gn && kp.gn_ctas_per_img > 1) {
      // the CTA groups that hold the pieces of one image (all its n tiles) must run in the same iteration of the tile loop:
      // the grid is a whole number of such units (group_tiles = images x unit by construction)
      const int unit = (kp.gn_ctas_per_img / tc.cg) * kp.num_n_tiles;
      groups = groups / unit * unit;
    }
    P->grid = tc.cg * groups;
  }
  *out = P;
  return 0;
}

void conv_release(ConvPrepared* p) { delete p; }
int launch_conv(const dmc_conv_desc& d, const ConvPrepared* P, cudaStream_t st) {
  ConvKParams kp = P->kp;
  kp.out_nchw = d.out_f32_nchw;  // the only re-bindable pointer (dmc_plan_rebind which=2)
  if (P->CG == 2) return P->BN == 256 ? launch_conv_256_1_2(P, kp, st) : launch_conv_128_2_2(P, kp, st);
  if (P->MT == 2) return launch_conv_128_2_1(P, kp, st);
  switch (P->BN) {
    case 256: return launch_conv_256_1_1(P, kp, st);
    case 128: return launch_conv_128_1_1(P, kp, st);
    case 64: return launch_conv_64_1_1(P, kp, st);
    default: return launch_conv_32_1_1(P, kp, st);
  }
}

}  // namespace dmc
