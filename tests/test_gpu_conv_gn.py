"""GPU: GroupNorm(+SiLU) of a convolution's output fused into that convolution's epilogue (csrc/conv_umma_kernel.cuh VAR_GN,
dmc_conv_desc.gn_*) -- the normalisation the NEXT layer of the reference applies (models/unet.py:35-36,51-52,80,238-239) and
the torch.cat of models/unet.py:284 (versions are channel slices of a wider tensor).  Reference of the op: plain PyTorch
fp32 conv + F.group_norm + F.silu on the same bf16-rounded operands.  Covered geometry classes (pixels per image vs rows per
CTA): several whole images per 128-row sub-tile (4x4, 8x8), one image per CTA (16x16 with two sub-tiles), one image across the
CTAs of a pair (16x16, 256-channel tile) and across several CTA groups (32x32: the self-resetting global counters), stride 2,
3x3 / 1x1, residual, conditioning, with and without the raw output, one or two versions with different group sizes, the
TMA-store and the per-thread-store epilogues, batch sizes that leave tiles partly / wholly out of range."""

import ctypes as C

import pytest
import torch
import torch.nn.functional as F

from diffusion_models_collection_b200 import _lib
from tests.gpu_util import Plan, nchw_f32, nhwc_bf16, pack3, rel_l2

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True)
def _no_tf32():
    a, b = torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    yield
    torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = a, b


def _rand(shape, seed, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(*shape, generator=g) * scale).cuda()


def _q(x):
    return x.to(torch.bfloat16).float()


def run_conv_gn(x, w, taps, *, stride=1, bias=None, cond=None, residual=None, versions=(), raw=True, runs=1):
    """x fp32 NCHW (bf16-rounded), w fp32 [Cout, Cin, k, k]; versions: list of dict(pitch, coff, gamma, beta, gsize, silu).
    Returns (raw bf16 NHWC or None, [version tensors bf16 NHWC of `pitch` channels], stats)."""
    dev = x.device
    B, cin, H, W = x.shape
    cout = w.shape[0]
    src = nhwc_bf16(x)
    wq = (pack3(w) if taps == 9 else w.reshape(cout, cin)).to(torch.bfloat16).contiguous()
    d = _lib.ConvDesc()
    d.nsrc = 1
    d.src[0], d.src_c[0], d.src_taps[0] = src.data_ptr(), cin, taps
    d.B, d.Hin, d.Win, d.stride, d.up_phase = B, H, W, stride, -1
    d.weight, d.Cout, d.Cout_pad, d.Ktot = wq.data_ptr(), cout, cout, wq.shape[1]
    d.bias = bias.data_ptr() if bias is not None else None
    if cond is not None:
        d.cond, d.cond_stride = cond.data_ptr(), cond.shape[1]
    Ho, Wo = H // stride, W // stride
    res = None
    if residual is not None:
        res = nhwc_bf16(residual)
        d.residual = res.data_ptr()
    out = torch.full((B, Ho, Wo, cout), float("nan"), device=dev, dtype=torch.bfloat16) if raw else None
    if raw:
        d.out_bf16 = out.data_ptr()
    slots = max(1, Ho * Wo // 32)
    st = torch.full((B, slots, cout // 8, 2), float("nan"), device=dev)
    d.stats, d.stats_slots = st.data_ptr(), slots
    outs = []
    d.gn_nver, d.gn_eps = len(versions), 1e-5
    for i, v in enumerate(versions):
        t = torch.full((B, Ho, Wo, v["pitch"]), 7.0, device=dev, dtype=torch.bfloat16)  # 7 = "untouched" marker
        outs.append(t)
        d.gn_out[i], d.gn_pitch[i], d.gn_coff[i] = t.data_ptr(), v["pitch"], v["coff"]
        d.gn_gamma[i], d.gn_beta[i] = v["gamma"].data_ptr(), v["beta"].data_ptr()
        d.gn_gsize[i], d.gn_silu[i] = v["gsize"], int(v["silu"])
    cnt = torch.zeros(2 * B * max(1, cout // 32), dtype=torch.int32, device=dev)
    d.gn_counters = cnt.data_ptr()
    p = Plan()
    p.add("conv", d)
    for _ in range(runs):
        p.run()
    assert int(cnt.abs().sum()) == 0, "image counters must reset themselves"
    p.keep = [src, wq, res, cnt]
    return out, outs, st


def reference(x, w, taps, stride, bias, cond, residual, versions):
    y = F.conv2d(x, w, bias, stride=stride, padding=1 if taps == 9 else 0)
    if cond is not None:
        y = y + cond[:, :w.shape[0], None, None]
    if residual is not None:
        y = y + residual
    refs = []
    for v in versions:
        C = y.shape[1]
        n = F.group_norm(y, C // v["gsize"], v["gamma"], v["beta"], eps=1e-5)
        refs.append(F.silu(n) if v["silu"] else n)
    return y, refs


def _versions(cout, spec, seed):
    out = []
    for i, (pitch, coff, gsize, silu) in enumerate(spec):
        out.append(dict(pitch=pitch, coff=coff, gsize=gsize, silu=silu, gamma=1.0 + _rand((cout,), seed + 2 * i, 0.2),
                        beta=_rand((cout,), seed + 2 * i + 1, 0.1)))
    return out


# (cin, cout, H, stride, taps, B, residual, raw, versions[(pitch, coff, gsize, silu)], DMC_CONV_CG)
CASES = [
    # 4x4 / 8x8: whole images inside a 128-row sub-tile
    (256, 256, 4, 1, 9, 5, False, False, [(256, 0, 32, True)], "2"),
    (256, 256, 4, 1, 9, 16, True, True, [(256, 0, 32, True), (512, 256, 64, True)], "2"),
    (256, 256, 8, 1, 9, 3, True, True, [(256, 0, 32, False), (512, 0, 64, True)], "2"),
    (256, 256, 8, 1, 9, 2, False, False, [(256, 0, 32, True)], "1"),
    (256, 256, 16, 2, 9, 4, False, True, [(256, 0, 32, True), (512, 256, 64, True)], "2"),  # Downsample 16 -> 8
    # 16x16: one image = the two CTAs of a pair (BN 256) / the two sub-tiles of one CTA (BN 128)
    (256, 256, 16, 1, 9, 3, True, True, [(256, 0, 32, False)], "2"),
    (128, 256, 16, 1, 9, 2, False, False, [(256, 0, 32, True)], "2"),
    (256, 256, 16, 1, 1, 5, True, True, [(256, 0, 32, True), (512, 256, 64, True)], "2"),  # attention proj (1x1, resident weights)
    (256, 256, 16, 1, 9, 2, False, False, [(256, 0, 32, True)], "1"),
    (128, 128, 32, 2, 9, 3, False, True, [(128, 0, 16, True)], "2"),                       # Downsample 32 -> 16
    # 32x32: one image spans several CTA groups (global image counters)
    (128, 128, 32, 1, 9, 2, False, False, [(128, 0, 16, True)], "2"),
    (128, 128, 32, 1, 9, 5, True, True, [(128, 0, 16, True), (256, 128, 32, True)], "2"),
    (128, 128, 32, 1, 9, 3, True, True, [(128, 0, 16, True)], "1"),
    (256, 128, 32, 1, 9, 40, False, False, [(128, 0, 16, True)], "2"),                     # >= 148 CTAs: the configuration of the bench
    (128, 128, 32, 1, 9, 75, True, True, [(128, 0, 16, True), (256, 0, 32, True)], "2"),   # more image units than CTA groups: 2 iterations
    # 64x64 (the shipped custom-dataset image size): 4096 pixels per image = 16 CTAs, 128 statistics slots
    (128, 128, 64, 1, 9, 2, False, False, [(128, 0, 16, True)], "2"),
    (128, 128, 64, 1, 9, 3, True, True, [(128, 0, 16, True), (256, 128, 32, False)], "1"),
]


@pytest.mark.parametrize("ts", ["1", "0"])
@pytest.mark.parametrize("cin,cout,H,stride,taps,B,use_res,raw,vspec,pairs", CASES)
def test_conv_with_fused_groupnorm(cin, cout, H, stride, taps, B, use_res, raw, vspec, pairs, ts, monkeypatch):
    monkeypatch.setenv("DMC_CONV_CG", pairs)
    monkeypatch.setenv("DMC_CONV_TMA_STORE", "2" if ts == "1" else "0")
    x = _q(_rand((B, cin, H, H), 1))
    w = _q(_rand((cout, cin, 3, 3) if taps == 9 else (cout, cin, 1, 1), 2, (cin * taps) ** -0.5))
    bias = _rand((cout,), 3, 0.1)
    cond = _rand((B, cout + 8), 4, 0.3) if taps == 9 else None
    Ho = H // stride
    residual = _q(_rand((B, cout, Ho, Ho), 5)) if use_res else None
    versions = _versions(cout, vspec, 10)
    y, refs = reference(x, w, taps, stride, bias, cond, residual, versions)
    out, outs, st = run_conv_gn(x, w, taps, stride=stride, bias=bias, cond=cond, residual=residual, versions=versions, raw=raw,
                                runs=2)  # the second run proves the counters reset themselves
    if raw:
        assert rel_l2(nchw_f32(out), y) < 4e-3
    assert torch.isfinite(st).all()
    for v, t, r in zip(versions, outs, refs):
        got = nchw_f32(t[..., v["coff"]:v["coff"] + cout])
        assert torch.isfinite(got).all()
        err = rel_l2(got, r)
        assert err < 5e-3, (v["gsize"], v["coff"], err)
        # the rest of a wider tensor (the other source of the concat) is left alone
        other = torch.cat([t[..., :v["coff"]], t[..., v["coff"] + cout:]], dim=-1)
        assert bool((other == 7.0).all())
    # deterministic, batch invariant: an image alone gives the same bits as inside the batch
    if B >= 3:
        k = B - 2
        _, outs1, _ = run_conv_gn(x[k:k + 1], w, taps, stride=stride, bias=bias, cond=None if cond is None else cond[k:k + 1],
                                  residual=None if residual is None else residual[k:k + 1], versions=versions, raw=raw)
        for t, t1 in zip(outs, outs1):
            assert torch.equal(t[k], t1[0])


def test_fused_groupnorm_matches_the_stand_alone_pass():
    """same conv, GroupNorm + SiLU by the epilogue vs by gn_apply_kernel on the stored bf16 output: the fused path normalises
    the fp32 accumulator (no intermediate rounding), so the two agree to bf16 rounding and the fused one is closer to fp32"""
    B, C, H = 4, 256, 16
    x = _q(_rand((B, C, H, H), 1))
    w = _q(_rand((C, C, 3, 3), 2, (C * 9) ** -0.5))
    bias = _rand((C,), 3, 0.1)
    versions = _versions(C, [(C, 0, 32, True)], 20)
    y, refs = reference(x, w, 9, 1, bias, None, None, versions)
    out, outs, st = run_conv_gn(x, w, 9, bias=bias, versions=versions, raw=True)
    d = _lib.GnApplyDesc()
    d.nsrc = 1
    d.src[0], d.src_c[0], d.stats[0], d.stats_slots[0] = out.data_ptr(), C, st.data_ptr(), st.shape[1]
    d.B, d.HW, d.groups = B, H * H, 8
    d.gamma, d.beta, d.eps, d.silu = versions[0]["gamma"].data_ptr(), versions[0]["beta"].data_ptr(), 1e-5, 1
    sep = torch.empty((B, H, H, C), device="cuda", dtype=torch.bfloat16)
    d.out = sep.data_ptr()
    p = Plan()
    p.add("gn_apply", d)
    p.run()
    e_fused, e_sep = rel_l2(nchw_f32(outs[0]), refs[0]), rel_l2(nchw_f32(sep), refs[0])
    print(f"fused {e_fused:.3e} stand-alone {e_sep:.3e}")
    assert rel_l2(nchw_f32(outs[0]), nchw_f32(sep)) < 6e-3
    assert e_fused < 4e-3 and e_fused <= e_sep * 1.05


def test_gn_support_query_and_bad_requests():
    lib = _lib.load()
    assert lib.dmc_conv_gn_supported(2048, 32, 32, 128, 16) == 1
    assert lib.dmc_conv_gn_supported(2, 4, 4, 256, 64) == 1
    assert lib.dmc_conv_gn_supported(2, 2, 2, 256, 32) == 0      # 4 pixels per image: below the statistics granularity
    assert lib.dmc_conv_gn_supported(2, 32, 32, 96, 16) == 0     # no 64+-channel tile divides 96
    assert lib.dmc_conv_gn_supported(2, 32, 32, 128, 48) == 0    # group size not 16 / 32 / 64
    x = _q(_rand((2, 128, 8, 8), 1))
    w = _q(_rand((128, 128, 3, 3), 2, 0.03))
    v = _versions(128, [(128, 0, 16, True)], 5)
    v[0]["coff"] = 8  # slice offset not a multiple of the group size
    with pytest.raises(_lib.DmcError):
        run_conv_gn(x, w, 9, versions=v)


@pytest.mark.parametrize("pairs", ["1", "2"])
@pytest.mark.parametrize("C,H,B", [(256, 16, 3), (256, 16, 40), (256, 8, 5), (256, 4, 9), (128, 16, 2), (256, 32, 2)])
def test_groupnorm_applied_to_the_qkv_operand_is_bit_identical_to_the_stand_alone_pass(C, H, B, pairs, monkeypatch):
    """AttentionBlock norm -> qkv (models/unet.py:80-81,86-87): dmc_plan_add_gn_coeff + dmc_conv_desc.a_affine (two idle warps of
    the GEMM kernel rewrite the landed A tiles as bf16(x * scale + shift)) against gn_apply + the plain GEMM: the same bits"""
    monkeypatch.setenv("DMC_CONV_CG", pairs)
    lib = _lib.load()
    if not lib.dmc_conv_affine_supported(B, H, H, C, 3 * C):
        pytest.skip("this geometry does not run with resident weights")
    x = nhwc_bf16(_rand((B, C, H, H), 1))
    gamma, beta = 1.0 + _rand((C,), 2, 0.2), _rand((C,), 3, 0.1)
    w = (_rand((3 * C, C), 4, C ** -0.5)).to(torch.bfloat16).contiguous()
    bias = _rand((3 * C,), 5, 0.1)
    slots = (H * H + 127) // 128
    st = torch.full((B, slots, C // 8, 2), float("nan"), device="cuda")
    an = torch.full((B, H, H, C), float("nan"), device="cuda", dtype=torch.bfloat16)
    coeff = torch.full((B, C, 2), float("nan"), device="cuda")
    outs = [torch.full((B, H, H, 3 * C), float("nan"), device="cuda", dtype=torch.bfloat16) for _ in range(2)]
    gs = _lib.GnStatsDesc()
    gs.src, gs.B, gs.HW, gs.C, gs.stats = x.data_ptr(), B, H * H, C, st.data_ptr()
    ga = _lib.GnApplyDesc()
    ga.nsrc = 1
    ga.src[0], ga.src_c[0], ga.stats[0], ga.stats_slots[0] = x.data_ptr(), C, st.data_ptr(), slots
    ga.B, ga.HW, ga.groups, ga.gamma, ga.beta, ga.eps, ga.silu, ga.out = B, H * H, 8, gamma.data_ptr(), beta.data_ptr(), 1e-5, 0, an.data_ptr()
    gc = _lib.GnCoeffDesc()
    gc.stats, gc.stats_slots, gc.B, gc.HW, gc.C, gc.groups = st.data_ptr(), slots, B, H * H, C, 8
    gc.gamma, gc.beta, gc.eps, gc.out = gamma.data_ptr(), beta.data_ptr(), 1e-5, coeff.data_ptr()

    def conv(src, out, aff):
        d = _lib.ConvDesc()
        d.nsrc = 1
        d.src[0], d.src_c[0], d.src_taps[0] = src.data_ptr(), C, 1
        d.B, d.Hin, d.Win, d.stride, d.up_phase = B, H, H, 1, -1
        d.weight, d.Cout, d.Cout_pad, d.Ktot = w.data_ptr(), 3 * C, 3 * C, C
        d.bias, d.out_bf16 = bias.data_ptr(), out.data_ptr()
        if aff is not None:
            d.a_affine = aff.data_ptr()
        return d

    p = Plan()
    p.add("gn_stats", gs)
    p.add("gn_apply", ga)
    p.add("conv", conv(an, outs[0], None))
    p.add("gn_coeff", gc)
    p.add("conv", conv(x, outs[1], coeff))
    p.run()
    p.run()  # the ring / barrier phases survive a second launch
    assert torch.isfinite(outs[0].float()).all() and torch.isfinite(outs[1].float()).all()
    assert torch.equal(outs[0], outs[1])
    # and against plain PyTorch
    xf = nchw_f32(x)
    ref = F.conv2d(F.group_norm(xf, 8, gamma, beta, eps=1e-5).to(torch.bfloat16).float(), w.float()[:, :, None, None], bias)
    assert rel_l2(nchw_f32(outs[1]), ref) < 5e-3
