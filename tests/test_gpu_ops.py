"""GPU: every kernel of the UNet forward through the C ABI (one-op plans) vs a plain PyTorch fp32 reference of the
same op on the same bf16-rounded operands.  Tolerances: outputs are bf16 (rel. rounding 2^-9 = 2e-3), accumulation
is fp32 -> per-op relative L2 < 4e-3 (fp32 outputs < 1e-4); the tcgen05 kernel must also agree with the CUDA-core
debug implementation of the same contract (impl=1)."""

import pytest
import torch
import torch.nn.functional as F

from tests.gpu_util import Plan, nchw_f32, nhwc_bf16, pack3, rel_l2, run_conv

pytestmark = pytest.mark.gpu

TOL_BF16 = 4e-3


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


def _q(x):  # bf16 rounding, back in fp32
    return x.to(torch.bfloat16).float()


# (Cin list, Cout, H, stride, B)  -- SURVEY.md A.2 shape classes (3x3)
CONV3 = [
    ([128], 128, 32, 1, 2), ([128], 128, 32, 2, 3), ([128], 256, 16, 1, 2), ([256], 256, 16, 1, 1),
    ([256], 256, 16, 2, 5), ([256], 256, 8, 1, 3), ([256], 256, 8, 2, 9), ([256], 256, 4, 1, 5), ([256], 256, 4, 1, 16),
    ([256, 256], 256, 4, 1, 3), ([256, 256], 256, 8, 1, 2), ([256, 256], 256, 16, 1, 1), ([256, 128], 256, 16, 1, 2),
    ([256, 128], 128, 32, 1, 1), ([128, 128], 128, 32, 1, 2), ([256], 256, 32, 1, 1), ([64], 64, 32, 1, 2),
    ([64], 128, 16, 1, 1), ([128, 64], 64, 32, 1, 1),
]


@pytest.mark.parametrize("pairs", ["1", "2"])
@pytest.mark.parametrize("cins,cout,H,stride,B", CONV3)
def test_conv3x3_classes(cins, cout, H, stride, B, pairs, monkeypatch):
    """pairs = "2": force the cta_group::2 kernel (two SMs per 256-pixel tile) wherever Cout allows it; "1": never"""
    monkeypatch.setenv("DMC_CONV_CG", pairs)
    cin = sum(cins)
    x = _q(_rand((B, cin, H, H), 1))
    w = _q(_rand((cout, cin, 3, 3), 2, (cin * 9) ** -0.5))
    bias = _rand((cout,), 3, 0.1)
    cond = _rand((B, cout + 8), 4, 0.2)
    ref = F.conv2d(x, w, bias, stride=stride, padding=1) + cond[:, :cout, None, None]
    # the GEMM K order is source by source: split the channel range like the concat does
    srcs, wparts, c0 = [], [], 0
    for c in cins:
        srcs.append(nhwc_bf16(x[:, c0:c0 + c]))
        wparts.append(pack3(w[:, c0:c0 + c]))
        c0 += c
    wmat = torch.cat(wparts, dim=1)
    out, st = run_conv(srcs, [9] * len(cins), wmat, cout, stride=stride, bias=bias, cond=cond, stats=True)
    got = nchw_f32(out)
    assert torch.isfinite(got).all()
    assert rel_l2(got, ref) < TOL_BF16
    dbg, _ = run_conv(srcs, [9] * len(cins), wmat, cout, stride=stride, bias=bias, cond=cond, impl=1)
    assert rel_l2(got, nchw_f32(dbg)) < 2.5e-3
    # GroupNorm partial sums of the fp32 (pre-rounding) output per (image, 8-channel block): every slot written once
    assert torch.isfinite(st).all()
    tot = st.sum(dim=1)
    want_s = ref.reshape(B, cout // 8, 8, -1).sum(dim=(2, 3))
    want_ss = (ref * ref).reshape(B, cout // 8, 8, -1).sum(dim=(2, 3))
    assert rel_l2(tot[..., 1], want_ss) < 1e-3
    assert float((tot[..., 0] - want_s).abs().max()) < 2e-2 * float(want_ss.sqrt().max())
    # deterministic: a second run gives bit-identical statistics and outputs (no atomics)
    out2, st2 = run_conv(srcs, [9] * len(cins), wmat, cout, stride=stride, bias=bias, cond=cond, stats=True)
    assert torch.equal(out, out2) and torch.equal(st, st2)


@pytest.mark.parametrize("cin,cout,H,B", [(256, 768, 16, 2), (256, 768, 8, 3), (256, 768, 4, 5), (256, 256, 16, 1),
                                          (256, 256, 4, 9), (128, 256, 16, 2), (64, 192, 16, 3)])
def test_conv1x1_with_residual(cin, cout, H, B, monkeypatch):
    monkeypatch.setenv("DMC_CONV_CG", "2" if B % 2 else "1")
    x = _q(_rand((B, cin, H, H), 5))
    w = _q(_rand((cout, cin, 1, 1), 6, cin ** -0.5))
    bias = _rand((cout,), 7, 0.1)
    res = _q(_rand((B, cout, H, H), 8))
    ref = F.conv2d(x, w, bias) + res
    out, _ = run_conv([nhwc_bf16(x)], [1], w.reshape(cout, cin), cout, bias=bias, residual=nhwc_bf16(res))
    assert rel_l2(nchw_f32(out), ref) < TOL_BF16


@pytest.mark.parametrize("cins,cout,H,B", [([128], 256, 16, 2), ([256, 256], 256, 4, 3), ([256, 128], 128, 32, 1),
                                           ([256, 128], 256, 16, 2)])
def test_conv2_with_fused_shortcut(cins, cout, H, B, monkeypatch):
    monkeypatch.setenv("DMC_CONV_CG", "2")
    """conv2(3x3 over a2) + shortcut(1x1 over the raw concat inputs) as one GEMM with extra K columns"""
    a2 = _q(_rand((B, cout, H, H), 9))
    xs = [_q(_rand((B, c, H, H), 10 + i)) for i, c in enumerate(cins)]
    w2 = _q(_rand((cout, cout, 3, 3), 20, (cout * 9) ** -0.5))
    wsc = _q(_rand((cout, sum(cins), 1, 1), 21, sum(cins) ** -0.5))
    bias = _rand((cout,), 22, 0.1)
    ref = F.conv2d(a2, w2, None, padding=1) + F.conv2d(torch.cat(xs, 1), wsc) + bias[None, :, None, None]
    wmat = torch.cat([pack3(w2), wsc.reshape(cout, -1)], dim=1)
    out, _ = run_conv([nhwc_bf16(a2)] + [nhwc_bf16(v) for v in xs], [9] + [1] * len(xs), wmat, cout, bias=bias)
    assert rel_l2(nchw_f32(out), ref) < TOL_BF16


VARIANTS = [  # kernel-variant switches of conv_umma.cu, each against its opposite default
    {"DMC_CONV_SLAB": "0"}, {"DMC_CONV_BRES": "0"}, {"DMC_CONV_TMA_STORE": "0"}, {"DMC_CONV_TMA_STORE": "2"},
    {"DMC_CONV_SLAB": "0", "DMC_CONV_BRES": "0", "DMC_CONV_TMA_STORE": "0"}, {"DMC_CONV_CG": "1", "DMC_CONV_TMA_STORE": "2"},
    {"DMC_CONV_CG": "2", "DMC_CONV_TMA_STORE": "2"}, {"DMC_CONV_CG": "2", "DMC_CONV_TMA_STORE": "2", "DMC_CONV_RES_TMA": "0"},
    {"DMC_CONV_CG": "2", "DMC_CONV_TMA_STORE": "2", "DMC_CONV_STORE_BUFS": "1"},
]


@pytest.mark.parametrize("variant", VARIANTS, ids=lambda v: ",".join(f"{k[9:]}={x}" for k, x in v.items()))
@pytest.mark.parametrize("kind,cins,cout,H,B", [
    ("3x3", [128], 128, 32, 3), ("3x3", [256], 256, 16, 3), ("3x3", [128], 256, 16, 2), ("3x3+sc", [256, 128], 128, 32, 2),
    ("3x3+sc", [256, 256], 256, 16, 3), ("1x1", [256], 768, 16, 3), ("1x1", [256], 256, 16, 5), ("1x1", [384], 1152, 16, 2),
    ("3x3", [64], 64, 64, 1), ("3x3", [256], 256, 8, 4), ("3x3+res", [128], 128, 32, 3), ("3x3+res", [256], 256, 16, 2),
])
def test_conv_kernel_variants(kind, cins, cout, H, B, variant, monkeypatch):
    """row-slab loads (3x3), resident weights (short K), TMA-store epilogue: every switch on and off gives the reference
    result, bit-identically reproducible, with the GroupNorm partial sums intact; B is odd for some cases so that the last
    CTA pair works on a tile that is past the end of the tensor"""
    for k, v in variant.items():
        monkeypatch.setenv(k, v)
    bias = _rand((cout,), 3, 0.1)
    if kind == "1x1":
        cin = cins[0]
        x = _q(_rand((B, cin, H, H), 5))
        w = _q(_rand((cout, cin, 1, 1), 6, cin ** -0.5))
        res = _q(_rand((B, cout, H, H), 8))
        ref = F.conv2d(x, w, bias) + res
        args = ([nhwc_bf16(x)], [1], w.reshape(cout, cin), cout)
        kw = dict(bias=bias, residual=nhwc_bf16(res), stats=True)
    elif kind == "3x3":
        cin = cins[0]
        x = _q(_rand((B, cin, H, H), 1))
        w = _q(_rand((cout, cin, 3, 3), 2, (cin * 9) ** -0.5))
        cond = _rand((B, cout + 8), 4, 0.2)
        ref = F.conv2d(x, w, bias, padding=1) + cond[:, :cout, None, None]
        args = ([nhwc_bf16(x)], [9], pack3(w), cout)
        kw = dict(bias=bias, cond=cond, stats=True)
    elif kind == "3x3+res":  # ResidualBlock.conv2 with the identity shortcut (models/unet.py:72)
        cin = cins[0]
        x = _q(_rand((B, cin, H, H), 1))
        w = _q(_rand((cout, cin, 3, 3), 2, (cin * 9) ** -0.5))
        res = _q(_rand((B, cout, H, H), 8))
        ref = F.conv2d(x, w, bias, padding=1) + res
        args = ([nhwc_bf16(x)], [9], pack3(w), cout)
        kw = dict(bias=bias, residual=nhwc_bf16(res), stats=True)
    else:
        a2 = _q(_rand((B, cout, H, H), 9))
        xs = [_q(_rand((B, c, H, H), 10 + i)) for i, c in enumerate(cins)]
        w2 = _q(_rand((cout, cout, 3, 3), 20, (cout * 9) ** -0.5))
        wsc = _q(_rand((cout, sum(cins), 1, 1), 21, sum(cins) ** -0.5))
        ref = F.conv2d(a2, w2, None, padding=1) + F.conv2d(torch.cat(xs, 1), wsc) + bias[None, :, None, None]
        args = ([nhwc_bf16(a2)] + [nhwc_bf16(v) for v in xs], [9] + [1] * len(xs),
                torch.cat([pack3(w2), wsc.reshape(cout, -1)], dim=1), cout)
        kw = dict(bias=bias, stats=True)
    out, st = run_conv(*args, **kw)
    got = nchw_f32(out)
    assert torch.isfinite(got).all() and torch.isfinite(st).all()
    assert rel_l2(got, ref) < TOL_BF16
    tot = st.sum(dim=1)
    want_ss = (ref * ref).reshape(B, cout // 8, 8, -1).sum(dim=(2, 3))
    assert rel_l2(tot[..., 1], want_ss) < 1e-3
    out2, st2 = run_conv(*args, **kw)
    assert torch.equal(out, out2) and torch.equal(st, st2)


@pytest.mark.parametrize("kind,cin,cout,H,B,k", [("bias", 256, 768, 16, 3, 1), ("bias", 384, 1152, 16, 2, 1), ("res+stats", 256, 256, 16, 5, 1),
                                                 ("res+stats", 128, 128, 32, 3, 3), ("bias+stats", 256, 256, 8, 4, 1)])
def test_straight_line_epilogue_is_bit_identical_to_the_generic_loop(kind, cin, cout, H, B, k, monkeypatch):
    """ConvKParams.epi_fast (one 64-channel TMA box per iteration, no per-chunk decisions) against the generic chunk loop
    (DMC_CONV_EPI_FAST=0): same arithmetic in the same order -> the same bits, output and GroupNorm partial sums"""
    monkeypatch.setenv("DMC_CONV_TMA_STORE", "2")
    x = _q(_rand((B, cin, H, H), 5))
    w = _q(_rand((cout, cin, k, k), 6, (cin * k * k) ** -0.5))
    bias = _rand((cout,), 3, 0.1)
    res = nhwc_bf16(_q(_rand((B, cout, H, H), 8))) if kind.startswith("res") else None
    args = ([nhwc_bf16(x)], [k * k], pack3(w) if k == 3 else w.reshape(cout, cin), cout)
    kw = dict(bias=bias, residual=res, stats="stats" in kind)
    out = {}
    for fast in ("1", "0"):
        monkeypatch.setenv("DMC_CONV_EPI_FAST", fast)
        out[fast] = run_conv(*args, **kw)
    assert torch.equal(out["1"][0], out["0"][0])
    if "stats" in kind:
        assert torch.equal(out["1"][1], out["0"][1])
    ref = F.conv2d(x, w, bias, padding=k // 2)
    if res is not None:
        ref = ref + nchw_f32(res)
    assert rel_l2(nchw_f32(out["1"][0]), ref) < TOL_BF16


@pytest.mark.parametrize("B", [1, 3, 8])
def test_head_conv_fp32_nchw(B):
    x = _q(_rand((B, 128, 32, 32), 30))
    w = _q(_rand((3, 128, 3, 3), 31, (128 * 9) ** -0.5))
    bias = _rand((3,), 32, 0.1)
    ref = F.conv2d(x, w, bias, padding=1)
    wmat = torch.cat([pack3(w), torch.zeros(29, 128 * 9, device="cuda")], dim=0)
    out, _ = run_conv([nhwc_bf16(x)], [9], wmat, 3, bias=bias, out_nchw=True)
    assert rel_l2(out, ref) < 1e-4


@pytest.mark.parametrize("slab", ["1", "0"])
@pytest.mark.parametrize("C,H,B", [(256, 4, 3), (256, 8, 2), (128, 16, 2), (256, 16, 3), (256, 16, 80)])
def test_upsample_phase_convs(C, H, B, slab, monkeypatch):
    """nearest-2x + conv3x3 (models/unet.py:118-120) == four 2x2 phase convolutions on the low-res tensor; with row slabs (one box
    of BH + 1 rows serves both vertical taps of the 2 x 2 window; 16x16 maps) and with regular steps"""
    monkeypatch.setenv("DMC_CONV_SLAB_PHASE", slab)
    x = _q(_rand((B, C, H, H), 40))
    w = _rand((C, C, 3, 3), 41, (C * 9) ** -0.5)
    bias = _rand((C,), 42, 0.1)
    from diffusion_models_collection_b200.models.unet import phase_weights

    out = torch.full((B, 2 * H, 2 * H, C), float("nan"), device="cuda", dtype=torch.bfloat16)
    wq = []
    for ph in range(4):
        wm = phase_weights(w, ph)  # [C, 4*C] fp32, rounded to bf16 by the packer
        wq.append(wm.to(torch.bfloat16).float())
        o, st = run_conv([nhwc_bf16(x)], [4], wm, C, bias=bias, up_phase=ph, out_tensor=out, stats=True)
        assert torch.isfinite(st[:, ph * (st.shape[1] // 4):(ph + 1) * (st.shape[1] // 4)]).all()
    got = nchw_f32(out)
    assert torch.isfinite(got).all()
    # reference with the same (phase-summed, then bf16-rounded) weights: rebuild per-phase dense conv in fp32
    up = F.interpolate(x, scale_factor=2, mode="nearest")
    ref_exact = F.conv2d(up, w, bias, padding=1)
    assert rel_l2(got, ref_exact) < 6e-3  # bf16 rounding of the summed taps instead of each tap


def test_stem_conv():
    B = 5
    x = _rand((3, 3, 32, 32), 50)  # x_batch = 3 < B: images n read x[n % 3] (CFG halves share x)
    w = _rand((128, 3, 3, 3), 51, 27 ** -0.5)
    bias = _rand((128,), 52, 0.1)
    from diffusion_models_collection_b200 import _lib

    out = torch.full((B, 32, 32, 128), float("nan"), device="cuda", dtype=torch.bfloat16)
    d = _lib.StemDesc()
    d.x, d.x_batch, d.B, d.Cin, d.H, d.W, d.Cout = x.data_ptr(), 3, B, 3, 32, 32, 128
    d.weight, d.bias, d.out = w.data_ptr(), bias.data_ptr(), out.data_ptr()
    p = Plan()
    p.add("stem", d)
    p.run()
    ref = F.conv2d(x[torch.arange(B) % 3], w, bias, padding=1)
    assert rel_l2(nchw_f32(out), ref) < TOL_BF16


@pytest.mark.parametrize("cs,H,B,silu", [([128], 32, 2, 1), ([256], 16, 3, 0), ([256, 256], 4, 5, 1), ([256, 128], 16, 2, 1),
                                         ([256, 128], 32, 1, 1), ([256], 8, 2, 1), ([128, 64], 32, 2, 1)])
def test_groupnorm_stats_apply_concat(cs, H, B, silu):
    """GroupNorm(8, C)(+SiLU) over the concat of two tensors; 384 = 256 + 128 has a group straddling the boundary"""
    from diffusion_models_collection_b200 import _lib

    C_ = sum(cs)
    xs = [_q(_rand((B, c, H, H), 60 + i, 1.5) + 0.3) for i, c in enumerate(cs)]
    gamma, beta = 1 + 0.2 * _rand((C_,), 70), 0.1 * _rand((C_,), 71)
    ref = F.group_norm(torch.cat(xs, 1), 8, gamma, beta, eps=1e-5)
    if silu:
        ref = F.silu(ref)
    srcs = [nhwc_bf16(v) for v in xs]
    slots = (H * H + 127) // 128
    stats = [torch.full((B, slots, c // 8, 2), float("nan"), device="cuda") for c in cs]
    out = torch.full((B, H, H, C_), float("nan"), device="cuda", dtype=torch.bfloat16)
    p = Plan()
    for s, st, c in zip(srcs, stats, cs):
        d = _lib.GnStatsDesc()
        d.src, d.B, d.HW, d.C, d.stats = s.data_ptr(), B, H * H, c, st.data_ptr()
        p.add("gn_stats", d)
    d = _lib.GnApplyDesc()
    d.nsrc = len(cs)
    for i in range(len(cs)):
        d.src[i], d.src_c[i], d.stats[i], d.stats_slots[i] = srcs[i].data_ptr(), cs[i], stats[i].data_ptr(), slots
    d.B, d.HW, d.groups, d.gamma, d.beta, d.eps, d.silu, d.out = B, H * H, 8, gamma.data_ptr(), beta.data_ptr(), 1e-5, silu, out.data_ptr()
    p.add("gn_apply", d)
    p.run()
    assert rel_l2(nchw_f32(out), ref) < TOL_BF16


@pytest.mark.parametrize("impl", [0, 1])
@pytest.mark.parametrize("L,heads,hd,B", [(256, 4, 64, 2), (64, 4, 64, 3), (16, 4, 64, 5), (256, 6, 64, 2), (1024, 6, 64, 1),
                                          (256, 2, 64, 1), (16, 4, 64, 3), (64, 4, 64, 1), (128, 2, 64, 3), (32, 1, 64, 7),
                                          (256, 4, 64, 37), (512, 2, 64, 3), (1024, 6, 64, 2), (2048, 1, 64, 1),
                                          (256, 4, 64, 80), (256, 6, 64, 75), (256, 1, 64, 300)])
def test_attention(L, heads, hd, B, impl):
    from diffusion_models_collection_b200 import _lib

    C_ = heads * hd
    qkv = _q(_rand((B, L, 3 * C_), 80))
    out = torch.full((B, L, C_), float("nan"), device="cuda", dtype=torch.bfloat16)
    d = _lib.AttnDesc()
    keep = qkv.to(torch.bfloat16).contiguous()
    d.qkv, d.out, d.B, d.L, d.heads, d.C, d.impl = keep.data_ptr(), out.data_ptr(), B, L, heads, C_, impl
    p = Plan()
    p.add("attention", d)
    p.run()
    q, k, v = (qkv[..., i * C_:(i + 1) * C_].reshape(B, L, heads, hd).transpose(1, 2) for i in range(3))
    a = torch.softmax(q @ k.transpose(-1, -2) / hd ** 0.5, dim=-1)
    ref = (a @ v).transpose(1, 2).reshape(B, L, C_)
    assert rel_l2(out.float(), ref) < 6e-3


@pytest.mark.parametrize("heads,B", [(4, 3), (6, 40)])
def test_attention_kernel_generations_agree(heads, B, monkeypatch):
    """L = 256: P in tensor memory (default), the ping-pong kernel with P in shared memory (DMC_ATTN_PP=1) and the round-1
    one-warpgroup-per-tile kernel (DMC_ATTN_PP=0) compute the same softmax(QK^T/8)V; the first two share the summation order of the
    row sums and differ only in where P lives -> identical bits"""
    from diffusion_models_collection_b200 import _lib

    C_ = heads * 64
    qkv = _q(_rand((B, 256, 3 * C_), 81)).to(torch.bfloat16).contiguous()
    outs = {}
    for mode in ("2", "1", "0"):
        monkeypatch.setenv("DMC_ATTN_PP", mode)
        out = torch.full((B, 256, C_), float("nan"), device="cuda", dtype=torch.bfloat16)
        d = _lib.AttnDesc()
        d.qkv, d.out, d.B, d.L, d.heads, d.C, d.impl = qkv.data_ptr(), out.data_ptr(), B, 256, heads, C_, 0
        p = Plan()
        p.add("attention", d)
        p.run()
        outs[mode] = out
    assert torch.equal(outs["2"], outs["1"])
    assert rel_l2(outs["2"].float(), outs["0"].float()) < 2e-3


@pytest.mark.parametrize("uniform_t,has_y", [(1, True), (0, True), (0, False), (1, False)])
def test_conditioning_table(uniform_t, has_y):
    """time_embed + all time_mlp/label_proj projections vs fp32 torch (models/unet.py:18-25,40-48,167-172,256-260)"""
    import math

    from diffusion_models_collection_b200 import _lib

    B, half, temb, ncols, ncls = 6, 64, 512, 1024, 10
    t = torch.tensor([500] * B if uniform_t else [0, 1, 20, 500, 979, 999]).cuda()
    y = torch.tensor([0, 1, 5, 10, 11, 99]).cuda()
    w1, b1 = _rand((temb, 2 * half), 90, (2 * half) ** -0.5), _rand((temb,), 91, 0.1)
    w2, b2 = _rand((temb, temb), 92, temb ** -0.5), _rand((temb,), 93, 0.1)
    wt, bt = _rand((ncols, temb), 94, temb ** -0.5), _rand((ncols,), 95, 0.1)
    emb = _rand((ncls + 1, temb), 96)
    emb[0] = 0
    wy = _rand((ncols, temb), 97, temb ** -0.5)
    freqs = torch.exp(torch.arange(half, device="cuda") * -(math.log(10000) / (half - 1))).float()
    arg = t[:, None] * freqs[None]
    e = torch.cat([arg.sin(), arg.cos()], -1)
    temb_v = F.linear(F.silu(F.linear(e, w1, b1)), w2, b2)
    ref = F.linear(F.silu(temb_v), wt, bt)
    ytab = (F.silu(emb) @ wy.t()).contiguous()
    if has_y:
        ref = ref + ytab[y.clamp(0, ncls)]
    R = 1 if uniform_t else B
    scratch = torch.empty(2 * R * temb + R * ncols, device="cuda")
    cond = torch.full((B, ncols), float("nan"), device="cuda")
    d = _lib.CondDesc()
    d.t, d.y, d.B, d.uniform_t, d.num_classes = t.data_ptr(), (y.data_ptr() if has_y else None), B, uniform_t, ncls
    d.half, d.temb, d.ncols, d.freqs = half, temb, ncols, freqs.data_ptr()
    d.w1, d.b1, d.w2, d.b2 = w1.data_ptr(), b1.data_ptr(), w2.data_ptr(), b2.data_ptr()
    d.wt_all, d.bt_all, d.ytab = wt.data_ptr(), bt.data_ptr(), (ytab.data_ptr() if has_y else None)
    d.scratch, d.cond = scratch.data_ptr(), cond.data_ptr()
    p = Plan()
    p.add("cond", d)
    p.run()
    assert rel_l2(cond, ref) < 2e-5
    if has_y:  # null label (row 0 is zero) adds exactly nothing
        d.y = None
        d.ytab = None
        cond2 = torch.empty_like(cond)
        d.cond = cond2.data_ptr()
        p2 = Plan()
        p2.add("cond", d)
        p2.run()
        assert torch.equal(cond2[0], cond[0])


def test_upsample_nearest():
    from diffusion_models_collection_b200 import _lib

    x = _q(_rand((3, 64, 8, 8), 99))
    src = nhwc_bf16(x)
    out = torch.empty((3, 16, 16, 64), device="cuda", dtype=torch.bfloat16)
    d = _lib.UpsampleDesc()
    d.src, d.out, d.B, d.H, d.W, d.C = src.data_ptr(), out.data_ptr(), 3, 8, 8, 64
    p = Plan()
    p.add("upsample", d)
    p.run()
    assert torch.equal(nchw_f32(out), F.interpolate(x, scale_factor=2, mode="nearest"))


def test_bad_arguments_are_reported_not_crashed():
    from diffusion_models_collection_b200 import _lib

    lib = _lib.load()
    d = _lib.ConvDesc()
    d.nsrc = 1
    p = Plan()
    with pytest.raises(_lib.DmcError):
        p.add("conv", d)
    assert b"conv" in lib.dmc_last_error()


def _split(x):
    hi = x.to(torch.bfloat16)
    return hi, (x - hi.float()).to(torch.bfloat16)


@pytest.mark.parametrize("cin,cout,H,B,k", [(128, 128, 32, 2, 3), (256, 256, 16, 3, 3), (256, 256, 8, 2, 3), (256, 768, 16, 2, 1),
                                            (64, 64, 32, 1, 3)])
def test_conv_split_bf16_three_products(cin, cout, H, B, k):
    """split-bf16 mode of one convolution: sources (hi, lo, hi) against [W_hi | W_hi | W_lo], residual and output as
    (hi, lo) pairs -> fp32-level accuracy (relative L2 < 2e-5 vs the fp32 convolution of the fp32 operands)"""
    x = _rand((B, cin, H, H), 1)
    w = _rand((cout, cin, k, k), 2, (cin * k * k) ** -0.5)
    bias = _rand((cout,), 3, 0.1)
    res = _rand((B, cout, H, H), 4)
    ref = F.conv2d(x, w, bias, padding=k // 2) + res
    xh, xl = _split(x.permute(0, 2, 3, 1).contiguous())
    rh, rl = _split(res.permute(0, 2, 3, 1).contiguous())
    wm = pack3(w) if k == 3 else w.reshape(cout, cin)
    wh = wm.to(torch.bfloat16)
    wl = (wm - wh.float()).to(torch.bfloat16)
    w3 = torch.cat([wh, wh, wl], dim=1).contiguous()
    from diffusion_models_collection_b200 import _lib
    d = _lib.ConvDesc()
    d.nsrc = 3
    for i, s in enumerate((xh, xl, xh)):
        d.src[i], d.src_c[i], d.src_taps[i] = s.data_ptr(), cin, k * k
    d.B, d.Hin, d.Win, d.stride, d.up_phase = B, H, H, 1, -1
    d.weight, d.Cout, d.Cout_pad, d.Ktot = w3.data_ptr(), cout, cout, w3.shape[1]
    d.bias, d.residual, d.residual_lo = bias.data_ptr(), rh.data_ptr(), rl.data_ptr()
    oh = torch.full((B, H, H, cout), float("nan"), device="cuda", dtype=torch.bfloat16)
    ol = torch.full_like(oh, float("nan"))
    st = torch.full((B, max(1, H * H // 32), cout // 8, 2), float("nan"), device="cuda")
    d.out_bf16, d.out_lo, d.stats, d.stats_slots = oh.data_ptr(), ol.data_ptr(), st.data_ptr(), st.shape[1]
    p = Plan()
    p.add("conv", d)
    p.run()
    got = nchw_f32(oh) + nchw_f32(ol)
    assert torch.isfinite(got).all() and torch.isfinite(st).all()
    assert rel_l2(got, ref) < 2e-5
    want_ss = (ref * ref).reshape(B, cout // 8, 8, -1).sum(dim=(2, 3))
    assert rel_l2(st.sum(dim=1)[..., 1], want_ss) < 1e-4


@pytest.mark.parametrize("C_,H,B,cout", [(128, 32, 3, 3), (64, 32, 2, 3), (128, 16, 2, 1), (64, 64, 1, 8), (128, 32, 1, 4), (128, 64, 1, 3)])
def test_fused_output_head(C_, H, B, cout):
    """GroupNorm(8) + SiLU + conv3x3 C -> cout + bias -> fp32 NCHW in one kernel vs PyTorch fp32 on the bf16-rounded input"""
    from diffusion_models_collection_b200 import _lib
    x = _q(_rand((B, C_, H, H), 1, 1.5) + 0.3)
    gamma, beta = _rand((C_,), 2, 0.3) + 1.0, _rand((C_,), 3, 0.2)
    w, bias = _rand((cout, C_, 3, 3), 4, (C_ * 9) ** -0.5), _rand((cout,), 5, 0.1)
    ref = F.conv2d(F.silu(F.group_norm(x, 8, gamma, beta, eps=1e-5)), w, bias, padding=1)
    src = nhwc_bf16(x)
    slots = (H * H + 127) // 128
    stats = torch.empty((B, slots, C_ // 8, 2), device="cuda")
    gs = _lib.GnStatsDesc()
    gs.src, gs.B, gs.HW, gs.C, gs.stats = src.data_ptr(), B, H * H, C_, stats.data_ptr()
    out = torch.full((B, cout, H, H), float("nan"), device="cuda")
    d = _lib.HeadDesc()
    d.src, d.stats, d.stats_slots, d.B, d.H, d.W, d.C, d.Cout, d.groups = src.data_ptr(), stats.data_ptr(), slots, B, H, H, C_, cout, 8
    d.gamma, d.beta, d.eps, d.weight, d.bias, d.out = gamma.data_ptr(), beta.data_ptr(), 1e-5, w.data_ptr(), bias.data_ptr(), out.data_ptr()
    wfrag = torch.empty(9 * (C_ // 16) * 256, dtype=torch.uint8, device="cuda")
    d.wfrag = wfrag.data_ptr()
    assert _lib.load().dmc_head_supported(d) == 1
    p = Plan()
    p.add("gn_stats", gs)
    p.add("head", d)
    p.run()
    assert torch.isfinite(out).all()
    assert rel_l2(out, ref) < TOL_BF16  # normalised activations and weights are rounded to bf16 for the tensor cores
    out2 = torch.empty_like(out)
    d.out = out2.data_ptr()
    p2 = Plan()
    p2.add("gn_stats", gs)
    p2.add("head", d)
    p2.run()
    assert torch.equal(out, out2)


@pytest.mark.parametrize("C,H,B,cout", [(128, 32, 3, 3), (128, 32, 1, 3), (64, 16, 2, 3), (128, 8, 5, 1)])
def test_head_as_1x1_gemm_plus_tap_gather(C, H, B, cout):
    """the output conv3x3 (C -> cout) as ONE 1x1 GEMM with 9 * cout (tap, cout) columns writing fp32 NHWC, followed by the
    9-tap gather of dmc_head_taps_desc (models/unet.py:240) vs F.conv2d on the same bf16-rounded operands"""
    from diffusion_models_collection_b200 import _lib

    x = _q(_rand((B, C, H, H), 31))
    w = _q(_rand((cout, C, 3, 3), 32, (C * 9) ** -0.5))
    bias = _rand((cout,), 33, 0.1)
    ref = F.conv2d(x, w, bias, padding=1)
    ypitch = (9 * cout + 31) // 32 * 32
    wt = torch.zeros(ypitch, C, device="cuda")
    wt[: 9 * cout] = w.permute(2, 3, 0, 1).reshape(9 * cout, C)
    wq = wt.to(torch.bfloat16).contiguous()
    src = nhwc_bf16(x)
    y = torch.full((B, H, H, ypitch), float("nan"), device="cuda")
    d = _lib.ConvDesc()
    d.nsrc = 1
    d.src[0], d.src_c[0], d.src_taps[0] = src.data_ptr(), C, 1
    d.B, d.Hin, d.Win, d.stride, d.up_phase = B, H, H, 1, -1
    d.weight, d.Cout, d.Cout_pad, d.Ktot = wq.data_ptr(), ypitch, ypitch, C
    d.out_f32_nhwc = y.data_ptr()
    out = torch.full((B, cout, H, H), float("nan"), device="cuda")
    t = _lib.HeadTapsDesc()
    t.y, t.B, t.H, t.W, t.Cout, t.ypitch = y.data_ptr(), B, H, H, cout, ypitch
    t.bias, t.out = bias.data_ptr(), out.data_ptr()
    p = Plan()
    p.add("conv", d)
    p.add("head_taps", t)
    p.run()
    assert torch.isfinite(out).all()
    assert rel_l2(out, ref) < 1e-4  # fp32 output of bf16 operands: only the accumulation order differs


@pytest.mark.parametrize("B,x_batch,H", [(3, 3, 32), (4, 2, 32), (2, 2, 16), (1, 1, 64)])
def test_stem_as_gathered_columns_plus_1x1_gemm(B, x_batch, H):
    """the input conv3x3 (3 -> 128) as dmc_stem_cols (27 (tap, channel) columns of the fp32 input as a bf16 (hi, lo) pair) +
    ONE 1x1 tcgen05 GEMM against [W | W | 0] (models/unet.py:188,263): the input keeps 16 mantissa bits, so the result matches
    F.conv2d on the fp32 input with bf16-rounded weights to the bf16 rounding of the output; image n reads x[n % x_batch] (CFG)"""
    from diffusion_models_collection_b200 import _lib

    cin, cout = 3, 128
    x = _rand((x_batch, cin, H, H), 41)  # NOT bf16-rounded: the (hi, lo) pair must carry it
    w = _q(_rand((cout, cin, 3, 3), 42, (cin * 9) ** -0.5))
    bias = _rand((cout,), 43, 0.1)
    ref = F.conv2d(x[[n % x_batch for n in range(B)]], w, bias, padding=1)
    cols = torch.full((B, H, H, 64), float("nan"), device="cuda", dtype=torch.bfloat16)
    s = _lib.StemColsDesc()
    s.x, s.x_batch, s.B, s.Cin, s.H, s.W, s.out = x.data_ptr(), x_batch, B, cin, H, H, cols.data_ptr()
    wp = pack3(w)
    wmat = torch.cat([wp, wp, wp.new_zeros(cout, 64 - 54)], dim=1).to(torch.bfloat16).contiguous()
    out = torch.full((B, H, H, cout), float("nan"), device="cuda", dtype=torch.bfloat16)
    slots = max(1, H * H // 32)
    st = torch.full((B, slots, cout // 8, 2), float("nan"), device="cuda")
    d = _lib.ConvDesc()
    d.nsrc = 1
    d.src[0], d.src_c[0], d.src_taps[0] = cols.data_ptr(), 64, 1
    d.B, d.Hin, d.Win, d.stride, d.up_phase = B, H, H, 1, -1
    d.weight, d.Cout, d.Cout_pad, d.Ktot = wmat.data_ptr(), cout, cout, 64
    d.bias, d.out_bf16 = bias.data_ptr(), out.data_ptr()
    d.stats, d.stats_slots = st.data_ptr(), slots
    p = Plan()
    p.add("stem_cols", s)
    p.add("conv", d)
    p.run()
    assert torch.isfinite(cols.float()).all() and float(cols[..., 54:].float().abs().max()) == 0.0
    # hi + lo reproduces the fp32 input to 2^-16 relative
    centre = (cols[..., 4 * cin: 5 * cin].float() + cols[..., 27 + 4 * cin: 27 + 5 * cin].float()).permute(0, 3, 1, 2)
    xin = x[[n % x_batch for n in range(B)]]
    assert float((centre - xin).abs().max()) <= 2.0 ** -15 * float(xin.abs().max())
    got = nchw_f32(out)
    assert rel_l2(got, ref) < 3e-3
    tot = st.sum(dim=1)
    want_ss = (ref * ref).reshape(B, cout // 8, 8, -1).sum(dim=(2, 3))
    assert rel_l2(tot[..., 1], want_ss) < 1e-3
