"""GPU: building blocks of the training step against PyTorch autograd in fp32 on the same
bf16-rounded operands.  The full step is not wired yet; these pin the kernels it will be made of."""

import ctypes as C

import pytest
import torch
import torch.nn.functional as F

from diffusion_models_collection_b200 import _lib
from tests.gpu_util import nhwc_bf16, pack3, rel_l2, run_conv

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


@pytest.mark.parametrize("cin,cout,H,B,k,stride", [
    (128, 128, 32, 2, 3, 1), (256, 256, 16, 3, 3, 1), (128, 256, 16, 2, 3, 1), (256, 256, 8, 5, 3, 1), (256, 256, 4, 7, 3, 1),
    (256, 768, 16, 2, 1, 1), (256, 256, 8, 3, 1, 1), (128, 128, 32, 2, 3, 2), (512, 256, 16, 1, 3, 1), (64, 128, 32, 1, 3, 1),
    (384, 128, 32, 2, 3, 1), (256, 128, 32, 1, 1, 1), (128, 128, 8, 130, 3, 1),
])
def test_conv_weight_gradient(cin, cout, H, B, k, stride):
    """dW of conv(x, W) for a random upstream gradient vs autograd (fp32 accumulation over up to B*H*W = 4096 pixels)"""
    lib = _lib.load()
    x = _q(_rand((B, cin, H, H), 1))
    w = _rand((cout, cin, k, k), 2, 0.05).requires_grad_(True)
    y = F.conv2d(x, w, None, stride=stride, padding=k // 2)
    dy = _q(_rand(tuple(y.shape), 3))
    (ref,) = torch.autograd.grad(y, w, dy)
    xs, dys = nhwc_bf16(x), nhwc_bf16(dy)
    d = _lib.WgradDesc()
    d.x, d.dy, d.B, d.Hin, d.Win, d.Cin, d.Cout, d.stride, d.taps = xs.data_ptr(), dys.data_ptr(), B, H, H, cin, cout, stride, k * k
    d.splits = _lib.check(lib.dmc_conv_wgrad_splits(C.byref(d)), "splits")
    partial = torch.full((d.splits, cout, k * k, cin), float("nan"), device="cuda")
    dw = torch.full((cout, cin, k, k), float("nan"), device="cuda")
    d.partial, d.dw, d.accumulate = partial.data_ptr(), dw.data_ptr(), 0
    _lib.check(lib.dmc_conv_wgrad(C.byref(d), _lib.stream_ptr()), "wgrad")
    torch.cuda.synchronize()
    assert torch.isfinite(dw).all()
    assert rel_l2(dw, ref) < 1e-5  # exact bf16 products, fp32 accumulation in a different order
    dw2 = dw.clone()
    d.dw, d.accumulate = dw2.data_ptr(), 1
    _lib.check(lib.dmc_conv_wgrad(C.byref(d), _lib.stream_ptr()), "wgrad accumulate")
    torch.cuda.synchronize()
    assert torch.equal(dw2, dw + dw)  # deterministic, accumulates on top


@pytest.mark.parametrize("cin,cout,H,B,k", [(128, 128, 32, 2, 3), (256, 128, 16, 2, 3), (256, 256, 8, 3, 3), (256, 768, 16, 1, 1)])
def test_conv_input_gradient_is_a_forward_conv_with_flipped_transposed_weights(cin, cout, H, B, k):
    """dX of a stride-1 conv = the SAME implicit-GEMM kernel over dY with W[co, ci, r, s] -> W'[ci, co, 2-r, 2-s]"""
    x = _q(_rand((B, cin, H, H), 1)).requires_grad_(True)
    w = _q(_rand((cout, cin, k, k), 2, (cin * k * k) ** -0.5))
    y = F.conv2d(x, w, None, padding=k // 2)
    dy = _q(_rand(tuple(y.shape), 3))
    (ref,) = torch.autograd.grad(y, x, dy)
    wt = w.flip(2, 3).permute(1, 0, 2, 3).contiguous()  # [cin, cout, k, k]
    wmat = pack3(wt) if k == 3 else wt.reshape(cin, cout)
    out, _ = run_conv([nhwc_bf16(dy)], [k * k], wmat, cin)
    got = out.float().permute(0, 3, 1, 2)
    assert rel_l2(got, ref) < 4e-3  # the result is stored in bf16


def _gn_forward(xs, gamma, beta, silu, drop_p=0.0, seed=0):
    """runs the forward gn_stats + gn_apply kernels; returns (out bf16 NHWC, stats list, slots)"""
    from tests.gpu_util import Plan
    B, H, W, _ = xs[0].shape
    HW = H * W
    slots = (HW + 127) // 128
    stats = []
    p = Plan()
    for x in xs:
        st = torch.empty((B, slots, x.shape[3] // 8, 2), device="cuda")
        g = _lib.GnStatsDesc()
        g.src, g.B, g.HW, g.C, g.stats = x.data_ptr(), B, HW, x.shape[3], st.data_ptr()
        p.add("gn_stats", g)
        stats.append(st)
    Ctot = sum(x.shape[3] for x in xs)
    out = torch.full((B, H, W, Ctot), float("nan"), device="cuda", dtype=torch.bfloat16)
    d = _lib.GnApplyDesc()
    d.nsrc = len(xs)
    for i, x in enumerate(xs):
        d.src[i], d.src_c[i], d.stats[i], d.stats_slots[i] = x.data_ptr(), x.shape[3], stats[i].data_ptr(), slots
    d.B, d.HW, d.groups, d.gamma, d.beta, d.eps, d.silu, d.out = B, HW, 8, gamma.data_ptr(), beta.data_ptr(), 1e-5, silu, out.data_ptr()
    d.drop_p, d.seed = drop_p, seed
    p.add("gn_apply", d)
    p.run()
    return out, stats, slots


@pytest.mark.parametrize("cs,H,B,silu", [([128], 32, 2, 1), ([256], 16, 3, 0), ([256, 256], 4, 5, 1), ([256, 128], 16, 2, 1),
                                         ([64], 8, 3, 1), ([128, 128], 32, 1, 1)])
def test_groupnorm_silu_backward(cs, H, B, silu):
    lib = _lib.load()
    C_ = sum(cs)
    xs = [_q(_rand((B, c, H, H), 10 + i, 1.5) + 0.3) for i, c in enumerate(cs)]
    gamma = (_rand((C_,), 2, 0.3) + 1.0).requires_grad_(True)
    beta = _rand((C_,), 3, 0.2).requires_grad_(True)
    xr = [x.clone().requires_grad_(True) for x in xs]
    y = F.group_norm(torch.cat(xr, 1), 8, gamma, beta, eps=1e-5)
    if silu:
        y = F.silu(y)
    dy = _q(_rand(tuple(y.shape), 4))
    refs = torch.autograd.grad(y, xr + [gamma, beta], dy)
    xn = [nhwc_bf16(x) for x in xs]
    _, stats, slots = _gn_forward(xn, gamma.detach(), beta.detach(), silu)
    dout = nhwc_bf16(dy)
    dsrc = [torch.full_like(x, float("nan")) for x in xn]
    dgamma, dbeta = torch.full((C_,), float("nan"), device="cuda"), torch.full((C_,), float("nan"), device="cuda")
    d = _lib.GnBwdDesc()
    d.nsrc = len(cs)
    for i in range(len(cs)):
        d.src[i], d.src_c[i], d.stats[i], d.stats_slots[i] = xn[i].data_ptr(), cs[i], stats[i].data_ptr(), slots
        d.dsrc[i], d.accumulate[i] = dsrc[i].data_ptr(), 0
    d.dout, d.B, d.HW, d.groups = dout.data_ptr(), B, H * H, 8
    d.gamma, d.beta, d.eps, d.silu = gamma.detach().data_ptr(), beta.detach().data_ptr(), 1e-5, silu
    scratch = torch.empty(int(lib.dmc_gn_backward_scratch(C.byref(d))), device="cuda")
    d.dgamma, d.dbeta, d.scratch = dgamma.data_ptr(), dbeta.data_ptr(), scratch.data_ptr()
    _lib.check(lib.dmc_gn_backward(C.byref(d), _lib.stream_ptr()), "gn_backward")
    torch.cuda.synchronize()
    for i in range(len(cs)):
        got = dsrc[i].float().permute(0, 3, 1, 2)
        assert rel_l2(got, refs[i]) < 5e-3, i  # bf16 gradient storage
    assert rel_l2(dgamma, refs[-2]) < 1e-4 and rel_l2(dbeta, refs[-1]) < 1e-4
    # accumulate mode adds on top
    for i in range(len(cs)):
        d.accumulate[i] = 1
    before = [t.clone() for t in dsrc]
    _lib.check(lib.dmc_gn_backward(C.byref(d), _lib.stream_ptr()), "gn_backward acc")
    torch.cuda.synchronize()
    for i in range(len(cs)):
        assert rel_l2(dsrc[i].float(), 2 * before[i].float()) < 5e-3


def test_dropout_mask_is_the_same_in_forward_and_backward():
    """the forward pass scales kept activations by 1/(1-p) and zeroes the rest; the backward pass regenerates the same
    counter-based mask: its gradients equal autograd's through  silu(gn(x)) * mask / (1-p)  with the mask read off the
    forward output"""
    lib = _lib.load()
    B, C_, H, p, seed = 2, 128, 16, 0.25, 1234
    xf = _q(_rand((B, C_, H, H), 1, 1.5))
    x = nhwc_bf16(xf)
    gamma, beta = _rand((C_,), 2, 0.3) + 1.0, _rand((C_,), 3, 0.2)
    out0, stats, slots = _gn_forward([x], gamma, beta, 1)
    out1, _, _ = _gn_forward([x], gamma, beta, 1, drop_p=p, seed=seed)
    kept = out1 != 0
    assert abs(float(kept.float().mean()) - (1 - p)) < 0.02
    assert rel_l2(out1.float()[kept], out0.float()[kept] / (1 - p)) < 5e-3
    mask = kept.float().permute(0, 3, 1, 2)
    xr = xf.clone().requires_grad_(True)
    g_, b_ = gamma.clone().requires_grad_(True), beta.clone().requires_grad_(True)
    y = F.silu(F.group_norm(xr, 8, g_, b_, eps=1e-5)) * mask / (1 - p)
    dy = _q(_rand((B, C_, H, H), 4))
    rx, rg, rb = torch.autograd.grad(y, [xr, g_, b_], dy)
    dout = nhwc_bf16(dy)
    dsrc = torch.empty_like(x)
    dg, db = torch.empty(C_, device="cuda"), torch.empty(C_, device="cuda")
    d = _lib.GnBwdDesc()
    d.nsrc = 1
    d.src[0], d.src_c[0], d.stats[0], d.stats_slots[0], d.dsrc[0] = x.data_ptr(), C_, stats[0].data_ptr(), slots, dsrc.data_ptr()
    d.dout, d.B, d.HW, d.groups = dout.data_ptr(), B, H * H, 8
    d.gamma, d.beta, d.eps, d.silu, d.drop_p, d.seed = gamma.data_ptr(), beta.data_ptr(), 1e-5, 1, p, seed
    scratch = torch.empty(int(lib.dmc_gn_backward_scratch(C.byref(d))), device="cuda")
    d.dgamma, d.dbeta, d.scratch = dg.data_ptr(), db.data_ptr(), scratch.data_ptr()
    _lib.check(lib.dmc_gn_backward(C.byref(d), _lib.stream_ptr()), "gn_backward")
    torch.cuda.synchronize()
    assert rel_l2(dsrc.float().permute(0, 3, 1, 2), rx) < 5e-3
    assert rel_l2(dg, rg) < 1e-4 and rel_l2(db, rb) < 1e-4


@pytest.mark.parametrize("L,heads,B", [(256, 4, 2), (64, 4, 3), (16, 4, 2), (128, 2, 1), (100, 4, 2), (200, 1, 1)])
def test_attention_backward(L, heads, B):
    lib = _lib.load()
    hd, C_ = 64, heads * 64
    qkv = _q(_rand((B, L, 3 * C_), 80)).requires_grad_(True)
    q, k, v = (qkv[..., i * C_:(i + 1) * C_].reshape(B, L, heads, hd).transpose(1, 2) for i in range(3))
    a = torch.softmax(q @ k.transpose(-1, -2) / hd ** 0.5, dim=-1)
    out = (a @ v).transpose(1, 2).reshape(B, L, C_)
    dout = _q(_rand((B, L, C_), 81))
    (ref,) = torch.autograd.grad(out, qkv, dout)
    qb, ob, db = qkv.detach().to(torch.bfloat16).contiguous(), out.detach().to(torch.bfloat16).contiguous(), dout.to(torch.bfloat16).contiguous()
    dqkv = torch.full_like(qb, float("nan"))
    d = _lib.AttnBwdDesc()
    d.qkv, d.out, d.dout, d.dqkv, d.B, d.L, d.heads, d.C = qb.data_ptr(), ob.data_ptr(), db.data_ptr(), dqkv.data_ptr(), B, L, heads, C_
    _lib.check(lib.dmc_attention_backward(C.byref(d), _lib.stream_ptr()), "attention_backward")
    torch.cuda.synchronize()
    assert torch.isfinite(dqkv).all()
    # o is read in bf16 for D = do . o, P and dS are rounded to bf16 for the second GEMMs, outputs stored in bf16
    assert rel_l2(dqkv.float(), ref) < 1.2e-2
    for i in range(3):  # q, k and v parts separately
        assert rel_l2(dqkv[..., i * C_:(i + 1) * C_].float(), ref[..., i * C_:(i + 1) * C_]) < 1.5e-2, i


def test_small_training_kernels():
    lib = _lib.load()
    st = _lib.stream_ptr()
    B, H, C_ = 3, 8, 128
    t = _rand((B, H * H, C_), 1).to(torch.bfloat16)
    out = torch.zeros(C_, device="cuda")
    scratch = torch.empty(B * C_, device="cuda")
    _lib.check(lib.dmc_channel_sum(t.data_ptr(), out.data_ptr(), B, H * H, C_, 0, 0, scratch.data_ptr(), st), "channel_sum")
    per = torch.zeros(B, C_, device="cuda")
    _lib.check(lib.dmc_channel_sum(t.data_ptr(), per.data_ptr(), B, H * H, C_, 1, 0, None, st), "channel_sum per image")
    big = _rand((5, 32 * 32, 768), 7).to(torch.bfloat16)  # several channel chunks, unrolled pixel loop + tail
    big_out, big_scr = torch.full((768,), 3.0, device="cuda"), torch.empty(5 * 768, device="cuda")
    _lib.check(lib.dmc_channel_sum(big.data_ptr(), big_out.data_ptr(), 5, 1024, 768, 0, 1, big_scr.data_ptr(), st), "channel_sum acc")
    dil = torch.full((B, 2 * H, 2 * H, C_), float("nan"), device="cuda", dtype=torch.bfloat16)
    _lib.check(lib.dmc_dilate2x(t.data_ptr(), dil.data_ptr(), B, H, H, C_, st), "dilate2x")
    hi = _rand((B, 2 * H, 2 * H, C_), 2).to(torch.bfloat16)
    lo = torch.zeros((B, H, H, C_), device="cuda", dtype=torch.bfloat16)
    _lib.check(lib.dmc_block_sum2x2(hi.data_ptr(), lo.data_ptr(), B, H, H, C_, 0, st), "block_sum")
    src = _rand((B, 3, H, H), 3)
    dst = torch.full((B, H, H, 64), float("nan"), device="cuda", dtype=torch.bfloat16)
    _lib.check(lib.dmc_nchw_f32_to_nhwc_bf16(src.data_ptr(), dst.data_ptr(), B, 3, H * H, 64, st), "pack")
    torch.cuda.synchronize()
    assert rel_l2(out, t.float().sum(dim=(0, 1))) < 1e-5
    assert rel_l2(per, t.float().sum(dim=1)) < 1e-5
    assert rel_l2(big_out, 3.0 + big.float().sum(dim=(0, 1))) < 1e-5
    want_d = torch.zeros_like(dil)
    want_d[:, ::2, ::2] = t.reshape(B, H, H, C_)
    assert torch.equal(dil, want_d)
    want = hi.float().reshape(B, H, 2, H, 2, C_).sum(dim=(2, 4))
    assert rel_l2(lo.float(), want) < 5e-3
    assert torch.equal(dst[..., :3].float(), src.permute(0, 2, 3, 1).to(torch.bfloat16).float()) and float(dst[..., 3:].abs().max()) == 0.0


@pytest.mark.parametrize("cin,cout,H,B", [(128, 128, 32, 2), (256, 256, 16, 1)])
def test_strided_conv_input_gradient(cin, cout, H, B):
    lib = _lib.load()
    x = _q(_rand((B, cin, H, H), 1)).requires_grad_(True)
    w = _rand((cout, cin, 3, 3), 2, (cin * 9) ** -0.5)
    y = F.conv2d(x, w, None, stride=2, padding=1)
    dy = _q(_rand(tuple(y.shape), 3))
    (ref,) = torch.autograd.grad(y, x, dy)
    dx = torch.full((B, H, H, cin), float("nan"), device="cuda", dtype=torch.bfloat16)
    dyn = nhwc_bf16(dy)
    _lib.check(lib.dmc_conv_dgrad_strided(dyn.data_ptr(), w.data_ptr(), dx.data_ptr(), B, H, H, cin, cout, 2, 0, _lib.stream_ptr()), "dgrad")
    torch.cuda.synchronize()
    assert rel_l2(dx.float().permute(0, 3, 1, 2), ref) < 4e-3


@pytest.mark.parametrize("cin,cout,H,B", [(128, 128, 32, 2), (256, 256, 16, 3), (256, 256, 8, 2)])
def test_strided_conv_input_gradient_as_dilate_plus_conv(cin, cout, H, B):
    """Downsample backward on tensor cores: dY spread onto the input grid, then the stride-1 flipped/transposed-weight conv"""
    lib = _lib.load()
    x = _q(_rand((B, cin, H, H), 1)).requires_grad_(True)
    w = _q(_rand((cout, cin, 3, 3), 2, (cin * 9) ** -0.5))
    y = F.conv2d(x, w, None, stride=2, padding=1)
    dy = _q(_rand(tuple(y.shape), 3))
    (ref,) = torch.autograd.grad(y, x, dy)
    dyn = nhwc_bf16(dy)
    dil = torch.empty((B, H, H, cout), device="cuda", dtype=torch.bfloat16)
    _lib.check(lib.dmc_dilate2x(dyn.data_ptr(), dil.data_ptr(), B, H // 2, H // 2, cout, _lib.stream_ptr()), "dilate2x")
    wt = w.flip(2, 3).permute(1, 0, 2, 3).contiguous()
    out, _ = run_conv([dil], [9], pack3(wt), cin)
    assert rel_l2(out.float().permute(0, 3, 1, 2), ref) < 4e-3


def test_conv_accumulates_in_place_through_its_residual():
    """the training engine accumulates input gradients with out == residual (each element read, then written, by one thread)"""
    B, cin, cout, H = 2, 256, 128, 16
    x = nhwc_bf16(_q(_rand((B, cin, H, H), 1)))
    w = _q(_rand((cout, cin, 3, 3), 2, (cin * 9) ** -0.5))
    base = nhwc_bf16(_q(_rand((B, cout, H, H), 3)))
    sep, _ = run_conv([x], [9], pack3(w), cout, residual=base)
    inplace = base.clone()
    run_conv([x], [9], pack3(w), cout, residual=inplace, out_tensor=inplace)
    assert torch.equal(sep, inplace)
    w1 = _q(_rand((cout, cin, 1, 1), 4, cin ** -0.5)).reshape(cout, cin)  # short-K variant (resident weights, TMA epilogue)
    sep1, _ = run_conv([x], [1], w1, cout, residual=base)
    inplace1 = base.clone()
    run_conv([x], [1], w1, cout, residual=inplace1, out_tensor=inplace1)
    assert torch.equal(sep1, inplace1)


def test_weight_pack_kernel_matches_torch_bit_for_bit():
    """the one-launch re-pack of the GEMM operands after an optimizer step: forward layout [co][tap][ci] (also as a column range
    of a K-concatenated matrix) and input-gradient layout [ci][flipped tap][co padded]"""
    lib = _lib.load()
    w3 = _rand((96, 160, 3, 3), 1)       # not multiples of 32 tiles on purpose
    w1 = _rand((96, 200, 1, 1), 2)
    head = _rand((3, 128, 3, 3), 3)
    K = 9 * 160 + 200
    fwd = torch.full((96, K), float("nan"), device="cuda", dtype=torch.bfloat16)      # fused conv2 + shortcut operand
    dg3 = torch.full((160, 9 * 96), float("nan"), device="cuda", dtype=torch.bfloat16)
    dg1 = torch.full((64, 96), float("nan"), device="cuda", dtype=torch.bfloat16)      # input channels 100 .. 163 of the 1x1
    dgh = torch.zeros((128, 9 * 128), device="cuda", dtype=torch.bfloat16)             # head: co padded 3 -> 128
    items = []

    def item(src, dst, ci0, cin, taps, mode, ld, col0, cpad):
        it = _lib.PackItem()
        it.src, it.dst = src.data_ptr(), dst.data_ptr()
        it.cout, it.cin_total, it.ci0, it.cin, it.taps, it.mode, it.ld, it.col0, it.cpad = (src.shape[0], src.shape[1], ci0, cin, taps,
                                                                                         mode, ld, col0, cpad)
        items.append(it)

    item(w3, fwd, 0, 160, 9, 0, K, 0, 0)
    item(w1, fwd, 0, 200, 1, 0, K, 9 * 160, 0)
    item(w3, dg3, 0, 160, 9, 1, 9 * 96, 0, 96)
    item(w1, dg1, 100, 64, 1, 1, 96, 0, 96)
    item(head, dgh, 0, 128, 9, 1, 9 * 128, 0, 128)
    table, n = _lib.pack_table(items, torch.device("cuda"))
    _lib.check(lib.dmc_pack_weights(table.data_ptr(), n, _lib.stream_ptr()), "pack")
    torch.cuda.synchronize()
    want_fwd = torch.cat([pack3(w3), w1.reshape(96, 200)], dim=1).to(torch.bfloat16)
    assert torch.equal(fwd, want_fwd)
    assert torch.equal(dg3, pack3(w3.flip(2, 3).permute(1, 0, 2, 3)).to(torch.bfloat16))
    assert torch.equal(dg1, w1.reshape(96, 200)[:, 100:164].t().to(torch.bfloat16))
    want_h = torch.zeros((128, 9, 128), device="cuda")
    want_h[:, :, :3] = head.flip(2, 3).permute(1, 2, 3, 0).reshape(128, 9, 3)
    assert torch.equal(dgh, want_h.reshape(128, -1).to(torch.bfloat16))


def test_weight_gradient_written_into_a_window_of_the_parameter():
    """zero-padded operands (head: 3 real output channels of 128) and one source of a fused 1x1 shortcut over concatenated inputs
    write straight into the parameter's gradient instead of a staging tensor"""
    lib = _lib.load()
    B, H = 2, 16
    x = _q(_rand((B, 128, H, H), 1))
    dy = torch.zeros((B, 128, H, H), device="cuda")
    dy[:, :3] = _q(_rand((B, 3, H, H), 2))
    w = torch.zeros((3, 128, 3, 3), device="cuda", requires_grad=True)
    (ref,) = torch.autograd.grad(F.conv2d(x, w, None, padding=1), w, dy[:, :3])
    xs, dys = nhwc_bf16(x), nhwc_bf16(dy)
    d = _lib.WgradDesc()
    d.x, d.dy, d.B, d.Hin, d.Win, d.Cin, d.Cout, d.stride, d.taps = xs.data_ptr(), dys.data_ptr(), B, H, H, 128, 128, 1, 9
    d.splits = _lib.check(lib.dmc_conv_wgrad_splits(C.byref(d)), "splits")
    partial = torch.empty((d.splits, 128, 9, 128), device="cuda")
    dw = torch.full((3, 128, 3, 3), float("nan"), device="cuda")
    guard = torch.full((64,), 7.0, device="cuda")
    d.partial, d.dw, d.accumulate = partial.data_ptr(), dw.data_ptr(), 0
    d.dw_cout, d.dw_cin, d.dw_cin_total, d.dw_ci0 = 3, 128, 128, 0
    _lib.check(lib.dmc_conv_wgrad(C.byref(d), _lib.stream_ptr()), "wgrad window")
    torch.cuda.synchronize()
    assert rel_l2(dw, ref) < 1e-5 and float(guard.min()) == 7.0
    # 1x1 over the second of two concatenated sources: columns 256 .. 384 of a [128, 384, 1, 1] parameter
    x2 = _q(_rand((B, 128, H, H), 3))
    dy2 = _q(_rand((B, 128, H, H), 4))
    w2 = torch.zeros((128, 128, 1, 1), device="cuda", requires_grad=True)
    (ref2,) = torch.autograd.grad(F.conv2d(x2, w2), w2, dy2)
    full = torch.full((128, 384, 1, 1), float("nan"), device="cuda")
    x2s, dy2s = nhwc_bf16(x2), nhwc_bf16(dy2)
    d2 = _lib.WgradDesc()
    d2.x, d2.dy, d2.B, d2.Hin, d2.Win, d2.Cin, d2.Cout, d2.stride, d2.taps = x2s.data_ptr(), dy2s.data_ptr(), B, H, H, 128, 128, 1, 1
    d2.splits = _lib.check(lib.dmc_conv_wgrad_splits(C.byref(d2)), "splits")
    partial2 = torch.empty((d2.splits, 128, 1, 128), device="cuda")
    d2.partial, d2.dw, d2.accumulate = partial2.data_ptr(), full.data_ptr(), 0
    d2.dw_cout, d2.dw_cin, d2.dw_cin_total, d2.dw_ci0 = 128, 128, 384, 256
    _lib.check(lib.dmc_conv_wgrad(C.byref(d2), _lib.stream_ptr()), "wgrad window 2")
    torch.cuda.synchronize()
    assert rel_l2(full[:, 256:], ref2) < 1e-5 and torch.isnan(full[:, :256]).all()
