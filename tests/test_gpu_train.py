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
