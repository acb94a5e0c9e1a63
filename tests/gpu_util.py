"""Helpers for the GPU parity tests: one-op plans driven through the C ABI (ctypes), bf16 NHWC conversions."""

from __future__ import annotations

import ctypes as C

import torch

from diffusion_models_collection_b200 import _lib


def rel_l2(a, b):
    a = a.detach().double().cpu()
    b = b.detach().double().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def nhwc_bf16(x):
    """fp32 NCHW -> bf16 NHWC contiguous"""
    return x.permute(0, 2, 3, 1).contiguous().to(torch.bfloat16)


def nchw_f32(x_nhwc):
    return x_nhwc.float().permute(0, 3, 1, 2).contiguous()


class Plan:
    def __init__(self):
        self.lib = _lib.load()
        self.h = C.c_void_p()
        _lib.check(self.lib.dmc_plan_create(C.byref(self.h)), "create")
        self.keep = []

    def add(self, name, desc):
        return _lib.check(getattr(self.lib, "dmc_plan_add_" + name)(self.h, C.byref(desc)), name)

    def run(self):
        _lib.check(self.lib.dmc_plan_run(self.h, _lib.stream_ptr()), "run")
        torch.cuda.synchronize()

    def __del__(self):
        try:
            self.lib.dmc_plan_destroy(self.h)
        except Exception:
            pass


def pack3(w):
    return w.permute(0, 2, 3, 1).reshape(w.shape[0], -1)


def run_conv(srcs, taps, wmat, Cout, *, stride=1, bias=None, cond=None, residual=None, out_nchw=False, stats=False,
             impl=0, up_phase=-1, out_tensor=None):
    """srcs: list of bf16 NHWC tensors; wmat: fp32 [Cout_pad, K]; returns (out, stats or None)."""
    dev = srcs[0].device
    B, H, W, _ = srcs[0].shape
    wq = wmat.to(torch.bfloat16).contiguous()
    d = _lib.ConvDesc()
    d.nsrc = len(srcs)
    for i, s in enumerate(srcs):
        d.src[i], d.src_c[i], d.src_taps[i] = s.data_ptr(), s.shape[3], taps[i]
    d.B, d.Hin, d.Win, d.stride, d.up_phase = B, H, W, stride, up_phase
    d.weight, d.Cout, d.Cout_pad, d.Ktot = wq.data_ptr(), Cout, wq.shape[0], wq.shape[1]
    d.bias = bias.data_ptr() if bias is not None else None
    if cond is not None:
        d.cond, d.cond_stride = cond.data_ptr(), cond.shape[1]
    Ho, Wo = (2 * H, 2 * W) if up_phase >= 0 else (H // stride, W // stride)
    if residual is not None:
        d.residual = residual.data_ptr()
    st = None
    if out_nchw:
        out = torch.full((B, Cout, Ho, Wo), float("nan"), device=dev)
        d.out_f32_nchw = out.data_ptr()
    else:
        out = out_tensor if out_tensor is not None else torch.full((B, Ho, Wo, Cout), float("nan"), device=dev,
                                                                   dtype=torch.bfloat16)
        d.out_bf16 = out.data_ptr()
    if stats:
        ppi = (H // stride) * (W // stride)
        slots = max(1, ppi // 32) * (4 if up_phase >= 0 else 1)
        st = torch.full((B, slots, Cout // 8, 2), float("nan"), device=dev)  # every slot must be written
        d.stats, d.stats_slots = st.data_ptr(), slots
    d.impl = impl
    p = Plan()
    p.add("conv", d)
    p.run()
    return out, st
