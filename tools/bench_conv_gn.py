#!/usr/bin/env python
"""Micro-benchmark of ONE convolution with / without the fused GroupNorm epilogue (CUDA events around repeated plan runs).
    python tools/bench_conv_gn.py            # the layer classes of the CIFAR UNet at 2048 images
DMC_GN_DEBUG bits (timing experiments only, wrong results): 1 no cross-CTA wait, 2 no pass 2, 4 no table, 8 no park in TMEM."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ctypes as C  # noqa: E402

import torch  # noqa: E402

from diffusion_models_collection_b200 import _lib  # noqa: E402
from tests.gpu_util import Plan  # noqa: E402


def bench(cin, cout, H, taps, B, nver, raw, residual, stride=1, iters=20):
    dev = "cuda"
    src = torch.randn(B, H, H, cin, device=dev).to(torch.bfloat16)
    K = taps * cin
    w = (torch.randn(cout, K, device=dev) * K ** -0.5).to(torch.bfloat16)
    bias = torch.randn(cout, device=dev) * 0.1
    d = _lib.ConvDesc()
    d.nsrc = 1
    d.src[0], d.src_c[0], d.src_taps[0] = src.data_ptr(), cin, taps
    d.B, d.Hin, d.Win, d.stride, d.up_phase = B, H, H, stride, -1
    d.weight, d.Cout, d.Cout_pad, d.Ktot = w.data_ptr(), cout, cout, K
    d.bias = bias.data_ptr()
    cond = torch.randn(B, cout, device=dev) * 0.1
    if os.environ.get("BENCH_COND", "1") == "1" and taps == 9:
        d.cond, d.cond_stride = cond.data_ptr(), cout
    Ho = H // stride
    keep = [src, w, bias, cond]
    if residual:
        r = torch.randn(B, Ho, Ho, cout, device=dev).to(torch.bfloat16)
        d.residual = r.data_ptr()
        keep.append(r)
    if raw:
        out = torch.empty(B, Ho, Ho, cout, device=dev, dtype=torch.bfloat16)
        d.out_bf16 = out.data_ptr()
        keep.append(out)
    slots = max(1, Ho * Ho // 32)
    st = torch.empty(B, slots, cout // 8, 2, device=dev)
    d.stats, d.stats_slots = st.data_ptr(), slots
    keep.append(st)
    gsz = cout // 8
    d.gn_nver, d.gn_eps = nver, 1e-5
    for i in range(nver):
        pitch = cout if i == 0 else 2 * cout
        t = torch.empty(B, Ho, Ho, pitch, device=dev, dtype=torch.bfloat16)
        g, b_ = torch.ones(cout, device=dev), torch.zeros(cout, device=dev)
        keep += [t, g, b_]
        d.gn_out[i], d.gn_pitch[i], d.gn_coff[i] = t.data_ptr(), pitch, 0
        d.gn_gamma[i], d.gn_beta[i] = g.data_ptr(), b_.data_ptr()
        d.gn_gsize[i], d.gn_silu[i] = gsz if i == 0 else 2 * gsz, 1
    cnt = torch.zeros(2 * B * max(1, cout // 32), dtype=torch.int32, device=dev)
    d.gn_counters = cnt.data_ptr()
    keep.append(cnt)
    p = Plan()
    p.add("conv", d)
    for _ in range(3):
        p.run()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        _lib.check(p.lib.dmc_plan_run(p.h, _lib.stream_ptr()), "run")
    e1.record()
    torch.cuda.synchronize()
    cnt.zero_()
    return e0.elapsed_time(e1) / iters


if __name__ == "__main__":
    B = int(os.environ.get("B", "2048"))
    layers = [("32x32 128->128 3x3 (K=18)", 128, 128, 32, 9), ("16x16 256->256 3x3 (K=36)", 256, 256, 16, 9),
              ("16x16 512->256 3x3 (K=72)", 512, 256, 16, 9), ("8x8 256->256 3x3", 256, 256, 8, 9), ("16x16 256->256 1x1", 256, 256, 16, 1)]
    for name, cin, cout, H, taps in layers:
        row = [f"{name:28s}"]
        for nver, raw, res in ((0, True, False), (1, False, False), (1, True, True), (2, True, True)):
            row.append(f"v{nver}{'R' if raw else '-'}{'r' if res else '-'} {bench(cin, cout, H, taps, B, nver, raw, res):.3f}")
        print("  ".join(row), flush=True)
