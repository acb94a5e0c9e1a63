#!/usr/bin/env python
"""Device time of ONE attention launch (B images, L = 256 tokens, heads x 64) under each DMC_ATTN_DEBUG timing switch of the
ping-pong kernel (pp:<bits>; profiling aid: results are wrong when a switch is set), for the kernel that keeps P in tensor memory
(ts:0, the default) and for the one-warpgroup-per-tile kernel (old:0)."""
import argparse, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=2048)
    ap.add_argument("--heads", type=int, default=4)
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--modes", default="ts:0,pp:0,old:0,ts:0,pp:0,old:0")
    args = ap.parse_args()
    from gpu_util import Plan
    from diffusion_models_collection_b200 import _lib

    B, L, H = args.batch, 256, args.heads
    C_ = H * 64
    qkv = torch.randn(B, L, 3 * C_, device="cuda").to(torch.bfloat16)
    out = torch.empty(B, L, C_, device="cuda", dtype=torch.bfloat16)
    for mode in args.modes.split(","):
        kind, dbg = mode.split(":")
        os.environ["DMC_ATTN_PP"] = {"ts": "2", "pp": "1", "old": "0"}[kind]
        os.environ["DMC_ATTN_DEBUG"] = dbg
        d = _lib.AttnDesc()
        d.qkv, d.out, d.B, d.L, d.heads, d.C, d.impl = qkv.data_ptr(), out.data_ptr(), B, L, H, C_, 0
        p = Plan()
        p.add("attention", d)
        for _ in range(3):
            p.run()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.iters):
            _lib.check(p.lib.dmc_plan_run(p.h, _lib.stream_ptr()), "run")
        e1.record()
        e1.synchronize()
        ms = e0.elapsed_time(e1) / args.iters
        torch.cuda.synchronize()
        print(f"{mode:10s} {ms * 1e3:8.1f} us   ({B * H} items, {ms * 1e-3 * 1.9e9 * 148 / (B * H):7.0f} clk / item at 1.9 GHz)", flush=True)


if __name__ == "__main__":
    main()
