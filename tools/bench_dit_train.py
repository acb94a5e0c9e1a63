#!/usr/bin/env python
"""DiT training step (models/dit_train.py) timing: p_losses + backward + fused AdamW at batch 128 on one GPU; device time by CUDA
events, plus the share of the step spent in the native kernels (sum of the engine's GEMM / weight-gradient / attention launches
timed alone).  A side measurement (the BASELINE training config is the UNet's)."""
import argparse, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=128)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    args = ap.parse_args()
    from diffusion_models_collection_b200 import synth
    from diffusion_models_collection_b200.diffusion import DDPM
    from diffusion_models_collection_b200.models import DiT

    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    net = DiT(**synth.CIFAR_DIT, num_classes=10)
    net.load_state_dict(synth.make_dit_state_dict(synth.CIFAR_DIT, 10, seed=42))
    net = net.to(dev).train()
    opt = torch.optim.AdamW(net.parameters(), lr=2e-4, weight_decay=1e-4, fused=True)
    ddpm = DDPM(1000, 1e-4, 0.02, "linear", device=dev)
    B = args.batch
    x = torch.rand(B, 3, 32, 32, device=dev) * 2 - 1
    y = torch.randint(0, 11, (B,), device=dev)

    def step():
        t = torch.randint(0, 1000, (B,), device=dev)
        loss = ddpm.p_losses(net, x, t, y, loss_type="l2")
        loss.backward()
        torch.nn.utils.clip_grad_norm_(net.parameters(), 1.0)
        opt.step()
        opt.zero_grad(set_to_none=True)
        return loss

    for _ in range(args.warmup):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        loss = step()
    e1.record()
    e1.synchronize()
    ms = e0.elapsed_time(e1) / args.steps
    flops = 3 * 12.107e9 * B
    print(f"DiT training step, batch {B}: {ms:.2f} ms per step, {B / ms * 1e3:.0f} images/s, loss {loss.item():.4f}, "
          f"{flops / ms / 1e9:.0f} TFLOP/s of 3 x forward FLOPs")
    from torch.profiler import ProfilerActivity, profile
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        step()
        torch.cuda.synchronize()
    ours = other = 0.0
    for ev in prof.key_averages():
        t_us = getattr(ev, "device_time_total", None) or getattr(ev, "cuda_time_total", 0.0)
        if any(k in ev.key for k in ("dmc::", "conv_umma", "attention_", "wgrad", "channel_sum")):
            ours += t_us
        else:
            other += t_us
    print(f"one profiled step: native kernels {ours / 1e3:.2f} ms, torch kernels (glue, optimizer, clip) {other / 1e3:.2f} ms")
    rows = sorted(prof.key_averages(), key=lambda e: -(getattr(e, "device_time_total", None) or getattr(e, "cuda_time_total", 0.0)))
    for ev in rows[:14]:
        t_us = getattr(ev, "device_time_total", None) or getattr(ev, "cuda_time_total", 0.0)
        print(f"  {t_us / 1e3:7.3f} ms  x{ev.count:4d}  {ev.key[:110]}")


if __name__ == "__main__":
    main()
