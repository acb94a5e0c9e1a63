#!/usr/bin/env python
"""Experiment: 4096 images per launch instead of 2048 (UNet.max_images_per_launch).  Checks that the eps of a 2048-image CFG batch
(4096 images) is bit-identical to the chunked run and times one forward per image both ways."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402


def main():
    from diffusion_models_collection_b200 import synth
    from diffusion_models_collection_b200.models import UNet
    dev = torch.device("cuda:0")
    net = UNet(**synth.CIFAR_UNET, num_classes=10)
    net.load_state_dict(synth.make_unet_state_dict(None, 10, seed=42))
    net = net.to(dev).eval()
    B = 2048
    g = torch.Generator().manual_seed(0)
    x = torch.randn(B, 3, 32, 32, generator=g).to(dev)
    y = (torch.randint(0, 10, (B,), generator=g) + 1).to(dev)
    t = torch.full((B,), 500, device=dev)
    outs = {}
    for cap in (2048, 4096):
        net.max_images_per_launch = cap
        with torch.no_grad(), net.uniform_timesteps():
            for _ in range(2):
                a, b = net.forward_cfg(x, t, y)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(5):
                a, b = net.forward_cfg(x, t, y)
            e1.record()
            e1.synchronize()
        ms = e0.elapsed_time(e1) / 5
        outs[cap] = (a.clone(), b.clone())
        print(f"max_images_per_launch {cap}: {ms:.2f} ms per 4096-image CFG forward = {ms / 4096 * 1e3:.3f} us / image", flush=True)
    same = torch.equal(outs[2048][0], outs[4096][0]) and torch.equal(outs[2048][1], outs[4096][1])
    print("bit-identical:", same, "finite:", bool(torch.isfinite(outs[4096][0]).all()))


if __name__ == "__main__":
    main()
