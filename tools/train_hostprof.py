"""Host-side profile (cProfile) of the training inner loop on the native UNet: where do the Python / launch microseconds go?"""
import cProfile
import pstats
import sys
import time

import torch

sys.path.insert(0, ".")
from diffusion_models_collection_b200 import synth  # noqa: E402
from diffusion_models_collection_b200.diffusion import DDPM  # noqa: E402
from diffusion_models_collection_b200.models import UNet  # noqa: E402


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 128
    dev = torch.device("cuda")
    net = UNet(**synth.CIFAR_UNET, num_classes=10)
    net.load_state_dict(synth.make_unet_state_dict(None, 10, seed=42))
    net = net.to(dev).train()
    ddpm = DDPM(1000, 1e-4, 0.02, "linear", device=dev)
    opt = torch.optim.AdamW(net.parameters(), lr=2e-4, weight_decay=1e-4, fused=True)
    x = torch.rand(B, 3, 32, 32, device=dev) * 2 - 1
    y = torch.randint(1, 11, (B,), device=dev)

    def step():
        t = torch.randint(0, 1000, (B,), device=dev).long()
        loss = ddpm.p_losses(net, x, t, y, loss_type="l2")
        loss.backward()
        torch.nn.utils.clip_grad_norm_(net.parameters(), 1.0)
        opt.step()
        opt.zero_grad()

    for _ in range(5):
        step()
    torch.cuda.synchronize()
    # coarse split first (wall clock of host enqueue per phase, no sync)
    acc = {"fwd": 0.0, "bwd": 0.0, "clip": 0.0, "opt": 0.0}
    n = 30
    for _ in range(n):
        t0 = time.perf_counter()
        t = torch.randint(0, 1000, (B,), device=dev).long()
        loss = ddpm.p_losses(net, x, t, y, loss_type="l2")
        t1 = time.perf_counter()
        loss.backward()
        t2 = time.perf_counter()
        torch.nn.utils.clip_grad_norm_(net.parameters(), 1.0)
        t3 = time.perf_counter()
        opt.step()
        opt.zero_grad()
        t4 = time.perf_counter()
        acc["fwd"] += t1 - t0; acc["bwd"] += t2 - t1; acc["clip"] += t3 - t2; acc["opt"] += t4 - t3
    torch.cuda.synchronize()
    print({k: round(v / n * 1e3, 3) for k, v in acc.items()}, "ms per step (host enqueue)")
    pr = cProfile.Profile()
    pr.enable()
    for _ in range(30):
        step()
    pr.disable()
    torch.cuda.synchronize()
    st = pstats.Stats(pr)
    st.sort_stats("tottime").print_stats(28)
    st.sort_stats("cumtime").print_stats(40)


if __name__ == "__main__":
    main()
