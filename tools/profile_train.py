"""Runs selected backward ops of one training step inside a cudaProfiler range (for ncu --profile-from-start off).
   python tools/profile_train.py kind[:name[:max]] ...   e.g.  conv_wgrad:down_blocks.4.0.conv2 gn_backward:up_blocks.9.0.conv1.0
   attention_backward::1   (an empty name matches every op of the kind; max caps how many of them run -- ncu --set full replays
   each kernel ~40 times, keep the selection to a dozen kernels)"""
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, ".")
from diffusion_models_collection_b200 import _lib, synth  # noqa: E402
from diffusion_models_collection_b200.models import UNet  # noqa: E402


def main():
    sel = [(a.split(":") + ["", ""])[:3] for a in sys.argv[1:]]
    left = {i: (int(k[2]) if k[2] else 10 ** 9) for i, k in enumerate(sel)}
    B = 128
    torch.manual_seed(0)
    net = UNet(**synth.CIFAR_UNET, num_classes=10)
    net.load_state_dict(synth.make_unet_state_dict(None, 10, seed=42))
    net = net.cuda().train()
    x = torch.randn(B, 3, 32, 32, device="cuda")
    t = torch.randint(0, 1000, (B,), device="cuda")
    y = torch.randint(0, 11, (B,), device="cuda")
    for _ in range(2):
        F.mse_loss(torch.randn_like(x), net(x, t, y)).backward()
    torch.cuda.synchronize()
    eng = next(iter(net._train_engines.values()))
    st = _lib.stream_ptr()
    picked = []
    for s in reversed(eng.segs):
        for fn, args, m in eng.bwd[s]:
            if args is None:
                continue
            for i, k in enumerate(sel):
                if m["kind"] == k[0] and (not k[1] or m["name"] == k[1]) and left[i] > 0:
                    left[i] -= 1
                    picked.append((fn, args, m))
                    break
    print("profiling", [(m["kind"], m["name"]) for _, _, m in picked])
    torch.cuda.profiler.start()
    for fn, args, m in picked:
        fn(*args, st)
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()


if __name__ == "__main__":
    main()
