#!/usr/bin/env python
"""Per-op device timings of one UNet CFG forward under the CURRENT environment (DMC_CONV_* switches) -- an A/B aid:
    DMC_CONV_SLAB=0 python tools/bench_ops.py --batch 1024 --out gpurun_out/ops_noslab.json"""
import argparse, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=1024)
    ap.add_argument("--out", default=None)
    ap.add_argument("--iters", type=int, default=5)
    args = ap.parse_args()
    from diffusion_models_collection_b200 import synth
    from diffusion_models_collection_b200.models import UNet
    dev = torch.device("cuda:0")
    net = UNet(**synth.CIFAR_UNET, num_classes=10)
    net.load_state_dict(synth.make_unet_state_dict(None, 10, seed=42))
    net = net.to(dev).eval()
    B = args.batch
    x = torch.randn(B, 3, 32, 32, device=dev)
    y = torch.randint(1, 11, (B,), device=dev)
    t = torch.full((B,), 500, device=dev)
    with torch.no_grad(), net.uniform_timesteps():
        net.forward_cfg(x, t, y)
        plan = net.plan_info(B, cfg=True, device=dev)
        ops = plan.time_ops(iters=args.iters)
    fam = {}
    for o in ops:
        fam[o["kind"]] = fam.get(o["kind"], 0.0) + o["ms"]
    tag = {k: v for k, v in os.environ.items() if k.startswith("DMC_")}
    print(tag, {k: round(v, 3) for k, v in fam.items()}, "total", round(sum(fam.values()), 3))
    if args.out:
        json.dump({"images": 2 * B, "env": tag, "ops": ops}, open(args.out, "w"), indent=1)


if __name__ == "__main__":
    main()
