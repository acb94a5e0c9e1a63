#!/usr/bin/env python
"""Profiling driver for ncu (tools only; not part of the product or of the bench):

    python tools/profile_forward.py --batch 1024 --mode forward          # one whole CFG forward inside the profiler range
    python tools/profile_forward.py --batch 1024 --mode ops --ops a,b,c  # the named plan ops, one launch each

Everything before `cudaProfilerStart` (weights, plan building, warm-up forwards that leave realistic activations in the
workspace) is outside the range: run ncu with `--profile-from-start off`.
Prints the plan's op table (index, name, kind, algorithmic FLOPs / bytes) so ncu launches can be matched to layers."""

import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=1024, help="images (the CFG forward runs 2x this many)")
    ap.add_argument("--mode", default="forward", choices=["forward", "ops"])
    ap.add_argument("--ops", default="")
    ap.add_argument("--model", default="unet", choices=["unet", "dit"])
    ap.add_argument("--table-out", default=None)
    args = ap.parse_args()

    from diffusion_models_collection_b200 import _lib, synth
    from diffusion_models_collection_b200.models import DiT, UNet

    dev = torch.device("cuda:0")
    torch.cuda.set_device(dev)
    dit = args.model == "dit"
    if dit:  # BASELINE configs[3]: unconditional DiT-32, one forward of `batch` images
        net = DiT(**synth.CIFAR_DIT, num_classes=None)
        net.load_state_dict(synth.make_dit_state_dict(synth.CIFAR_DIT, None, seed=42))
    else:
        net = UNet(**synth.CIFAR_UNET, num_classes=10)
        net.load_state_dict(synth.make_unet_state_dict(None, 10, seed=42))
    net = net.to(dev).eval()
    B = args.batch
    g = torch.Generator().manual_seed(0)
    x = torch.randn(B, 3, 32, 32, generator=g).to(dev)
    y = (torch.randint(0, 10, (B,), generator=g) + 1).to(dev)
    t = torch.full((B,), 500, device=dev)
    fwd = (lambda: net(x, t, None)) if dit else (lambda: net.forward_cfg(x, t, y))
    with torch.no_grad(), net.uniform_timesteps():
        for _ in range(3):
            fwd()
        plan = net.plan_info(B, cfg=not dit, device=dev)
        lib = _lib.load()
        n = lib.dmc_plan_num_ops(plan.handle)
        table = [dict(index=i, name=plan.op_names[i], kind=_lib.OP_KINDS[lib.dmc_plan_op_kind(plan.handle, i)],
                      flops=lib.dmc_plan_op_flops(plan.handle, i), bytes=lib.dmc_plan_op_bytes(plan.handle, i))
                 for i in range(n)]
        if args.table_out:
            json.dump(dict(images=B if dit else 2 * B, ops=table), open(args.table_out, "w"), indent=1)
        torch.cuda.synchronize()
        torch.cuda.profiler.start()
        if args.mode == "forward":
            fwd()
        else:
            want = [s for s in args.ops.split(",") if s]
            for name in want:
                idx = [o["index"] for o in table if o["name"] == name]
                if not idx:
                    raise SystemExit(f"no op named {name}")
                _lib.check(lib.dmc_plan_run_op(plan.handle, idx[0], _lib.stream_ptr()), name)
        torch.cuda.synchronize()
        torch.cuda.profiler.stop()
    print("profiled", args.mode, "ops:", n)


if __name__ == "__main__":
    main()
