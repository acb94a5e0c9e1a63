#!/usr/bin/env python
"""bench.py -- DDIM-50 UNet CIFAR-10 sampling throughput (BASELINE.json metric) on N B200s of one node.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--batch B] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W

Workload = BASELINE.json configs[2]: class-conditional UNet (10 classes + null label), DDIM 50 steps, classifier-free
guidance 3.0 with the reference's default dynamic thresholding (p = 0.995), global batch 4096 images sharded over the
N ranks (contiguous slices, no data-path collective; one NCCL all-gather of the final images).  One "step" = one complete
DDIM-50 sampling of the global batch.  `value` is images/s with x_T and the labels already resident in HBM; `e2e` is the
same call fed from pinned HOST buffers (labels + x_T) with the images read back to the host inside the timed region.

`--impl reference` times the REFERENCE ITSELF on the host cores: its own `DDIM.sample_with_cfg(UNet, ...)`, imported from
oracle/_ref (the reference's sources packed by oracle/make_ref.py; /root/reference does not exist on the GPU box), on a
bounded sample of the same workload (`--ref-batch` images per step, all 50 DDIM steps, no extrapolation), plus BASELINE
configs[0] verbatim (`cpu_baseline_config1`: uncond DDIM-50, batch 16, default init under seed 42).
"""

from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

os.environ.setdefault("TQDM_DISABLE", "1")  # the reference's samplers draw tqdm bars on stderr

import torch  # noqa: E402

METRIC = "ddim50_unet_cifar10_images_per_sec"
UNIT = "images/s"
FLOPS_PER_IMAGE_FORWARD = 12.638e9  # cond UNet, SURVEY.md 8(d) (hooks on the reference modules)
FALLBACK_PEAKS = {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        d["_source"] = "measured"
        return d
    d = dict(FALLBACK_PEAKS)
    d["_source"] = "fallback"
    return d


class ClockSampler:
    """samples nvidia-smi clocks / throttle reasons of one GPU while the timed region runs"""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.stop_flag, self.th = index, [], threading.Event(), None

    def _run(self):
        while not self.stop_flag.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([v.strip() for v in out.split(",")])
            except Exception:
                pass
            self.stop_flag.wait(0.2)

    def __enter__(self):
        self.th = threading.Thread(target=self._run, daemon=True)
        self.th.start()
        return self

    def __exit__(self, *a):
        self.stop_flag.set()
        self.th.join(timeout=6)

    def summary(self):
        sm = sorted(float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit())
        reasons = set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            for n, v in zip(names, r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx[0] if mx else None,
                "reasons": sorted(reasons), "samples": len(self.rows)}


# ------------------------------------------------------------------------------------------------------
# CPU baseline: the REFERENCE ITSELF (oracle/_ref: its own sources, packed by oracle/make_ref.py) on the host cores
# ------------------------------------------------------------------------------------------------------
_REF = {}


def _reference():
    """the reference's own classes (oracle/ref_loader.py: /root/reference in the build container, the archive
    oracle/_ref/reference_src.zip on the GPU box).  Only when NEITHER exists (a checkout of this repo that never ran
    `python oracle/make_ref.py`): the oracle port behind the same three call signatures, labelled kind = "port"."""
    if not _REF:
        from oracle import ref_loader

        if ref_loader.available():
            _REF.update(ref_loader.import_reference())
            _REF["cpu_kind"] = "reference"
        else:
            _REF.update(_port_adapters())
    return _REF


def _port_adapters():
    """UNet / DDIM / DDPM look-alikes over the CPU oracle restatement (oracle/model_oracle.py, oracle/sched_oracle.py): the
    few calls the CPU legs make -- construct, load_state_dict, eval / train, sample, sample_with_cfg, p_losses"""
    from diffusion_models_collection_b200 import synth
    from oracle import model_oracle, sched_oracle as so

    class PortUNet(torch.nn.Module):
        def __init__(self, num_classes=None, **cfg):
            super().__init__()
            self.cfg, self.num_classes = cfg, num_classes
            sd = synth.make_unet_state_dict(cfg, num_classes, seed=42)
            self.names = list(sd)
            self.ps = torch.nn.ParameterList([torch.nn.Parameter(v) for v in sd.values()])

        def load_state_dict(self, sd, strict=True):
            with torch.no_grad():
                for n, p in zip(self.names, self.ps):
                    p.copy_(sd[n])

        def forward(self, x, t, y=None):
            sd = dict(zip(self.names, self.ps))
            fn = model_oracle.unet_forward if not torch.is_grad_enabled() else model_oracle.unet_forward.__wrapped__
            return fn(sd, self.cfg, x, t, y, self.num_classes)

    class PortDDIM:
        def __init__(self, T, S, b0, b1, sched, eta=0.0, device="cpu"):
            self.tb, self.ts = so.make_tables(T, b0, b1, sched), so.ddim_timesteps(T, S)

        def sample(self, model, shape):
            return so.ddim_sample(model, self.tb, self.ts, torch.randn(shape))

        def sample_with_cfg(self, model, shape, y, cfg_scale=3.0):
            return so.ddim_sample_cfg(model, self.tb, self.ts, torch.randn(shape), y, cfg_scale=cfg_scale)

    class PortDDPM:
        def __init__(self, T, b0, b1, sched, device="cpu"):
            self.tb = so.make_tables(T, b0, b1, sched)

        def p_losses(self, model, x0, t, y=None, loss_type="l2"):
            noise = torch.randn_like(x0)
            return torch.nn.functional.mse_loss(noise, model(so.q_sample(self.tb, x0, t, noise), t, y))

    return dict(UNet=PortUNet, DDIM=PortDDIM, DDPM=PortDDPM, kind="oracle port: oracle/_ref is missing", cpu_kind="port")


def _set_seed(seed):
    # utils/helpers.py:12-19 of the reference (set_seed), minus the CUDA / cudnn lines that do nothing on the host
    import random

    import numpy as np

    random.seed(seed)
    np.random.seed(seed)
    torch.manual_seed(seed)


def cpu_reference_config1():
    """BASELINE.json configs[0] verbatim, through the reference's own code: set_seed(42); UNet(**cifar10_unet model_params,
    num_classes=None) default init, eval; DDIM(1000, 50, 1e-4, 0.02, 'linear', eta=0).sample(model, (16, 3, 32, 32)) --
    all 50 steps, fp32, every host thread.  No extrapolation."""
    from diffusion_models_collection_b200 import synth

    ref = _reference()
    torch.set_num_threads(os.cpu_count() or 1)
    _set_seed(42)
    net = ref["UNet"](**synth.CIFAR_UNET, num_classes=None).eval()
    d = ref["DDIM"](1000, 50, 1e-4, 0.02, "linear", eta=0.0, device="cpu")
    t0 = time.perf_counter()
    with torch.no_grad():
        img = d.sample(net, (16, 3, 32, 32))
    dt = time.perf_counter() - t0
    assert tuple(img.shape) == (16, 3, 32, 32) and bool(torch.isfinite(img).all())
    return {"value": 16 / dt, "unit": UNIT, "cores": torch.get_num_threads(), "kind": ref["cpu_kind"], "seconds": dt,
            "sample": "BASELINE.json configs[0] verbatim: the reference's own DDIM(50 steps).sample(UNet uncond, default init "
                      f"under seed 42, (16, 3, 32, 32)), fp32 on the host, all 50 steps, no extrapolation ({ref['kind']})"}


_CFG_NET = {}


def cpu_reference_images_per_sec(batch=4):
    """images/s of THIS bench's workload (cond UNet, DDIM-50, CFG 3.0, dynamic threshold 0.995) through the reference's own
    DDIM.sample_with_cfg on `batch` images: all 50 steps, two UNet forwards each, fp32, every host thread -- a bounded
    sample of the 4096-image step (per-image cost does not depend on the batch), not an extrapolation over steps."""
    from diffusion_models_collection_b200 import synth

    ref = _reference()
    torch.set_num_threads(os.cpu_count() or 1)
    if "net" not in _CFG_NET:
        net = ref["UNet"](**synth.CIFAR_UNET, num_classes=10)
        net.load_state_dict(synth.make_unet_state_dict(None, 10, seed=42))  # the native arm's weights
        _CFG_NET["net"] = net.eval()
        _CFG_NET["ddim"] = ref["DDIM"](1000, 50, 1e-4, 0.02, "linear", eta=0.0, device="cpu")
    g = torch.Generator().manual_seed(42)
    y = torch.randint(0, 10, (batch,), generator=g) + 1
    torch.manual_seed(42)
    t0 = time.perf_counter()
    with torch.no_grad():
        img = _CFG_NET["ddim"].sample_with_cfg(_CFG_NET["net"], (batch, 3, 32, 32), y, cfg_scale=3.0)
    dt = time.perf_counter() - t0
    assert tuple(img.shape) == (batch, 3, 32, 32) and bool(torch.isfinite(img).all())
    return batch / dt, {"value": batch / dt, "unit": UNIT, "cores": torch.get_num_threads(), "kind": ref["cpu_kind"],
                        "sample": f"{batch} images of the 4096-image step through the reference's own DDIM.sample_with_cfg (all 50 "
                                  f"steps, 2 UNet forwards each, CFG 3.0, dynamic threshold 0.995), fp32 on the host ({ref['kind']})"}


def cpu_reference_train_images_per_sec(batch=8, steps=1):
    """images/s of the reference's training step on the host through its OWN code: DDPM.p_losses(UNet cond, train mode, l2) ->
    backward -> clip_grad_norm_(1.0) -> AdamW(lr 2e-4, wd 1e-4) (utils/trainer.py:221-262) on `batch` images, every host thread."""
    from diffusion_models_collection_b200 import synth

    ref = _reference()
    torch.set_num_threads(os.cpu_count() or 1)
    torch.manual_seed(42)
    net = ref["UNet"](**synth.CIFAR_UNET, num_classes=10)
    net.load_state_dict(synth.make_unet_state_dict(None, 10, seed=42))
    net.train()
    ddpm = ref["DDPM"](1000, 1e-4, 0.02, "linear", device="cpu")
    opt = torch.optim.AdamW(net.parameters(), lr=2e-4, weight_decay=1e-4)
    g = torch.Generator().manual_seed(42)
    x0 = torch.rand(batch, 3, 32, 32, generator=g) * 2 - 1
    y = torch.randint(0, 10, (batch,), generator=g) + 1
    best = None
    for _ in range(steps):
        t0 = time.perf_counter()
        t = torch.randint(0, 1000, (batch,), generator=g)
        loss = ddpm.p_losses(net, x0, t, y, loss_type="l2")
        loss.backward()
        torch.nn.utils.clip_grad_norm_(net.parameters(), 1.0)
        opt.step()
        opt.zero_grad()
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    return batch / best, {"value": batch / best, "unit": UNIT, "cores": torch.get_num_threads(), "kind": ref["cpu_kind"],
                          "sample": f"{steps} training step(s) of {batch} images through the reference's own DDPM.p_losses + "
                                    f"backward + clip_grad_norm_ + AdamW, fp32 on the host ({ref['kind']})"}


def run_reference_arm(args, rank):
    if rank != 0:
        return
    if args.workload == "train":  # side measurement: the reference's training step on the host cores
        for _ in range(min(args.warmup, 1)):
            cpu_reference_train_images_per_sec(batch=2, steps=1)
        t0 = time.perf_counter()
        v, cb = cpu_reference_train_images_per_sec(batch=args.ref_batch * 2, steps=args.steps)
        wall = time.perf_counter() - t0
        print(json.dumps({"impl": "reference", "metric": "train_unet_cifar10_images_per_sec", "value": v, "unit": UNIT,
                          "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * wall / args.steps,
                          "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                          "config": {"workload": "UNet cond CIFAR-10 training step (BASELINE.json configs[4]), the reference's own "
                                                 "code on the host cores", "per_gpu_batch": args.ref_batch * 2},
                          "cpu_baseline": cb, "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                          "gpu_launches": 0}), flush=True)
        return
    # every step = the reference's own DDIM-50 + CFG sampler on `rb` images of the 4096-image step (all 50 denoising steps);
    # rb = --ref-batch, halved until the K + W steps fit ~200 s of host time (measured on the first run)
    rb = max(1, args.ref_batch)
    t0 = time.perf_counter()
    cpu_reference_images_per_sec(batch=rb)  # warm-up 1 (always: it sizes the sample)
    t_one = time.perf_counter() - t0
    while rb > 1 and (args.steps + max(args.warmup, 1) - 1) * t_one > 200.0:
        rb, t_one = rb // 2, t_one / 2
    for _ in range(max(args.warmup, 1) - 1):
        cpu_reference_images_per_sec(batch=rb)
    cb = None
    t0 = time.perf_counter()
    for _ in range(args.steps):
        _, cb = cpu_reference_images_per_sec(batch=rb)
    wall = time.perf_counter() - t0
    v = rb * args.steps / wall
    cb["value"] = v
    shard = (args.batch + args.gpus - 1) // args.gpus
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * wall / args.steps, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": workload_config(args, shard),
            "cpu_baseline": cb, "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    if not args.no_cpu_baseline:
        line["cpu_baseline_config1"] = cpu_reference_config1()
    print(json.dumps(line), flush=True)


def workload_config(args, per_rank):
    return {"workload": "UNet cond (10+1 null) CIFAR-10 32x32, DDIM-50, CFG 3.0, dynamic threshold 0.995 "
                        "(BASELINE.json configs[2])", "global_batch": args.batch, "per_gpu_batch": per_rank,
            "sampler": "ddim50", "cfg_scale": 3.0, "parallelism": f"sample-sharded x{args.gpus}",
            "l2": "inputs larger than L2 (activations of one forward are >10x the 126 MB L2)",
            "schedule_tables": "built on the device with the reference's own torch expressions: bit-equal to what the reference "
                               "computes on the same device, last-bit differences against its CPU tables (linspace / cumprod / sqrt)"}


# ------------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--batch", type=int, default=None, help="global batch (images per step); default 4096 (256 for ddpm1000)")
    ap.add_argument("--workload", default=None, choices=["ddim50_cfg", "ddpm1000", "eval_ddpm1000_cfg", "dit_ddim50", "dit64_ddim50", "train"],
                    help="ddim50_cfg: BASELINE configs[2], the bench line (default); ddpm1000: configs[1] (uncond UNet, DDPM "
                         "1000 steps, batch 256); dit_ddim50: configs[3] (same as --model dit).  The last two are side "
                         "measurements recorded under profiles/, not the headline metric")
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--model", default="unet", choices=["unet", "dit"],
                    help="unet: BASELINE configs[2] (the bench line); dit: configs[3] (DiT patch-2 DDIM-50 uncond, --batch 1024), "
                         "a side measurement, not the headline")
    ap.add_argument("--ref-batch", type=int, default=8,
                    help="images per step of the CPU reference legs (every step runs all 50 DDIM steps on them)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-roofline", action="store_true")
    ap.add_argument("--ops-out", default=None, help="write the per-op timing table (JSON) here")
    args = ap.parse_args()
    if args.workload is None:
        args.workload = "dit_ddim50" if args.model == "dit" else "ddim50_cfg"
    if args.batch is None:
        args.batch = 256 if args.workload in ("ddpm1000", "dit64_ddim50") else (
            128 if args.workload == "train" else (512 if args.workload == "eval_ddpm1000_cfg" else 4096))

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference_arm(args, rank)
        return
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            # convenience: re-launch ourselves under torchrun
            cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
                   "--master-addr", "127.0.0.1", "--master-port", os.environ.get("MASTER_PORT", "29511"),
                   os.path.abspath(__file__)] + sys.argv[1:]
            sys.exit(subprocess.call(cmd))
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")

    if args.workload == "train":
        run_train_workload(args, rank, local_rank, world)
        return

    import torch.distributed as dist

    from diffusion_models_collection_b200 import synth
    from diffusion_models_collection_b200.diffusion import DDIM, DDPM
    from diffusion_models_collection_b200.models import DiT, UNet
    from diffusion_models_collection_b200.sharding import sharded_sample, sharded_sample_with_cfg, shard_bounds

    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    is_dit = args.workload in ("dit_ddim50", "dit64_ddim50")
    dit64 = args.workload == "dit64_ddim50"
    img = 64 if dit64 else 32
    is_eval = args.workload == "eval_ddpm1000_cfg"  # evaluate.py:201-219: DDPM-1000 + CFG at the published batch 512
    is_ddpm = args.workload == "ddpm1000" or is_eval
    uncond = is_dit or (is_ddpm and not is_eval)
    sampler_steps = 1000 if is_ddpm else 50
    if is_dit:
        dcfg = dict(synth.CIFAR_DIT, img_size=(img, img))
        net = DiT(**dcfg, num_classes=None)
        net.load_state_dict(synth.make_dit_state_dict(dcfg, None, seed=42))
    elif is_ddpm and not is_eval:
        net = UNet(**synth.CIFAR_UNET, num_classes=None)
        net.load_state_dict(synth.make_unet_state_dict(None, None, seed=42))
    else:
        net = UNet(**synth.CIFAR_UNET, num_classes=10)
        net.load_state_dict(synth.make_unet_state_dict(None, 10, seed=42))
    net = net.to(dev).eval()
    if is_ddpm:
        ddim = DDPM(1000, 1e-4, 0.02, "linear", device=dev)  # the sampler object of this workload
    else:
        ddim = DDIM(1000, 50, 1e-4, 0.02, "linear", eta=0.0, device=dev)
    ddim.progress = False

    B = args.batch
    lo, hi = shard_bounds(B, rank, world)
    nb = hi - lo
    g = torch.Generator().manual_seed(42)
    y_host = (torch.randint(0, 10, (B,), generator=g) + 1).pin_memory()
    xT_host = torch.randn(B, 3, img, img, generator=g).pin_memory()
    y_dev, xT_dev = y_host.to(dev), xT_host.to(dev)
    shape = (B, 3, img, img)

    def step_resident():
        if uncond:
            return sharded_sample(ddim, net, shape, None, noise=xT_dev, rank=rank, world=world)
        return sharded_sample_with_cfg(ddim, net, shape, y_dev, cfg_scale=3.0, noise=xT_dev, rank=rank, world=world)

    def step_e2e():
        # host -> device of this rank's slice of the inputs, device -> host of the gathered images
        yl = y_host[lo:hi].to(dev, non_blocking=True)
        xl = xT_host[lo:hi].to(dev, non_blocking=True)
        if uncond:
            out = sharded_sample(ddim, net, shape, None, noise=xl, rank=rank, world=world, sliced=True)
        else:
            out = sharded_sample_with_cfg(ddim, net, shape, yl, cfg_scale=3.0, noise=xl, rank=rank, world=world, sliced=True)
        return out.cpu() if rank == 0 else out[:1].cpu()

    def timed(fn, n):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
            dist.barrier()
        return float(ms.item())

    for _ in range(args.warmup):
        step_resident()
    with ClockSampler(local_rank) as cs:
        ms = timed(step_resident, args.steps)
    clocks = cs.summary()
    value = B * args.steps / (ms / 1e3)

    step_e2e()
    ms_e2e = timed(step_e2e, max(1, min(args.steps, 2)))
    e2e_value = B * max(1, min(args.steps, 2)) / (ms_e2e / 1e3)

    # launches of OUR kernels in the timed region: per DDIM step, per chunk: one plan run + one fused scheduler kernel
    launches = net.launches_per_forward(nb, cfg=not uncond) * sampler_steps * args.steps

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "bf16", "data": "synthetic", "config": workload_config(args, nb), "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(B * (3 * img * img * 4 + 8)),
                    "d2h_bytes_per_step": int(B * 3 * img * img * 4)},
            "gpu_launches": int(launches)}
    pk = peaks()
    # 2 forwards per DDIM step (cond + uncond) for the CFG UNet workload, 1 for the unconditional DiT one
    flops_step = (50 * (62.855e9 if dit64 else 12.107e9) * B) if is_dit else (
        (2 * 1000 * FLOPS_PER_IMAGE_FORWARD * B) if is_eval else ((1000 * 12.632e9 * B) if is_ddpm else (2 * 50 * FLOPS_PER_IMAGE_FORWARD * B)))
    line["model_flops_utilization"] = {"achieved_tflops": flops_step * args.steps / (ms / 1e3) / 1e12 / world,
                                       "peak_tflops": pk["bf16_tflops_sustained"], "peak_source": pk["_source"]}

    if is_dit:
        line["metric"] = "ddim50_dit_cifar10_images_per_sec"
        line["config"]["workload"] = (f"DiT patch-2 (hidden 384, depth 12, 6 heads) {img}x{img} unconditional, DDIM-50 "
                                      "(BASELINE.json configs[3]" + (", the shipped 64x64 image size" if dit64 else "") +
                                      "); side measurement, not the headline metric")
        line["config"]["cfg_scale"] = None
    if is_ddpm:
        line["metric"] = "ddpm1000_unet_cifar10_images_per_sec"
        line["config"]["workload"] = ("UNet unconditional CIFAR-10 32x32, DDPM 1000 steps, fresh N(0,1) noise every step "
                                      "(BASELINE.json configs[1]); side measurement, not the headline metric")
        line["config"]["sampler"] = "ddpm1000"
        line["config"]["cfg_scale"] = None
    if is_eval:
        line["metric"] = "ddpm1000_cfg_unet_cifar10_images_per_sec"
        line["config"]["workload"] = ("UNet cond (10+1 null) CIFAR-10 32x32, DDPM 1000 steps + CFG 3.0 + dynamic threshold, batch 512 per "
                                      "call: the generation loop of the reference's evaluate.py:201-219 at its published batch size "
                                      "(docs/cifar10_runs.md:129); side measurement, not the headline metric")
        line["config"]["cfg_scale"] = 3.0
    if rank == 0 and not args.no_roofline:
        line["roofline"] = roofline_leg(net, dev, nb, pk, args.ops_out, cfg=not uncond, step_ms=ms / args.steps / sampler_steps,
                                        chunks=-(-nb // max(1, net.max_images_per_launch // (1 if uncond else 2))))
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        # the reference itself on the host cores: this workload on a bounded sample, and BASELINE configs[0] verbatim
        _, cb = cpu_reference_images_per_sec(batch=args.ref_batch)
        line["cpu_baseline"] = cb
        if not is_dit and not is_ddpm:
            line["cpu_baseline_config1"] = cpu_reference_config1()
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def run_train_workload(args, rank, local_rank, world):
    """BASELINE.json configs[4] (side measurement): the reference trainer's inner loop (utils/trainer.py:221-262) on the native
    UNet -- labels + 1 with CFG label dropout 0.2, t ~ U[0, 1000), p_losses (q_sample + eps-MSE), backward, clip_grad_norm_(1.0),
    AdamW(lr 2e-4, wd 1e-4), EMA(0.9999) on rank 0 -- per-GPU batch `--batch` (default 128, weak scaling), DistributedDataParallel
    over NCCL for N > 1 (gradient all-reduce overlapped with the backward kernels)."""
    import torch.distributed as dist

    from diffusion_models_collection_b200 import synth
    from diffusion_models_collection_b200.diffusion import DDPM
    from diffusion_models_collection_b200.models import UNet

    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    torch.manual_seed(1234 + rank)
    net = UNet(**synth.CIFAR_UNET, num_classes=10)
    net.load_state_dict(synth.make_unet_state_dict(None, 10, seed=42))
    net = net.to(dev).train()
    ema = [p.detach().clone() for p in net.parameters()] if rank == 0 else None
    bucket_mb = float(os.environ.get("DMC_DDP_BUCKET_MB", "25"))  # 25 = torch default (the reference's setting)
    bucket_view = os.environ.get("DMC_DDP_BUCKET_VIEW", "0") == "1"  # default False = torch default (the reference's setting)
    native_ar = os.environ.get("DMC_NATIVE_ALLREDUCE", "0") == "1"  # UNet.set_gradient_allreduce() instead of DDP(model)
    if world > 1 and native_ar:
        model = net.set_gradient_allreduce()
    else:
        model = (torch.nn.parallel.DistributedDataParallel(net, bucket_cap_mb=bucket_mb, gradient_as_bucket_view=bucket_view)
                 if world > 1 else net)
    ddpm = DDPM(1000, 1e-4, 0.02, "linear", device=dev)
    # native clip + AdamW + EMA in two launches (optim.FusedAdamW, the default since round 2: it bumps the parameter versions,
    # so the engine re-packs its bf16 operands after every step); DMC_FUSED_OPT=0: torch's fused AdamW + foreach clip / EMA
    fused_opt = os.environ.get("DMC_FUSED_OPT", "1") == "1"
    if fused_opt:
        from diffusion_models_collection_b200.optim import FusedAdamW

        opt = FusedAdamW(net.parameters(), lr=2e-4, weight_decay=1e-4, max_grad_norm=1.0, ema_params=ema, ema_decay=0.9999)
    else:
        opt = torch.optim.AdamW(net.parameters(), lr=2e-4, weight_decay=1e-4, fused=True)
    B = args.batch
    g = torch.Generator().manual_seed(42 + rank)
    nbuf = 4  # rotating host batches (the DataLoader's pinned buffers)
    x_host = [torch.rand(B, 3, 32, 32, generator=g).mul_(2).sub_(1).pin_memory() for _ in range(nbuf)]
    y_host = [torch.randint(0, 10, (B,), generator=g).pin_memory() for _ in range(nbuf)]
    x_dev, y_dev = [t.to(dev) for t in x_host], [t.to(dev) for t in y_host]
    it = [0]

    def step(images, labels, read_loss):
        labels = labels + 1  # 0 is the null label (utils/trainer.py:225-230)
        drop = torch.rand(labels.shape, device=dev) < 0.2
        labels = torch.where(drop, torch.zeros_like(labels), labels)
        t = torch.randint(0, 1000, (B,), device=dev).long()
        loss = ddpm.p_losses(model, images, t, labels, loss_type="l2")
        loss.backward()
        if fused_opt:
            opt.step()
            opt.zero_grad()
            return loss.item() if read_loss else loss
        torch.nn.utils.clip_grad_norm_(net.parameters(), 1.0)
        opt.step()
        opt.zero_grad()
        if ema is not None:
            torch._foreach_mul_(ema, 0.9999)
            torch._foreach_add_(ema, [p.detach() for p in net.parameters()], alpha=1 - 0.9999)
        return loss.item() if read_loss else loss

    def step_resident():
        i = it[0] = (it[0] + 1) % nbuf
        return step(x_dev[i], y_dev[i], False)

    def step_e2e():
        i = it[0] = (it[0] + 1) % nbuf
        return step(x_host[i].to(dev, non_blocking=True), y_host[i].to(dev, non_blocking=True), True)

    def timed(fn, n):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
            dist.barrier()
        return float(ms.item())

    for _ in range(max(args.warmup, 3)):
        step_resident()
    with ClockSampler(local_rank) as cs:
        ms = timed(step_resident, args.steps)
    clocks = cs.summary()
    value = world * B * args.steps / (ms / 1e3)
    step_e2e()
    ms_e2e = timed(step_e2e, args.steps)
    e2e_value = world * B * args.steps / (ms_e2e / 1e3)
    loss_now = step_e2e()
    torch.cuda.synchronize()
    t0 = time.perf_counter()  # host enqueue time per step (no synchronisation inside): is the GPU ever waiting for Python?
    for _ in range(args.steps):
        step_resident()
    host_ms = (time.perf_counter() - t0) * 1e3 / args.steps
    torch.cuda.synchronize()

    eng = next(iter(net._train_engines.values()))
    info = eng.describe()
    line = {"metric": "train_unet_cifar10_images_per_sec", "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": "UNet cond (10+1 null) CIFAR-10 32x32 training step: CFG label dropout 0.2, t~U[0,1000), q_sample, "
                                   "eps-MSE, backward, clip_grad_norm 1.0, fused AdamW, EMA on rank 0 (BASELINE.json configs[4]); "
                                   "side measurement, not the headline metric",
                       "global_batch": world * B, "per_gpu_batch": B, "parallelism": (f"native per-entry all-reduce x{world} (NCCL, overlapped, no DDP wrapper)" if (world > 1 and os.environ.get("DMC_NATIVE_ALLREDUCE", "0") == "1") else (f"DDP(model) x{world}: the wrapper is handed all parameters but one to ignore, the engine's per-entry NCCL all-reduce averages them (overlapped)" if (world > 1 and getattr(net, "_ddp_sentinel", None)) else f"DDP x{world} (NCCL all-reduce overlapped)")),
                       "l2": "activations + gradients of one step are > 10x the 126 MB L2", "dropout": 0.1},
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(B * (3 * 32 * 32 * 4 + 8)), "d2h_bytes_per_step": 4},
            "gpu_launches": int((info["forward_launches"] + info["backward_launches"] + (3 if fused_opt else 0)) * args.steps),
            "loss": loss_now, "host_enqueue_ms_per_step": host_ms}
    line["config"]["optimizer"] = "FusedAdamW (native clip + AdamW + EMA)" if fused_opt else "torch fused AdamW + foreach clip / EMA"
    line["config"]["ddp_bucket_mb"] = bucket_mb if world > 1 else None
    line["config"]["ddp_gradient_as_bucket_view"] = bucket_view if world > 1 else None
    pk = peaks()
    flops = info["gemm_flops"]
    line["model_flops_utilization"] = {"achieved_tflops": flops * args.steps / (ms / 1e3) / 1e12, "gemm_flops_per_step": flops,
                                       "peak_tflops": pk["bf16_tflops_sustained"], "peak_source": pk["_source"]}
    if rank == 0 and not args.no_roofline:
        ops = eng.time_ops(iters=3)
        if args.ops_out:
            json.dump(ops, open(args.ops_out, "w"), indent=1)
        tc = [o for o in ops if o["flops"] > 0]
        tms, tfl = sum(o["ms"] for o in tc), sum(o["flops"] for o in tc)
        line["roofline"] = {"bound": "tensor", "achieved": tfl / (tms / 1e3) / 1e12, "peak": pk["bf16_tflops"], "unit": "TFLOP/s",
                            "frac": tfl / (tms / 1e3) / 1e12 / pk["bf16_tflops"], "traffic": None,
                            "kernel": "tcgen05 convolution family of one training step: forward + input-gradient (same kernel) + "
                                      "weight-gradient GEMMs, each timed alone (burst peak)",
                            "tensor_ms": tms, "all_ops_ms": sum(o["ms"] for o in ops), "step_ms": ms / args.steps,
                            "by_kind_ms": {k: sum(o["ms"] for o in ops if o["kind"] == k) for k in sorted({o["kind"] for o in ops})}}
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        _, line["cpu_baseline"] = cpu_reference_train_images_per_sec(batch=2 * args.ref_batch, steps=2)
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def ncu_traffic(nimg):
    """DRAM bytes per launch of the dominant kernel from the committed ncu capture (profiles/*_ncu_*.json written by
    tools/summarize_ncu.py), if one exists for this launch size."""
    import glob

    best = None
    import re

    def order(path):  # r02_forward_ncu_run20.json after r02_forward_ncu_run8.json: (round, run) as numbers
        return [int(v) for v in re.findall(r"\d+", os.path.basename(path))]

    for f in sorted(glob.glob(os.path.join(ROOT, "profiles", "*forward_ncu*.json")), key=order):
        try:
            d = json.load(open(f))
        except Exception:
            continue
        if d.get("images_per_launch") == nimg:
            best = d
    return best


def roofline_leg(net, dev, nb, pk, ops_out, cfg=True, step_ms=None, chunks=1):
    """Per-op device times of ONE forward (CUDA events on the launching stream, each op timed alone -> burst peak),
    aggregated per kernel family.  Dominant kernel: the tcgen05 implicit-GEMM convolution (all GEMM-shaped launches of
    one forward, FLOP-weighted: achieved = sum of algorithmic FLOPs / sum of launch durations)."""
    mult = 2 if cfg else 1
    nimg = min(mult * nb, net.max_images_per_launch)
    with net.uniform_timesteps():  # what the sampling loop runs
        plan = net.plan_info(nimg // mult, cfg=cfg, device=dev)
    ops = plan.time_ops(iters=3)
    fam = {}
    for o in ops:
        f = fam.setdefault(o["kind"], {"ms": 0.0, "flops": 0.0, "bytes": 0.0, "launches": 0})
        f["ms"] += o["ms"]
        f["flops"] += o["flops"]
        f["bytes"] += o["bytes"]
        f["launches"] += 1
    total_ms = sum(f["ms"] for f in fam.values())
    conv = fam.get("conv", {"ms": 1e-9, "flops": 0.0, "launches": 1})
    achieved = conv["flops"] / (conv["ms"] / 1e3) / 1e12
    peak = pk["bf16_tflops"]
    per_family = {}
    for k, f in fam.items():
        e = {"ms": round(f["ms"], 4), "share": round(f["ms"] / total_ms, 4), "launches": f["launches"]}
        if f["flops"] > 0 and k in ("conv", "attention", "gemm"):
            e["tflops"] = round(f["flops"] / (f["ms"] / 1e3) / 1e12, 2)
            e["frac_tensor"] = round(e["tflops"] / peak, 4)
            if k == "attention" and f["bytes"] > 0:
                # head dim 64: 128 FLOP per HBM byte at L = 256, half the chip's ridge -> the attention core is HBM-bound and
                # its roofline fraction is the bandwidth one ("Attention")
                e["gbs"] = round(f["bytes"] / (f["ms"] / 1e3) / 1e9, 1)
                e["frac_hbm"] = round(e["gbs"] / pk["hbm_gbs"], 4)
                e["bound"] = "hbm"
        else:
            e["gbs"] = round(f["bytes"] / (f["ms"] / 1e3) / 1e9, 1)
            e["frac_hbm"] = round(e["gbs"] / pk["hbm_gbs"], 4)
        per_family[k] = e
    if ops_out:
        with open(ops_out, "w") as fh:
            json.dump({"images": nimg, "ops": ops, "families": per_family}, fh, indent=1)
    out = {"bound": "tensor", "kernel": "conv_umma_kernel (all conv / GEMM launches of one forward, FLOP-weighted)",
           "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak, "traffic": None,
           "peak_source": pk["_source"] + " burst (ops timed alone)", "images_per_launch": nimg,
           "launches_per_forward": conv["launches"], "flops_per_launch": conv["flops"] / max(1, conv["launches"]),
           "forward_ms": total_ms, "per_family": per_family}
    tr = ncu_traffic(nimg)
    if tr is not None:
        out["traffic"] = tr["dram_bytes_per_launch"]
        out["traffic_algorithmic"] = tr["algorithmic_bytes_per_launch"]
        out["traffic_source"] = "profiles/" + tr["source"]
    if step_ms is not None:
        # the same kernels inside the timed sampling loop (power-capped clocks): time of one denoising step of one chunk,
        # apportioned by the per-family shares measured above -> fraction of the SUSTAINED peak
        in_step_ms = step_ms / max(1, chunks) * (conv["ms"] / total_ms)
        a = conv["flops"] / (in_step_ms / 1e3) / 1e12
        out["in_step"] = {"achieved": a, "peak": pk["bf16_tflops_sustained"], "frac": a / pk["bf16_tflops_sustained"],
                          "ms_per_forward": step_ms / max(1, chunks), "note": "step time x conv share of the per-op timing"}
    return out


if __name__ == "__main__":
    main()
