"""Writes the golden fixtures in this directory from the LIVE reference (/root/reference).

Run in the build container only (the GPU box has no /root/reference):

    python tests/golden/make_golden.py

The reference has no tests and no golden vectors of its own, so the pin for the
oracle -- and through it for the CUDA path -- is the reference itself, imported here file by file
under private module names (its top-level package names ``models`` / ``diffusion`` would otherwise
clash with the drop-in shims).  Nothing from the reference is copied: only inputs and outputs of its
public API are stored.

Fixtures written:
  unet_golden.npz    eps = UNet(x, t, y) for the cases in UNET_CASES (weights from synth.py seeds,
                     loaded with strict=True => state_dict key/shape contract is checked here)
  dit_golden.npz     eps = DiT(x, t, y) for DIT_CASES
  tables_golden.npz  every coefficient table of DDPM/DDIM for linear / cosine / quadratic schedules,
                     and the DDIM timestep subsets
  steps_golden.npz   single p_sample calls (DDIM eta 0 / 0.5, last step, DDPM t>0 / t==0), q_sample
  loops_golden.npz   whole sample() / sample_with_cfg() runs with a toy denoiser and recorded noise
  train_golden.npz   loss and parameter gradients of the reference's own DDPM.p_losses(UNet, ...) + backward() (eval mode: dropout
                     off) for two small batches: per-tensor l2 norm and sum for all 357 / 334 tensors, every 1-D tensor in full,
                     256 fixed entries of every larger tensor -- the pin of the training step (BASELINE configs[4])
  config1_golden.npz BASELINE configs[0] verbatim (default init under seed 42, uncond DDIM-50, batch 16): x_T, the reference's
                     state after 1 / 2 / 3 / 5 / 10 / 20 / 50 free-running steps, its first eps (a B = 16 whole-model golden)
  samples_golden.npz FINAL images of whole DDIM-50 runs of the REAL CIFAR UNet / DiT through the reference's own
                     sample() / sample_with_cfg() (fp32 CPU, x_T recorded): the end-to-end pin of BASELINE configs 1 / 3 / 4
"""

from __future__ import annotations

import importlib.util
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
REF = os.environ.get("DMC_REFERENCE_DIR", "/root/reference")

from diffusion_models_collection_b200 import synth  # noqa: E402
from oracle.sched_oracle import toy_model  # noqa: E402
from tests.golden_cases import (DIT_CASES, UNET_CASES, SMALL_UNET, TRAIN_CASES, TRAIN_DIT_CASES, case_inputs,  # noqa: E402
                                perturbed_state_dict, sample_index, train_inputs)


def _load(name, rel):
    spec = importlib.util.spec_from_file_location(name, os.path.join(REF, rel))
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


ref_unet = _load("_ref_unet", "models/unet.py")
ref_dit = _load("_ref_dit", "models/dit.py")
ref_ddpm = _load("_ref_ddpm", "diffusion/ddpm.py")
ref_ddim = _load("_ref_ddim", "diffusion/ddim.py")


class NoiseRecorder:
    """Records every torch.randn / torch.randn_like draw made while active (the reference API has no
    noise-injection argument, SURVEY.md section 8c trap (iv))."""

    def __init__(self):
        self.draws = []

    def __enter__(self):
        self._randn, self._randn_like = torch.randn, torch.randn_like

        def randn(*a, **k):
            out = self._randn(*a, **k)
            self.draws.append(out.detach().cpu().clone())
            return out

        def randn_like(*a, **k):
            out = self._randn_like(*a, **k)
            self.draws.append(out.detach().cpu().clone())
            return out

        torch.randn, torch.randn_like = randn, randn_like
        return self

    def __exit__(self, *exc):
        torch.randn, torch.randn_like = self._randn, self._randn_like


def gen_unet():
    out = {}
    for name, c in UNET_CASES.items():
        cfg = SMALL_UNET if c.get("small") else synth.CIFAR_UNET
        sd = synth.make_unet_state_dict(cfg, c["num_classes"], seed=c["wseed"], null_row_zero=c.get("null_row_zero", True))
        net = ref_unet.UNet(**cfg, num_classes=c["num_classes"]).eval()
        net.load_state_dict(sd, strict=True)
        x, t, y = case_inputs(c)
        with torch.no_grad():
            eps = net(x, t, y)
        out[name] = eps.numpy()
        print("unet", name, tuple(eps.shape), float(eps.std()))
    np.savez_compressed(os.path.join(HERE, "unet_golden.npz"), **out)


TEACHER_STEPS = [0, 1, 10, 25, 40, 48, 49]  # DDIM-50 steps whose (x_in, x_out) pair is stored for teacher forcing


def gen_samples():
    """whole sampling runs of the real models through the reference's own loops (tqdm bars silenced): the final images and,
    for a few steps s, the reference's own state before and after step s (teacher-forcing pairs)"""
    import contextlib
    import io

    out = {}

    def run(name, net, fn):
        torch.manual_seed(1234)
        with NoiseRecorder() as rec, contextlib.redirect_stderr(io.StringIO()), torch.no_grad():
            traj = fn(net)  # [S, B, C, H, W]: the state after every step (return_all_timesteps=True)
        xT = rec.draws[0]
        out[name] = traj[-1].numpy()
        out[name + ".xT"] = xT.numpy()
        for s_ in TEACHER_STEPS:
            out[f"{name}.in{s_}"] = (xT if s_ == 0 else traj[s_ - 1]).numpy()
            out[f"{name}.out{s_}"] = traj[s_].numpy()
        print("samples", name, tuple(traj.shape), float(traj[-1].abs().max()), float(traj[-1].std()))

    d50 = ref_ddim.DDIM(1000, 50, 1e-4, 0.02, "linear", eta=0.0, device="cpu")
    net = ref_unet.UNet(**synth.CIFAR_UNET, num_classes=None).eval()
    net.load_state_dict(synth.make_unet_state_dict(None, None, seed=42), strict=True)
    run("unet.uncond.ddim50", net, lambda m: d50.sample(m, (2, 3, 32, 32), return_all_timesteps=True))
    net = ref_unet.UNet(**synth.CIFAR_UNET, num_classes=10).eval()
    net.load_state_dict(synth.make_unet_state_dict(None, 10, seed=42), strict=True)
    y = torch.tensor([3, 10])
    run("unet.cond.ddim50.cfg3", net, lambda m: d50.sample_with_cfg(m, (2, 3, 32, 32), y, cfg_scale=3.0,
                                                                    return_all_timesteps=True))
    net = ref_dit.DiT(**synth.CIFAR_DIT, num_classes=None).eval()
    net.load_state_dict(synth.make_dit_state_dict(None, None, seed=42), strict=True)
    run("dit.uncond.ddim50", net, lambda m: d50.sample(m, (2, 3, 32, 32), return_all_timesteps=True))
    np.savez_compressed(os.path.join(HERE, "samples_golden.npz"), **out)


def gen_dit():
    out = {}
    for name, c in DIT_CASES.items():
        cfg = synth.CIFAR_DIT
        sd = synth.make_dit_state_dict(cfg, c["num_classes"], seed=c["wseed"])
        net = ref_dit.DiT(**cfg, num_classes=c["num_classes"]).eval()
        net.load_state_dict(sd, strict=True)
        x, t, y = case_inputs(c)
        with torch.no_grad():
            eps = net(x, t, y)
        out[name] = eps.numpy()
        print("dit", name, tuple(eps.shape), float(eps.std()))
    np.savez_compressed(os.path.join(HERE, "dit_golden.npz"), **out)


def gen_dim():
    """eps = DiM(x, t, y) of the reference's own DiM class as built in this container (mamba_ssm absent: models/dim.py:103-117
    picks nn.MultiheadAttention)"""
    from tests.golden_cases import DIM_CASES

    ref_dim = _load("_ref_dim", "models/dim.py")
    assert not ref_dim.MAMBA_AVAILABLE
    out = {}
    for name, c in DIM_CASES.items():
        cfg = dict(synth.CIFAR_DIM, hidden_size=c["hidden"], depth=c["depth"])
        net = ref_dim.DiM(**cfg, num_classes=c["num_classes"]).eval()
        net.load_state_dict(synth.make_dim_state_dict(cfg, c["num_classes"], seed=c["wseed"]), strict=True)
        x, t, y = case_inputs(c)
        with torch.no_grad():
            eps = net(x, t, y)
        out[name] = eps.numpy()
        print("dim", name, tuple(eps.shape), float(eps.std()))
    np.savez_compressed(os.path.join(HERE, "dim_golden.npz"), **out)


TABLE_NAMES = ["betas", "alphas", "alphas_cumprod", "alphas_cumprod_prev", "sqrt_alphas_cumprod",
               "sqrt_one_minus_alphas_cumprod", "sqrt_recip_alphas", "sqrt_recipm1_alphas_cumprod",
               "posterior_variance", "posterior_log_variance_clipped", "posterior_mean_coef1", "posterior_mean_coef2"]


def gen_tables():
    out = {}
    for sched in ("linear", "cosine", "quadratic"):
        d = ref_ddpm.DDPM(1000, 1e-4, 0.02, sched, device="cpu")
        for n in TABLE_NAMES:
            out[f"{sched}.{n}"] = getattr(d, n).numpy()
        di = ref_ddim.DDIM(1000, 50, 1e-4, 0.02, sched, device="cpu")
        for n in ("betas", "alphas", "alphas_cumprod", "sqrt_alphas_cumprod", "sqrt_one_minus_alphas_cumprod"):
            assert torch.equal(getattr(di, n), getattr(d, n)), (sched, n)
    d = ref_ddpm.DDPM(20, 1e-4, 0.02, "linear", device="cpu")
    for n in TABLE_NAMES:
        out[f"linear20.{n}"] = getattr(d, n).numpy()
    for T, S in ((1000, 50), (1000, 10), (1000, 1000), (1000, 7), (20, 7), (1000, 1)):
        out[f"timesteps.{T}.{S}"] = ref_ddim.DDIM(T, S, device="cpu").inference_timesteps.numpy()
    di = ref_ddim.DDIM(1000, 50, device="cpu")
    di.set_inference_steps(25)
    out["timesteps.set25"] = di.inference_timesteps.numpy()
    np.savez_compressed(os.path.join(HERE, "tables_golden.npz"), **out)
    print("tables", len(out))


def gen_steps():
    out = {}
    g = torch.Generator().manual_seed(1234)
    B = 2
    x = torch.randn(B, 3, 32, 32, generator=g)
    eps = torch.randn(B, 3, 32, 32, generator=g)
    out["x"], out["eps"] = x.numpy(), eps.numpy()
    fake = lambda *_a, **_k: eps  # noqa: E731
    full = lambda v: torch.full((B,), v, dtype=torch.long)  # noqa: E731
    for eta in (0.0, 0.5):
        di = ref_ddim.DDIM(1000, 50, eta=eta, device="cpu")
        for (t, tn) in ((999, 979), (510, 489), (20, 0), (0, -1)):
            for clip in (True, False):
                torch.manual_seed(7)
                with NoiseRecorder() as rec:
                    r = di.p_sample(fake, x, full(t), full(tn), None, clip_denoised=clip)
                key = f"ddim.eta{eta}.t{t}.n{tn}.clip{int(clip)}"
                out[key] = r.numpy()
                if rec.draws:
                    out[key + ".noise"] = rec.draws[0].numpy()
    dp = ref_ddpm.DDPM(1000, device="cpu")
    for t in (999, 500, 1, 0):
        for clip in (True, False):
            torch.manual_seed(11)
            with NoiseRecorder() as rec:
                r = dp.p_sample(fake, x, full(t), None, clip_denoised=clip)
            key = f"ddpm.t{t}.clip{int(clip)}"
            out[key] = r.numpy()
            out[key + ".noise"] = rec.draws[0].numpy()
    tq = torch.tensor([17, 999])
    out["q_sample.t"] = tq.numpy()
    out["q_sample.ddpm"] = dp.q_sample(x, tq, eps).numpy()
    out["q_sample.ddim"] = ref_ddim.DDIM(1000, 50, device="cpu").q_sample(x, tq, eps).numpy()
    np.savez_compressed(os.path.join(HERE, "steps_golden.npz"), **out)
    print("steps", len(out))


def gen_loops():
    out = {}
    B = 2
    shape = (B, 3, 32, 32)
    y = torch.tensor([3, 10])

    def run(key, fn):
        torch.manual_seed(42)
        with NoiseRecorder() as rec:
            r = fn()
        out[key] = r.numpy()
        out[key + ".noise"] = torch.stack(rec.draws).numpy()
        print("loop", key, tuple(r.shape), len(rec.draws))

    d50 = ref_ddim.DDIM(1000, 50, device="cpu")
    d10 = ref_ddim.DDIM(1000, 10, device="cpu")
    d10e = ref_ddim.DDIM(1000, 10, eta=0.3, device="cpu")
    run("ddim50.sample", lambda: d50.sample(toy_model, shape))
    run("ddim50.sample_y", lambda: d50.sample(toy_model, shape, y))
    run("ddim50.cfg3", lambda: d50.sample_with_cfg(toy_model, shape, y, cfg_scale=3.0))
    run("ddim10.sample.traj", lambda: d10.sample(toy_model, shape, return_all_timesteps=True))
    run("ddim10.cfg3.traj", lambda: d10.sample_with_cfg(toy_model, shape, y, cfg_scale=3.0, return_all_timesteps=True))
    run("ddim10.cfg1p5.nothr", lambda: d10.sample_with_cfg(toy_model, shape, y, cfg_scale=1.5, p_threshold=None))
    run("ddim10.cfg3.p90", lambda: d10.sample_with_cfg(toy_model, shape, y, cfg_scale=3.0, p_threshold=0.9))
    run("ddim10.eta0p3.sample", lambda: d10e.sample(toy_model, shape, y))
    run("ddim10.eta0p3.cfg3", lambda: d10e.sample_with_cfg(toy_model, shape, y, cfg_scale=3.0))
    p20 = ref_ddpm.DDPM(20, device="cpu")
    run("ddpm20.sample.traj", lambda: p20.sample(toy_model, shape, y, return_all_timesteps=True))
    run("ddpm20.cfg3.traj", lambda: p20.sample_with_cfg(toy_model, shape, y, cfg_scale=3.0, return_all_timesteps=True))
    run("ddpm20.cfg2.nothr", lambda: p20.sample_with_cfg(toy_model, shape, y, cfg_scale=2.0, p_threshold=None))
    # error behaviour (ddim.py:270-273, ddpm.py:273-276, ddpm.py:46)
    for bad in (lambda: d10.sample_with_cfg(toy_model, shape, None),
                lambda: d10.sample_with_cfg(toy_model, shape, y, p_threshold=1.0),
                lambda: p20.sample_with_cfg(toy_model, shape, None),
                lambda: ref_ddpm.DDPM(10, beta_schedule="nope", device="cpu")):
        try:
            bad()
            raise SystemExit("reference did not raise")
        except ValueError:
            pass
    np.savez_compressed(os.path.join(HERE, "loops_golden.npz"), **out)


def gen_train():
    out = {}
    for name, c in TRAIN_CASES.items():
        net = ref_unet.UNet(**synth.CIFAR_UNET, num_classes=c["num_classes"]).eval()
        net.load_state_dict(perturbed_state_dict(c["num_classes"]), strict=True)
        ddpm = ref_ddpm.DDPM(num_timesteps=1000, beta_start=1e-4, beta_end=0.02, beta_schedule="linear", device="cpu")
        x0, t, y, noise = train_inputs(c)
        loss = ddpm.p_losses(net, x0, t, y, noise=noise, loss_type="l2")
        loss.backward()
        out[name + "/loss"] = np.float64(loss.item())
        names, norms, sums = [], [], []
        for k, p in net.named_parameters():
            g = p.grad
            names.append(k)
            norms.append(float(g.double().norm()))
            sums.append(float(g.double().sum()))
            flat = g.reshape(-1)
            if g.dim() == 1:
                out[f"{name}/full/{k}"] = flat.numpy().copy()
            else:
                out[f"{name}/sample/{k}"] = flat[torch.from_numpy(sample_index(flat.numel()))].numpy().copy()
        out[name + "/names"] = np.array(names)
        out[name + "/norms"] = np.array(norms)
        out[name + "/sums"] = np.array(sums)
        print("train", name, "loss", loss.item(), "params", len(names), "total grad norm", float(np.sqrt((np.array(norms) ** 2).sum())))
    np.savez_compressed(os.path.join(HERE, "train_golden.npz"), **out)


def gen_train_dit():
    """train_dit_golden.npz: loss and parameter gradients of the reference's own DDPM.p_losses(DiT, ...) + backward() (eval mode)"""
    out = {}
    for name, c in TRAIN_DIT_CASES.items():
        cfg = synth.CIFAR_DIT
        net = ref_dit.DiT(**cfg, num_classes=c["num_classes"]).eval()
        net.load_state_dict(synth.make_dit_state_dict(cfg, c["num_classes"], seed=c["wseed"]), strict=True)
        ddpm = ref_ddpm.DDPM(num_timesteps=1000, beta_start=1e-4, beta_end=0.02, beta_schedule="linear", device="cpu")
        x0, t, y, noise = train_inputs(c)
        loss = ddpm.p_losses(net, x0, t, y, noise=noise, loss_type="l2")
        loss.backward()
        out[name + "/loss"] = np.float64(loss.item())
        names, norms = [], []
        for k, p in net.named_parameters():
            g = p.grad
            names.append(k)
            norms.append(float(g.double().norm()))
            flat = g.reshape(-1)
            if g.dim() == 1 and flat.numel() <= 4096:
                out[f"{name}/full/{k}"] = flat.numpy().copy()
            else:
                out[f"{name}/sample/{k}"] = flat[torch.from_numpy(sample_index(flat.numel()))].numpy().copy()
        out[name + "/names"] = np.array(names)
        out[name + "/norms"] = np.array(norms)
        print("train_dit", name, "loss", loss.item(), "params", len(names), "total grad norm", float(np.sqrt((np.array(norms) ** 2).sum())))
    np.savez_compressed(os.path.join(HERE, "train_dit_golden.npz"), **out)


CONFIG1_HORIZONS = [1, 2, 3, 5, 10, 20, 50]  # states after this many free-running DDIM steps


def gen_config1():
    """BASELINE.json configs[0] through the reference: set_seed(42) (sample.py:90), UNet(**cifar10_unet model_params,
    num_classes=None), DDIM(1000, 50, 1e-4, 0.02, 'linear', eta=0).sample(model, (16, 3, 32, 32)).  One departure from the
    verbatim config: the weights are synth.make_unet_state_dict(seed 42) instead of PyTorch's default init -- nn.init's
    un-seeded-generator path is not bit-reproducible across CPU models (measured: the GPU box's host draws other values than the
    build container for the same seed), the explicit-generator draws of synth.py are, and 148 MB of weights are no fixture.
    Stored: x_T, the reference's state after CONFIG1_HORIZONS free-running steps (the last one = the final images), its eps for
    the first forward (a B = 16 whole-model golden), and per-tensor checksums of the weights."""
    import contextlib
    import io
    import random

    random.seed(42)
    np.random.seed(42)
    torch.manual_seed(42)
    net = ref_unet.UNet(**synth.CIFAR_UNET, num_classes=None).eval()
    net.load_state_dict(synth.make_unet_state_dict(None, None, seed=42), strict=True)
    d50 = ref_ddim.DDIM(1000, 50, 1e-4, 0.02, "linear", eta=0.0, device="cpu")
    with NoiseRecorder() as rec, contextlib.redirect_stderr(io.StringIO()), torch.no_grad():
        traj = d50.sample(net, (16, 3, 32, 32), return_all_timesteps=True)
    xT = rec.draws[0]
    out = {"xT": xT.numpy(), "horizons": np.array(CONFIG1_HORIZONS)}
    for h in CONFIG1_HORIZONS:
        out[f"after{h}"] = traj[h - 1].numpy()
    with torch.no_grad():
        out["eps0"] = net(xT, torch.full((16,), 999, dtype=torch.long)).numpy()
    names = [k for k, _ in net.state_dict().items()]
    out["weight_names"] = np.array(names)
    # exact, order-independent checksums (integer sums of the fp32 bit patterns): a floating-point .sum() depends on the SIMD
    # width of the host CPU and differed between the build container and the GPU box for IDENTICAL weights
    out["weight_isums"] = np.array([int(v.contiguous().view(torch.int32).to(torch.int64).sum()) for v in net.state_dict().values()])
    np.savez_compressed(os.path.join(HERE, "config1_golden.npz"), **out)
    print("config1", tuple(traj.shape), float(traj[-1].abs().max()), float(traj[-1].std()), "eps0 std", float(out["eps0"].std()))


if __name__ == "__main__":
    torch.set_num_threads(os.cpu_count() or 1)
    which = sys.argv[1:] or ["tables", "steps", "loops", "unet", "dit", "samples"]
    for w in which:
        globals()["gen_" + w]()
