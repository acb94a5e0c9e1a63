"""CPU: the reference's UNMODIFIED sample.py, launched through dropin/run.py, resolves `models`, `diffusion`,
`utils.helpers` to the native classes: a synthetic checkpoint in the reference's format loads (strict state_dict
contract), the sampler is built, and -- there being no GPU in the build container -- the first denoising step fails
LOUDLY with DmcError instead of falling back to a CPU implementation.  Skipped where /root/reference is absent."""

import os
import subprocess
import sys

import pytest
import torch

from diffusion_models_collection_b200 import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("DMC_REFERENCE_DIR", "/root/reference")


@pytest.mark.skipif(not os.path.exists(os.path.join(REF, "sample.py")), reason="reference checkout not present")
@pytest.mark.skipif(torch.cuda.is_available(), reason="CPU-only check (on a GPU box the run would go through)")
def test_reference_sample_script_runs_through_the_shims(tmp_path):
    import importlib.util

    spec = importlib.util.spec_from_file_location("_refcfg", os.path.join(REF, "configs", "cifar10_unet.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    cfg = dict(mod.config)
    ncls = cfg["num_classes"] if cfg.get("conditional") else None
    ck = tmp_path / "ckpt.pth"
    torch.save({"epoch": 1, "model_state_dict": synth.make_unet_state_dict(None, ncls, seed=1), "config": cfg}, ck)
    r = subprocess.run([sys.executable, os.path.join(ROOT, "dropin", "run.py"), os.path.join(REF, "sample.py"),
                        "--checkpoint", str(ck), "--sampling_method", "ddim", "--num_inference_steps", "3", "--device", "cpu"],
                       cwd=tmp_path, capture_output=True, text=True, timeout=300)
    out = r.stdout + r.stderr
    assert r.returncode != 0
    assert "diffusion_models_collection_b200" in out          # our classes were the ones imported
    assert "DmcError" in out and "no CPU fallback" in out    # and they refuse to run without the CUDA path
    assert "Generating" in out                                # checkpoint loaded, model and sampler constructed


@pytest.mark.skipif(not os.path.exists(os.path.join(REF, "train.py")), reason="reference checkout not present")
@pytest.mark.skipif(torch.cuda.is_available(), reason="CPU-only check (on a GPU box the run would go on to train)")
def test_reference_train_script_resolves_through_the_shims(tmp_path):
    """the reference's UNMODIFIED train.py through dropin/run.py: `models` / `diffusion` / `utils.helpers` resolve to the native
    classes, `utils.trainer` to the reference's own trainer (a caller of the hot path, loaded by file path), `datasets` to the
    reference's package; without a GPU the script stops at its own GPU check in main() -- after every import succeeded"""
    r = subprocess.run([sys.executable, os.path.join(ROOT, "dropin", "run.py"), os.path.join(REF, "train.py"),
                        "--config", os.path.join(REF, "configs", "cifar10_unet.py")],
                       cwd=tmp_path, capture_output=True, text=True, timeout=300,
                       env={**os.environ, "WORLD_SIZE": "1", "DMC_REFERENCE_DIR": REF})
    out = r.stdout + r.stderr
    assert r.returncode != 0
    assert "GPU 0 not available. Only 0 GPU(s) detected." in out or "CUDA not available" in out, out[-3000:]
    assert "ModuleNotFoundError" not in out and "ImportError" not in out


def test_trainer_shim_reexports_the_reference_trainer():
    if not os.path.exists(os.path.join(REF, "utils", "trainer.py")):
        pytest.skip("reference checkout not present")
    code = ("import sys; sys.path[:0] = [%r, %r]; import utils.trainer as t, models; "
            "print(t.DiffusionTrainer.__module__, models.UNet.__module__)" % (os.path.join(ROOT, "dropin"), ROOT))
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300,
                       env={**os.environ, "DMC_REFERENCE_DIR": REF})
    assert r.returncode == 0, r.stderr[-3000:]
    assert "_dmc_reference_trainer" in r.stdout and "diffusion_models_collection_b200" in r.stdout


_TRAINER_SCRIPT = r'''
import contextlib, os, sys, torch
sys.path[:0] = [DROPIN, ROOT]
from diffusion_models_collection_b200 import _lib
from tests.test_train_plan import _Recorder
rec = _Recorder()
_lib.load, _lib.stream_ptr, _lib.check = (lambda: rec), (lambda: 0), (lambda rc, what="": rc)
torch.cuda.device = lambda d: contextlib.nullcontext()
from diffusion_models_collection_b200.models import unet_train
unet_train.USE_GRAPHS = False
from models import UNet                     # the shim -> native class
from diffusion import DDPM
from utils.trainer import DiffusionTrainer  # the shim -> the reference's own trainer
assert UNet.__module__.startswith("diffusion_models_collection_b200")
# host-logic run: no device, so route the forward straight to the training engine (its kernels are recorded, not run)
UNet._run = lambda self, x, t, y, cfg: self._run_train(x, t, y)
DDPM.q_sample = lambda self, x_start, t, noise=None: x_start  # (the q_sample kernel refuses CPU tensors, like every product path)
torch.manual_seed(0)
mp = dict(image_size=(32, 32), in_channels=3, model_channels=128, out_channels=3, num_res_blocks=2, attention_resolutions=(16, 8),
          dropout=0.1, channel_mult=(1, 2, 2, 2))
cfg = dict(epochs=1, save_dir=os.path.join(OUT, "ck"), sample_dir=os.path.join(OUT, "img"), loss_type="l2", use_ema=True,
           conditional=True, num_classes=10, image_size=32, model_type="unet", model_params=mp, sample_start_epoch=99,
           cfg_dropout_prob=0.2, save_interval=1)
model = UNet(**mp, num_classes=10)
ds = torch.utils.data.TensorDataset(torch.rand(8, 3, 32, 32) * 2 - 1, torch.randint(0, 10, (8,)))
loader = torch.utils.data.DataLoader(ds, batch_size=4, drop_last=True)
opt = torch.optim.AdamW(model.parameters(), lr=2e-4, weight_decay=1e-4)
tr = DiffusionTrainer(model=model, diffusion=DDPM(1000, 1e-4, 0.02, "linear", device="cpu"), train_loader=loader, optimizer=opt,
                      scheduler=None, device=torch.device("cpu"), config=cfg, rank=0, world_size=1)
assert type(tr.ema_model) is UNet and not any(p.requires_grad for p in tr.ema_model.parameters())
tr.train()
ck = torch.load(os.path.join(OUT, "ck", "current_model.pth"), weights_only=False)
assert set(ck) >= {"epoch", "model_state_dict", "optimizer_state_dict", "best_loss", "config", "ema_model_state_dict"}
assert len(ck["model_state_dict"]) == 357 and len(ck["ema_model_state_dict"]) == 357
n_fwd = sum(1 for n, _ in rec.calls if n == "dmc_plan_run")
n_wg = sum(1 for n, _ in rec.calls if n == "dmc_conv_wgrad")
print("TRAINER_OK", n_fwd, n_wg)
'''


def test_reference_trainer_drives_the_native_training_engine(tmp_path):
    """the reference's own DiffusionTrainer (utils/trainer.py, unmodified) around the native UNet for one epoch of two iterations:
    EMA model built with type(model)(**params), labels + 1 with CFG dropout, p_losses, backward, clip, AdamW, EMA update and
    checkpoint -- host logic only (the C library is a recorder: kernels are logged, not run; numbers are meaningless)"""
    if not os.path.exists(os.path.join(REF, "utils", "trainer.py")):
        pytest.skip("reference checkout not present")
    code = (f"DROPIN, ROOT, OUT = {os.path.join(ROOT, 'dropin')!r}, {ROOT!r}, {str(tmp_path)!r}\n" + _TRAINER_SCRIPT)
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=900, cwd=tmp_path,
                       env={**os.environ, "DMC_REFERENCE_DIR": REF})
    assert r.returncode == 0, (r.stdout[-2000:] + r.stderr[-4000:])
    tag = [l for l in r.stdout.splitlines() if l.startswith("TRAINER_OK")]
    assert tag, r.stdout[-2000:]
    _, n_fwd, n_wg = tag[0].split()
    assert int(n_fwd) == 2 and int(n_wg) == 2 * 99  # two iterations: two forward plans, 99 weight-gradient launches each
