"""GPU: the reference's UNMODIFIED sample.py (sample.py:67-279), executed through dropin/run.py on the B200 -- SURVEY.md
section 8 row f1.  The script comes from the reference itself (oracle/ref_loader.py: the checkout in the build container, the
archive oracle/_ref/reference_src.zip on the GPU box); `models`, `diffusion`, `utils.helpers`, `configs` resolve to the native
classes through dropin/.  Covered: checkpoint format (utils/trainer.py:339-351) incl. EMA weight selection, label + 1 shift,
DDIM-50 + CFG 3.0 with dynamic thresholding, the `return_all_timesteps` path (--create_gif / --save_intermediate: per-step
`.cpu()` copies, GIF + PNG frames), DDPM, and the DiT config.  The images the script writes are compared with the same call
made directly through the native API under the same seed (bit-equal after PNG quantisation)."""

import os
import subprocess
import sys

import numpy as np
import pytest
import torch

from diffusion_models_collection_b200 import synth

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ref_dir():
    from oracle import ref_loader

    if not ref_loader.available():
        pytest.skip("reference not available (no checkout, no oracle/_ref archive)")
    return ref_loader.reference_dir()


def _config(ref, which):
    import importlib.util

    spec = importlib.util.spec_from_file_location("_refcfg_" + which, os.path.join(ref, "configs", which + ".py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return dict(mod.config)


def _run(ref, tmp_path, args, timeout=900):
    r = subprocess.run([sys.executable, os.path.join(ROOT, "dropin", "run.py"), os.path.join(ref, "sample.py")] + args,
                       cwd=tmp_path, capture_output=True, text=True, timeout=timeout,
                       env={**os.environ, "DMC_REFERENCE_DIR": ref})
    out = r.stdout + r.stderr
    assert r.returncode == 0, out[-4000:]
    assert "Done!" in out
    return out


def _png(path):
    from PIL import Image

    return np.asarray(Image.open(path).convert("RGB"))


def _grid_png(images, nrow, path):
    from torchvision.utils import save_image

    save_image(torch.clamp((images.cpu() + 1) / 2, 0, 1), str(path), nrow=nrow)
    return _png(path)


def test_sample_py_ddim50_cfg3_labels_gif_on_b200(tmp_path):
    ref = _ref_dir()
    cfg = _config(ref, "cifar10_unet")
    cfg["conditional"], cfg["num_classes"] = True, 10
    sd = synth.make_unet_state_dict(None, 10, seed=1)
    ema = synth.make_unet_state_dict(None, 10, seed=2)
    ck = tmp_path / "ckpt.pth"
    torch.save({"epoch": 3, "model_state_dict": sd, "ema_model_state_dict": ema, "config": cfg}, ck)
    out = _run(ref, tmp_path, ["--checkpoint", str(ck), "--sampling_method", "ddim", "--num_inference_steps", "50",
                               "--cfg_scale", "3", "--labels", "0,1,2,9", "--num_samples", "16", "--batch_size", "8", "--use_ema",
                               "--create_gif", "--save_intermediate", "--output_dir", str(tmp_path / "o"), "--device", "cuda"])
    assert "Using EMA model" in out and "Using sampling method: DDIM" in out
    assert "labels: [1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 10, 10, 10, 10]" in out  # + 1 shift, one label per grid row
    got = _png(tmp_path / "o" / "samples.png")
    assert got.shape == (4 * 34 + 2, 4 * 34 + 2, 3)  # 4 x 4 grid of 32 x 32 images, padding 2
    assert (tmp_path / "o" / "samples.gif").stat().st_size > 10_000
    frames = sorted(os.listdir(tmp_path / "o" / "intermediate"))
    assert len(frames) == 50 and frames[0] == "step_0000.png" and frames[-1] == "step_0049.png"
    assert np.array_equal(_png(tmp_path / "o" / "intermediate" / "step_0049.png"), got)
    # the same two batches through the native API under the script's seed: set_seed(42) then batch after batch
    from diffusion_models_collection_b200.diffusion import DDIM
    from diffusion_models_collection_b200.models import UNet
    from diffusion_models_collection_b200.utils.helpers import set_seed

    set_seed(42)
    net = UNet(**cfg["model_params"], num_classes=10)
    net.load_state_dict(ema)
    net = net.cuda().eval()
    d = DDIM(cfg["num_timesteps"], 50, cfg["beta_start"], cfg["beta_end"], cfg["beta_schedule"], eta=0.0, device=torch.device("cuda"))
    d.progress = False
    labels = (torch.tensor([0, 1, 2, 9], device="cuda") + 1).repeat_interleave(4)
    imgs = [d.sample_with_cfg(net, (8, 3, 32, 32), labels[s:s + 8], cfg_scale=3.0, return_all_timesteps=True)[-1] for s in (0, 8)]
    want = _grid_png(torch.cat(imgs), 4, tmp_path / "want.png")
    assert np.array_equal(got, want)


def test_sample_py_ddpm_uncond_and_default_labels_on_b200(tmp_path):
    """--sampling_method ddpm (1000 steps, fresh noise each step) on an unconditional checkpoint, then a conditional one without
    --labels (the script draws one class per grid row, sample.py:160-163) and without CFG (cfg_scale 0 -> diffusion.sample)"""
    ref = _ref_dir()
    cfg = _config(ref, "cifar10_unet")
    cfg["conditional"], cfg["num_timesteps"] = False, 100  # a 100-step DDPM chain keeps the test short; the loop is the same
    ck = tmp_path / "u.pth"
    torch.save({"epoch": 1, "model_state_dict": synth.make_unet_state_dict(None, None, seed=5), "config": cfg}, ck)
    _run(ref, tmp_path, ["--checkpoint", str(ck), "--sampling_method", "ddpm", "--num_samples", "9", "--batch_size", "16",
                         "--output_dir", str(tmp_path / "u"), "--device", "cuda"])
    got = _png(tmp_path / "u" / "samples.png")
    assert got.shape == (3 * 34 + 2, 3 * 34 + 2, 3)
    from diffusion_models_collection_b200.diffusion import DDPM
    from diffusion_models_collection_b200.models import UNet
    from diffusion_models_collection_b200.utils.helpers import set_seed

    set_seed(42)
    net = UNet(**cfg["model_params"], num_classes=None)
    net.load_state_dict(synth.make_unet_state_dict(None, None, seed=5))
    net = net.cuda().eval()
    dp = DDPM(100, cfg["beta_start"], cfg["beta_end"], cfg["beta_schedule"], device=torch.device("cuda"))
    dp.progress = False
    want = _grid_png(dp.sample(net, (9, 3, 32, 32)), 3, tmp_path / "want.png")
    assert np.array_equal(got, want)
    cfg2 = _config(ref, "cifar10_unet")
    cfg2["conditional"], cfg2["num_classes"] = True, 10
    ck2 = tmp_path / "c.pth"
    torch.save({"epoch": 1, "model_state_dict": synth.make_unet_state_dict(None, 10, seed=6), "config": cfg2}, ck2)
    out = _run(ref, tmp_path, ["--checkpoint", str(ck2), "--sampling_method", "ddim", "--num_inference_steps", "10",
                               "--num_samples", "4", "--output_dir", str(tmp_path / "c"), "--device", "cuda"])
    assert "Using conditional generation with labels" in out
    assert _png(tmp_path / "c" / "samples.png").shape == (2 * 34 + 2, 2 * 34 + 2, 3)


def test_sample_py_dit_config_on_b200(tmp_path):
    ref = _ref_dir()
    cfg = _config(ref, "cifar10_dit")
    ncls = cfg.get("num_classes") if cfg.get("conditional") else None
    h, w = (cfg["image_size"], cfg["image_size"]) if isinstance(cfg["image_size"], int) else cfg["image_size"]
    dcfg = dict(cfg["model_params"])
    ck = tmp_path / "d.pth"
    torch.save({"epoch": 1, "model_state_dict": synth.make_dit_state_dict(dcfg, ncls, seed=8), "config": cfg}, ck)
    _run(ref, tmp_path, ["--checkpoint", str(ck), "--sampling_method", "ddim", "--num_inference_steps", "20", "--num_samples", "4",
                         "--output_dir", str(tmp_path / "d"), "--device", "cuda"])
    got = _png(tmp_path / "d" / "samples.png")
    assert got.shape == (2 * (h + 2) + 2, 2 * (w + 2) + 2, 3)
