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
