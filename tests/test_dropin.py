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
