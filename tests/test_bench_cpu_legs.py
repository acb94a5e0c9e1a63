"""CPU: the host-side legs of bench.py.  `--impl reference` / `cpu_baseline` run the reference itself from oracle/_ref; a
checkout that never built the archive falls back to the oracle port behind the same call signatures (kind = "port")."""

import importlib
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _bench():
    sys.path.insert(0, ROOT)
    return importlib.import_module("bench")


def test_cpu_legs_use_the_reference_itself_when_the_archive_exists():
    from oracle import ref_loader

    bench = _bench()
    bench._REF.clear()
    if not ref_loader.available():
        import pytest

        pytest.skip("no reference checkout / archive here")
    r = bench._reference()
    assert r["cpu_kind"] == "reference" and r["UNet"].__module__ == "models.unet" and r["DDIM"].__module__ == "diffusion.ddim"
    bench._REF.clear()


def test_cpu_legs_fall_back_to_the_oracle_port(monkeypatch):
    from diffusion_models_collection_b200 import synth
    from oracle import ref_loader

    bench = _bench()
    monkeypatch.setattr(ref_loader, "available", lambda: False)
    bench._REF.clear()
    try:
        r = bench._reference()
        assert r["cpu_kind"] == "port"
        net = r["UNet"](**synth.CIFAR_UNET, num_classes=10)
        net.load_state_dict(synth.make_unet_state_dict(None, 10, seed=42))
        d = r["DDIM"](1000, 1, 1e-4, 0.02, "linear", eta=0.0, device="cpu")
        with torch.no_grad():
            img = d.sample_with_cfg(net.eval(), (1, 3, 32, 32), torch.tensor([3]), cfg_scale=3.0)
        assert img.shape == (1, 3, 32, 32) and bool(torch.isfinite(img).all()) and float(img.abs().max()) <= 1.0 + 1e-6
    finally:
        bench._REF.clear()
