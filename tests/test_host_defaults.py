"""CPU: host-side defaults and opt-in switches added in the second session of round 2 (no kernels run)."""

import torch

from diffusion_models_collection_b200 import synth
from diffusion_models_collection_b200.models.dit import DiT
from diffusion_models_collection_b200.models.unet import UNet


def test_unet_launch_size_default_and_chunking():
    assert UNet.max_images_per_launch == 4096  # DMC_MAX_IMAGES_PER_LAUNCH overrides at import time
    net = UNet(**synth.CIFAR_UNET, num_classes=10)
    # CFG doubles the images of a launch: 4096 samples with CFG = 8192 images = 4 launches of 2048 samples
    cb = max(1, min(4096, net.max_images_per_launch // 2))
    assert cb == 2048


def test_dit_qkv_rows_are_padded_only_on_request(monkeypatch):
    cfg = dict(synth.CIFAR_DIT, depth=1)
    net = DiT(**cfg, num_classes=None)
    net.load_state_dict(synth.make_dit_state_dict(cfg, None, seed=1))
    pk = net._ensure_packed(torch.device("cpu"))
    assert pk["wshape"]["blocks.0.qkv"] == (3 * 384, 384)
    monkeypatch.setenv("DMC_DIT_PAD_QKV", "1")
    net._packed = None
    pk = net._ensure_packed(torch.device("cpu"))
    assert pk["wshape"]["blocks.0.qkv"] == (1280, 384)  # zero rows up to the next multiple of 256
    w = pk["wlog"]["blocks.0.qkv"]
    assert float(w[1152:].abs().max()) == 0.0 and torch.equal(w[:1152], pk["sd"]["blocks.0.attn.in_proj_weight"])


def test_ddp_ignore_list_is_invisible_outside_a_ddp_constructor():
    net = UNet(**synth.CIFAR_UNET, num_classes=None)
    assert not hasattr(net, "_ddp_params_and_buffers_to_ignore")
    assert net._ddp_sentinel is None and net._grad_allreduce is None
    import copy

    twin = copy.deepcopy(net)
    assert twin._ddp_sentinel is None
