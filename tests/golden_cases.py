"""Definition of the golden cases shared by tests/golden/make_golden.py (which runs the live
reference in the build container) and the tests that replay them (oracle on CPU, CUDA path on GPU).
Inputs are regenerated from seeds with the CPU generator; weights come from synth.py seeds."""

from __future__ import annotations

import numpy as np
import torch

# a second, smaller geometry: exercises 64/128/192-channel layers and a 2-level UNet
SMALL_UNET = dict(
    image_size=(32, 32), in_channels=3, model_channels=64, out_channels=3, num_res_blocks=1,
    attention_resolutions=(16,), dropout=0.1, channel_mult=(1, 2), use_attention=True,
)

UNET_CASES = {
    # name: weights seed, num_classes, batch, timesteps, labels
    "uncond_t500": dict(wseed=1, num_classes=None, xseed=10, B=2, t=[500, 500], y=None),
    "uncond_t0_t999": dict(wseed=1, num_classes=None, xseed=11, B=2, t=[0, 999], y=None),
    "cond_labels": dict(wseed=2, num_classes=10, xseed=12, B=4, t=[20, 20, 20, 20], y=[0, 1, 10, 11]),
    "cond_mixed_t": dict(wseed=2, num_classes=10, xseed=13, B=3, t=[1, 979, 306], y=[5, 0, 99]),
    "cond_y_none": dict(wseed=2, num_classes=10, xseed=14, B=2, t=[163, 163], y=None),
    "cond_row0_nonzero": dict(wseed=3, num_classes=10, xseed=15, B=2, t=[897, 897], y=[0, 3], null_row_zero=False),
    "small_cond": dict(wseed=4, num_classes=10, xseed=16, B=3, t=[999, 510, 0], y=[0, 7, 10], small=True),
    "small_uncond": dict(wseed=5, num_classes=None, xseed=17, B=2, t=[41, 41], y=None, small=True),
}

DIT_CASES = {
    "uncond": dict(wseed=6, num_classes=None, xseed=20, B=2, t=[500, 20], y=None),
    "cond": dict(wseed=7, num_classes=10, xseed=21, B=3, t=[999, 0, 347], y=[0, 10, 4]),
    "cond_y_none": dict(wseed=7, num_classes=10, xseed=22, B=2, t=[61, 61], y=None),
}


# DiM as the reference builds it without mamba_ssm (nn.MultiheadAttention, 8 heads): hidden 512 -> head dim 64 (tcgen05
# attention), hidden 256 -> head dim 32 (CUDA-core flash kernel)
DIM_CASES = {
    "cond_h512": dict(wseed=8, num_classes=10, xseed=23, B=3, t=[999, 0, 347], y=[0, 10, 4], hidden=512, depth=4),
    "uncond_h256": dict(wseed=9, num_classes=None, xseed=24, B=2, t=[500, 20], y=None, hidden=256, depth=2),
}


def case_inputs(c):
    g = torch.Generator(device="cpu")
    g.manual_seed(int(c["xseed"]))
    x = torch.randn(c["B"], 3, 32, 32, generator=g)
    t = torch.tensor(c["t"], dtype=torch.long)
    y = None if c["y"] is None else torch.tensor(c["y"], dtype=torch.long)
    return x, t, y


# ---- training step (BASELINE configs[4]): tests/golden/train_golden.npz -------------------------------------------
TRAIN_CASES = {"cond_b3": dict(num_classes=10, B=3, seed=7), "uncond_b2": dict(num_classes=None, B=2, seed=8)}


# DiT training step (reference models/dit.py under DDPM.p_losses + backward, eval mode): CIFAR DiT, all zero-init tensors
# re-randomised (synth.make_dit_state_dict)
TRAIN_DIT_CASES = {"dit_cond_b3": dict(num_classes=10, B=3, seed=17, wseed=11), "dit_uncond_b2": dict(num_classes=None, B=2, seed=18, wseed=12)}


def train_inputs(c):
    """inputs of one training iteration (utils/trainer.py:221-251): images in [-1, 1], labels already shifted (0 = null), t, noise"""
    g = torch.Generator().manual_seed(c["seed"])
    B = c["B"]
    x0 = torch.rand(B, 3, 32, 32, generator=g) * 2 - 1
    t = torch.randint(0, 1000, (B,), generator=g)
    y = torch.randint(0, 11, (B,), generator=g) if c["num_classes"] else None
    if y is not None:
        y[0] = 0  # one null label: the padding row of label_embed must get no gradient
    noise = torch.randn(B, 3, 32, 32, generator=g)
    return x0, t, y, noise


def sample_index(numel, k=256):
    """fixed pseudo-random entries of a flattened tensor (a multiplicative walk: reproducible without a generator)"""
    return (np.arange(k, dtype=np.int64) * 2654435761 + 12345) % numel


def perturbed_state_dict(num_classes):
    from diffusion_models_collection_b200 import synth

    sd = synth.make_unet_state_dict(None, num_classes, seed=42)
    g = torch.Generator().manual_seed(99)
    for k, v in sd.items():  # GroupNorm affines / biases off their trivial initial values
        if v.dim() == 1:
            v.add_(0.05 * torch.randn(v.shape, generator=g))
    return sd
