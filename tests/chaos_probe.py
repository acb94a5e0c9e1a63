#!/usr/bin/env python
"""Why final DDIM-50 images cannot pin an implementation when the weights are random-init (CPU, oracle only):
the fp32 run with eps perturbed by a tiny relative amount per step ends far from the unperturbed run.

    python tests/chaos_probe.py        # prints final max-abs / relative L2 for eps perturbations 1e-4, 1e-3, 6e-3
Measured here: 1e-4 -> 1.74 / 0.46, 1e-3 -> 1.95 / 0.67, 6e-3 -> 1.99 / 0.91 (images clamp to [-1, 1])."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
from diffusion_models_collection_b200 import synth
from oracle import model_oracle, sched_oracle as so

g = np.load(os.path.join(ROOT, "tests", "golden", "samples_golden.npz"))
sd = synth.make_unet_state_dict(None, None, seed=42)
tb, ts = so.make_tables(), so.ddim_timesteps(1000, 50)
xT, ref = torch.from_numpy(g["unet.uncond.ddim50.xT"]), torch.from_numpy(g["unet.uncond.ddim50"])


def perturbed(rel):
    def model(x, t, y=None):
        e = model_oracle.unet_forward(sd, synth.CIFAR_UNET, x, t, y, num_classes=None)
        n = torch.randn_like(e)
        return e + rel * e.norm() / n.norm() * n
    return model


torch.manual_seed(0)
for rel in (1e-4, 1e-3, 6e-3):
    img = so.ddim_sample(perturbed(rel), tb, ts, xT)
    print(f"eps perturbation {rel:g}: final max-abs {float((img - ref).abs().max()):.3f}, "
          f"relative L2 {float((img - ref).norm() / ref.norm()):.3f}")
