"""GPU: BASELINE.json configs[0] -- the reference's own CPU-runnable case (uncond UNet, DDIM-50, batch 16, seed 42) -- replayed
on the native path against tests/golden/config1_golden.npz (written from the live reference by tests/golden/make_golden.py
config1).  Weights: synth.make_unet_state_dict(seed 42), proven identical to the golden run's through stored per-tensor
checksums (PyTorch's default init is not bit-reproducible across host CPU models, see make_golden.gen_config1).

* whole-model eps at B = 16 (the batch size of configs[0]): bf16 <= 2e-2, split-bf16 ("bf16x3") <= 1e-3 relative L2;
* FREE-RUNNING DDIM from the reference's x_T: deviation from the reference's own state after 1, 2, 3, 5, 10, 20 and 50 steps.
  With random-init weights the sampler is chaotic (tests/chaos_probe.py: the reference against itself with a 1e-5 relative eps
  perturbation ends 1.9 max-abs away), so the bound that means something is the one over the first steps, before the
  divergence saturates; it is asserted in the fp32-accuracy mode at 3x the measured values, the final-image deviation is
  recorded next to the reference's own self-divergence."""

import os

import numpy as np
import pytest
import torch

from diffusion_models_collection_b200 import synth
from tests.gpu_util import rel_l2

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

# max-abs deviation (16 x 3072 values) of the free-running state from the reference's own state after h steps: gates = 3x the
# values measured on B200 (round 2, profiles/r02_eps_errors_run3.txt):
#   split-bf16 ("bf16x3", per-step eps error 1.6e-5):  h = 1: 3.0e-5, 2: 5.9e-5, 3: 9.8e-5, 5: 3.7e-4, 10: 4.3e-3, 20: 0.37, 50: 1.93
#   bf16 (per-step eps error 6.2e-3):                  h = 1: 1.0e-2, 2: 2.2e-2, 3: 4.3e-2, 5: 0.12,   10: 1.34,   20: 2.8,  50: 2.0
# The deviation grows ~10x every 5 steps whatever the precision: beyond ~15 steps it has saturated at the clamp range, exactly
# like the reference against itself (tests/chaos_probe.py).  Horizons past the gated ones are recorded, not asserted.
FREE_RUN_GATES_X3 = {1: 1e-4, 2: 2e-4, 3: 3e-4, 5: 1.2e-3, 10: 1.3e-2}
FREE_RUN_GATES_BF16 = {1: 3e-2, 2: 7e-2, 3: 1.3e-1, 5: 3.6e-1}


def _reference_weights(g):
    sd = synth.make_unet_state_dict(None, None, seed=42)
    isums = np.array([int(v.contiguous().view(torch.int32).to(torch.int64).sum()) for v in sd.values()])
    assert list(sd) == list(g["weight_names"])
    assert np.array_equal(isums, g["weight_isums"]), \
        "synth.make_unet_state_dict(seed 42) did not reproduce the weights the golden run used"
    return sd


def _log(line):
    if os.path.isdir(os.path.join(ROOT, "gpurun_out")):
        with open(os.path.join(ROOT, "gpurun_out", "eps_errors.txt"), "a") as fh:
            fh.write(line + "\n")
    print(line)


@pytest.fixture(scope="module")
def net_and_golden():
    from diffusion_models_collection_b200.models import UNet

    g = np.load(os.path.join(ROOT, "tests", "golden", "config1_golden.npz"))
    net = UNet(**synth.CIFAR_UNET, num_classes=None)
    net.load_state_dict(_reference_weights(g), strict=True)
    return net.cuda().eval(), g


def test_config1_eps_at_batch_16(net_and_golden):
    net, g = net_and_golden
    xT = torch.from_numpy(g["xT"]).cuda()
    t = torch.full((16,), 999, device="cuda", dtype=torch.long)
    ref = torch.from_numpy(g["eps0"])
    with torch.no_grad():
        e16 = net(xT, t)
        net.precision = "bf16x3"
        try:
            e32 = net(xT, t)
        finally:
            net.precision = "bf16"
    a, b = rel_l2(e16, ref), rel_l2(e32, ref)
    _log(f"config1_b16 eps rel-L2: bf16 {a:.4e} split-bf16 {b:.4e}")
    assert a < 2e-2 and b < 1e-3


@pytest.mark.parametrize("precision", ["bf16x3", "bf16"])
def test_config1_free_running_ddim50(net_and_golden, precision):
    from diffusion_models_collection_b200.diffusion import DDIM

    net, g = net_and_golden
    d = DDIM(1000, 50, 1e-4, 0.02, "linear", eta=0.0, device=torch.device("cuda"))
    d.progress = False
    xT = torch.from_numpy(g["xT"]).cuda()
    net.precision = precision
    try:
        traj = d.sample(net, (16, 3, 32, 32), noise=xT, return_all_timesteps=True)  # [50, 16, 3, 32, 32] on the host
        final = d.sample(net, (16, 3, 32, 32), noise=xT)                              # the graph loop
    finally:
        net.precision = "bf16"
    assert torch.isfinite(traj).all() and float(traj[-1].abs().max()) <= 1.0 + 1e-6
    assert torch.equal(final.cpu(), traj[-1])
    gates = FREE_RUN_GATES_X3 if precision == "bf16x3" else FREE_RUN_GATES_BF16
    worst = {}
    for h in [int(v) for v in g["horizons"]]:
        want = torch.from_numpy(g[f"after{h}"])
        mx, l2 = float((traj[h - 1] - want).abs().max()), rel_l2(traj[h - 1], want)
        worst[h] = mx
        _log(f"config1_free_running {precision} after {h} steps: max-abs {mx:.4e} rel-L2 {l2:.4e}")
    for h, gate in gates.items():
        assert worst[h] <= gate, (precision, h, worst[h], gate)
