"""GPU: the generation loop of the reference's evaluate.py (evaluate.py:201-219) at its published shape -- DDPM ancestral
sampling with classifier-free guidance, 512 images per `diffusion.sample_with_cfg` call (docs/cifar10_runs.md:129), i.e. 1024
images per UNet launch, per-step noise drawn inside the captured step graph -- SURVEY.md section 8 row f4.  A short chain
(T = 6) keeps the test fast; the loop, its kernels and its batch geometry are those of the 1000-step run
(`python bench.py --workload eval_ddpm1000_cfg` measures that one).  The DDPM + CFG + dynamic-threshold algebra itself is pinned
bit-exactly against the oracle and the reference's golden loops in tests/test_gpu_sched.py."""

import pytest
import torch

from diffusion_models_collection_b200 import synth

pytestmark = pytest.mark.gpu


def test_evaluate_py_generation_shape_ddpm_cfg_batch_512():
    from diffusion_models_collection_b200.diffusion import DDPM
    from diffusion_models_collection_b200.models import UNet

    net = UNet(**synth.CIFAR_UNET, num_classes=10)
    net.load_state_dict(synth.make_unet_state_dict(None, 10, seed=42))
    net = net.cuda().eval()
    B, T = 512, 6
    g = torch.Generator().manual_seed(11)
    y = (torch.randint(0, 10, (B,), generator=g) + 1).cuda()          # evaluate.py:198: real labels + 1
    xT = torch.randn(B, 3, 32, 32, generator=g)
    zs = torch.randn(T, B, 3, 32, 32, generator=g)
    d = DDPM(T, 1e-4, 0.02, "linear", device=torch.device("cuda"))
    d.progress = False
    # (1) the call evaluate.py makes: graph loop, noise from the torch generator
    torch.manual_seed(5)
    a = d.sample_with_cfg(net, (B, 3, 32, 32), y, cfg_scale=3.0)
    torch.manual_seed(5)
    b = d.sample_with_cfg(net, (B, 3, 32, 32), y, cfg_scale=3.0)
    assert a.shape == (B, 3, 32, 32) and torch.isfinite(a).all() and torch.equal(a, b)
    assert d._graph_cache is not None
    denorm = (a + 1) / 2                                               # evaluate.py:216
    assert float(denorm.min()) >= -1e-6 and float(denorm.max()) <= 1 + 1e-6  # dynamic thresholding keeps x0 in [-1, 1]
    # (2) batch invariance at this shape: rows of the 512-image run == the same rows sampled alone (injected noise)
    big = d.sample_with_cfg(net, (B, 3, 32, 32), y, cfg_scale=3.0, noise=xT.cuda(), step_noise=zs.cuda())
    for lo, hi in ((0, 3), (509, 512)):
        part = d.sample_with_cfg(net, (hi - lo, 3, 32, 32), y[lo:hi], cfg_scale=3.0, noise=xT[lo:hi].cuda(),
                                 step_noise=zs[:, lo:hi].cuda())
        assert torch.equal(big[lo:hi], part)
