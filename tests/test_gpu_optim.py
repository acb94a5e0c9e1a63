"""GPU: FusedAdamW (csrc/optim.cu) against torch.nn.utils.clip_grad_norm_ + torch.optim.AdamW + the reference trainer's EMA update
(utils/trainer.py:187-202, 256-262) on the same tensors.  fp32 everywhere; tolerance 2e-6 relative (different association of the
norm's sum, fused multiply-adds)."""

import pytest
import torch

from diffusion_models_collection_b200.optim import FusedAdamW

pytestmark = pytest.mark.gpu

SHAPES = [(3,), (128,), (256, 128, 3, 3), (11, 512), (512, 512), (256, 384, 1, 1), (1,), (40000,)]


def _tensors(seed, scale=1.0):
    g = torch.Generator(device="cuda").manual_seed(seed)
    return [torch.randn(s, device="cuda", generator=g) * scale for s in SHAPES]


@pytest.mark.parametrize("max_norm,ema", [(1.0, True), (0.0, False), (1e9, True)])
def test_fused_adamw_matches_torch(max_norm, ema):
    ours = [torch.nn.Parameter(t.clone()) for t in _tensors(1)]
    ref = [torch.nn.Parameter(t.clone()) for t in _tensors(1)]
    ema_o = [p.detach().clone() for p in ours] if ema else None
    ema_r = [p.detach().clone() for p in ref] if ema else None
    opt_o = FusedAdamW(ours, lr=2e-3, weight_decay=1e-2, max_grad_norm=max_norm or None, ema_params=ema_o, ema_decay=0.99)
    opt_r = torch.optim.AdamW(ref, lr=2e-3, weight_decay=1e-2)
    for step in range(4):
        grads = _tensors(10 + step, scale=3.0 if step % 2 else 0.01)  # alternately far above / below the clipping threshold
        for p, q, g in zip(ours, ref, grads):
            p.grad, q.grad = g.clone(), g.clone()
        if step == 3:  # a parameter without a gradient is left alone by both (its bias correction afterwards follows the group's
            # step count here, its own in torch: a parameter that misses steps is outside the exact-parity claim)
            ours[3].grad = ref[3].grad = None
        if max_norm:
            want_norm = torch.nn.utils.clip_grad_norm_(ref, max_norm)
        opt_r.step()
        if ema:
            for e, q in zip(ema_r, ref):
                e.mul_(0.99).add_(q.detach(), alpha=1 - 0.99)
        opt_o.step()
        torch.cuda.synchronize()
        if max_norm:
            assert abs(float(opt_o.last_grad_norm) - float(want_norm)) <= 2e-6 * float(want_norm)
        for i, (p, q) in enumerate(zip(ours, ref)):
            assert torch.allclose(p, q, rtol=2e-6, atol=1e-7), (step, i, float((p - q).abs().max()))
        if ema:
            # (the reference's EMA also moves for a parameter without a gradient; ours folds the EMA into the update pass, so
            # only compare tensors that took part in every step)
            for i, (e, f) in enumerate(zip(ema_o, ema_r)):
                if i != 3:
                    assert torch.allclose(e, f, rtol=2e-6, atol=1e-7), (step, i)
    sd = opt_o.state_dict()
    assert set(sd["state"][0]) == {"step", "exp_avg", "exp_avg_sq"}


def test_fused_adamw_refuses_cpu_tensors():
    from diffusion_models_collection_b200 import _lib

    p = torch.nn.Parameter(torch.zeros(4))
    p.grad = torch.ones(4)
    with pytest.raises(_lib.DmcError):
        FusedAdamW([p]).step()
