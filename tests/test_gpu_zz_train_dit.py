"""GPU: the DiT training step (models/dit_train.py: native GEMMs, weight gradients and attention as autograd nodes around
differentiable glue) against the reference's OWN recorded loss and gradients (tests/golden/train_dit_golden.npz, written from the
live reference by `tests/golden/make_golden.py train_dit`) -- same weights, images, timesteps, labels and noise, eval mode.

Tolerances (bf16 GEMM operands and activation gradients, fp32 weight-gradient accumulation, fp32 glue): loss 5e-3 relative; per-tensor
gradient norm 8e-2 relative and the recorded entries 1e-1 relative L2 for every tensor carrying at least 1e-2 of the total gradient
norm; total gradient norm 2e-2."""

import numpy as np
import pytest
import torch

from diffusion_models_collection_b200 import synth
from diffusion_models_collection_b200.diffusion import DDPM
from diffusion_models_collection_b200.models.dit import DiT
from tests.golden_cases import TRAIN_DIT_CASES, sample_index, train_inputs

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", list(TRAIN_DIT_CASES))
def test_native_dit_training_step_matches_the_reference_gradients(golden, name):
    c, g = TRAIN_DIT_CASES[name], golden["train_dit"]
    net = DiT(**synth.CIFAR_DIT, num_classes=c["num_classes"])
    net.load_state_dict(synth.make_dit_state_dict(synth.CIFAR_DIT, c["num_classes"], seed=c["wseed"]), strict=True)
    net = net.cuda().eval()
    x0, t, y, noise = (v.cuda() if v is not None else None for v in train_inputs(c))
    ddpm = DDPM(num_timesteps=1000, beta_start=1e-4, beta_end=0.02, beta_schedule="linear", device="cuda")
    for step in range(2):  # the second pass reuses the engine's static buffers; gradients must not accumulate into them
        net.zero_grad(set_to_none=True)
        loss = ddpm.p_losses(net, x0, t, y, noise=noise, loss_type="l2")
        loss.backward()
    want_loss = float(g[name + "/loss"])
    assert abs(loss.item() - want_loss) < 5e-3 * want_loss, (loss.item(), want_loss)
    names = [str(n) for n in g[name + "/names"]]
    assert names == [n for n, _ in net.named_parameters()]
    norms = g[name + "/norms"]
    total = float(np.sqrt((norms ** 2).sum()))
    got_total, worst = 0.0, (0.0, "")
    for i, n in enumerate(names):
        gr = net.get_parameter(n).grad
        assert gr is not None and torch.isfinite(gr).all(), n
        gn = float(gr.double().norm())
        got_total += gn * gn
        if norms[i] < 1e-2 * total:
            continue
        assert abs(gn - norms[i]) <= 8e-2 * norms[i], (n, gn, norms[i])
        flat = gr.reshape(-1).cpu()
        key = f"{name}/full/{n}"
        if key in g.files:
            want, got = torch.from_numpy(g[key]), flat
        else:
            want, got = torch.from_numpy(g[f"{name}/sample/{n}"]), flat[torch.from_numpy(sample_index(flat.numel()))]
        err = float((got.double() - want.double()).norm() / want.double().norm())
        worst = max(worst, (err, n))
    print(f"\n[DiT train fixture {name}] loss {loss.item():.5f} vs {want_loss:.5f}, total grad norm {got_total ** 0.5:.4f} vs {total:.4f}, "
          f"worst recorded-entry rel_l2 {worst[0]:.3e} ({worst[1]})")
    assert abs(got_total ** 0.5 - total) <= 2e-2 * total
    assert worst[0] < 1e-1, worst
    if c["num_classes"]:
        assert float(net.get_parameter("y_embedder.embedding_table.weight").grad[0].abs().max()) == 0.0  # padding row


def test_dit_sampling_after_a_training_step_uses_the_updated_weights():
    """optimizer step -> the inference plan re-packs its bf16 operands (parameter versions) and the engine its own"""
    net = DiT(**synth.CIFAR_DIT, num_classes=None)
    net.load_state_dict(synth.make_dit_state_dict(synth.CIFAR_DIT, None, seed=3), strict=True)
    net = net.cuda().train()
    net.dropout = 0.0
    opt = torch.optim.SGD(net.parameters(), lr=1e-2)
    x = torch.randn(4, 3, 32, 32, device="cuda")
    t = torch.randint(0, 1000, (4,), device="cuda")
    losses = []
    for _ in range(3):
        opt.zero_grad(set_to_none=True)
        loss = (net(x, t) - x).pow(2).mean()
        loss.backward()
        opt.step()
        losses.append(loss.item())
    assert losses[2] < losses[0]  # the step sees its own updates
    with torch.no_grad():
        a = net.eval()(x, t)
    net.train()
    b = net(x, t)  # training forward, same weights, dropout 0
    err = float((a - b.detach()).norm() / a.norm())
    assert err < 2e-2, err
