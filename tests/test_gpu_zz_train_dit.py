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
    net = net.cuda().train()  # dropout 0.1 (configs/cifar10_dit.py): the masks come from the native kernels
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
    net.dropout = 0.0
    b = net(x, t)  # training forward, same weights, dropout off
    err = float((a - b.detach()).norm() / a.norm())
    assert err < 2e-2, err
    net.dropout = 0.1
    torch.manual_seed(5)
    c1 = net(x, t).detach().clone()
    torch.manual_seed(5)
    c2 = net(x, t).detach().clone()
    c3 = net(x, t).detach().clone()
    assert torch.equal(c1, c2) and not torch.equal(c1, c3)  # masks follow torch's seed
    assert 1e-3 < float((c1 - a).norm() / a.norm()) < 1.0   # and they do something


@pytest.mark.parametrize("B,L,C_,with_y,drop", [(3, 256, 384, True, 0.0), (2, 256, 384, False, 0.0), (5, 64, 512, True, 0.0),
                                                 (1, 256, 1024, True, 0.0), (4, 16, 128, False, 0.0), (3, 256, 384, True, 0.25)])
def test_gated_residual_layernorm_modulate_kernel_and_its_backward(B, L, C_, with_y, drop):
    """dmc_dit_gate_ln_mod / _backward against PyTorch autograd in fp32 (models/dit.py:117-121): outputs, the stream gradient,
    the branch gradient and the three per-image sums; a second launch reproduces the first bit for bit (fixed summation order)"""
    import ctypes as C

    from diffusion_models_collection_b200 import _lib

    lib = _lib.load()
    g = torch.Generator(device="cuda").manual_seed(B * 1000 + L + C_)
    x_in = torch.randn(B, L, C_, device="cuda", generator=g) * 1.5 + 0.3
    y = (torch.randn(B, L, C_, device="cuda", generator=g)).to(torch.bfloat16) if with_y else None
    mod = torch.randn(B, 6 * C_, device="cuda", generator=g) * 0.5
    gtab = torch.randn(B, 2 * C_, device="cuda", generator=g) * 0.5  # the gate comes from another table (other row stride)
    shift, scale, gate = mod[:, :C_], mod[:, C_:2 * C_], gtab[:, C_:]
    dh = torch.randn(B, L, C_, device="cuda", generator=g).to(torch.bfloat16)
    dx_out = torch.randn(B, L, C_, device="cuda", generator=g) if with_y else None
    # native forward
    h = torch.full((B, L, C_), float("nan"), device="cuda", dtype=torch.bfloat16)
    x_out = torch.full((B, L, C_), float("nan"), device="cuda") if with_y else None
    d = _lib.DitGlmDesc()
    d.x_in, d.h, d.shift, d.scale, d.mod_stride = x_in.data_ptr(), h.data_ptr(), shift.data_ptr(), scale.data_ptr(), mod.stride(0)
    d.B, d.L, d.C, d.eps = B, L, C_, 1e-6
    if with_y:
        d.y, d.gate, d.gate_stride, d.x_out = y.data_ptr(), gate.data_ptr(), gtab.stride(0), x_out.data_ptr()
        d.drop_p, d.seed = drop, 12345
    _lib.check(lib.dmc_dit_gate_ln_mod(C.byref(d), _lib.stream_ptr()), "glm")
    keep = None
    if drop > 0:  # the mask the kernel drew, read back from its output: y_eff = (x_out - x_in) / gate is 0 or y / (1 - p)
        y_eff = (x_out - x_in) / gate[:, None]
        keep = (y_eff.abs() > 0.5 * y.float().abs()).float()
        sure = y.float().abs() > 1e-2
        frac = float(keep[sure].mean())
        assert abs(frac - (1 - drop)) < 5e-3, frac
    # reference
    xi = x_in.clone().requires_grad_(True)
    yf = y.float().requires_grad_(True) if with_y else None
    sh, sc, gt = (t.clone().requires_grad_(True) for t in (shift, scale, gate))
    y_used = yf * keep / (1 - drop) if keep is not None else yf
    xo = xi + gt[:, None] * y_used if with_y else xi
    hr = torch.nn.functional.layer_norm(xo, (C_,), eps=1e-6) * (1 + sc[:, None]) + sh[:, None]
    loss = (hr * dh.float()).sum() + ((xo * dx_out).sum() if with_y else 0.0)
    loss.backward()
    if with_y:
        assert float((x_out - xo.detach()).abs().max()) < 2e-5
    assert float((h.float() - hr.detach()).norm() / hr.detach().norm()) < 4e-3

    def backward(scratch=None):
        dx_in = torch.full((B, L, C_), float("nan"), device="cuda")
        dy = torch.full((B, L, C_), float("nan"), device="cuda", dtype=torch.bfloat16) if with_y else None
        sums = torch.full((3, B, C_), float("nan"), device="cuda")
        b = _lib.DitGlmBwdDesc()
        b.x, b.dh, b.scale, b.mod_stride = (x_out if with_y else x_in).data_ptr(), dh.data_ptr(), scale.data_ptr(), mod.stride(0)
        b.dx_in, b.dshift, b.dscale = dx_in.data_ptr(), sums[1].data_ptr(), sums[2].data_ptr()
        b.B, b.L, b.C, b.eps = B, L, C_, 1e-6
        if with_y:
            b.dx_out, b.y, b.gate, b.gate_stride = dx_out.data_ptr(), y.data_ptr(), gate.data_ptr(), gtab.stride(0)
            b.dy, b.dgate = dy.data_ptr(), sums[0].data_ptr()
            b.drop_p, b.seed = drop, 12345
        if scratch is not None:
            b.scratch = scratch.data_ptr()
        _lib.check(lib.dmc_dit_gate_ln_mod_backward(C.byref(b), _lib.stream_ptr()), "glm backward")
        torch.cuda.synchronize()
        return dx_in, dy, sums

    dx_in, dy, sums = backward()

    def rel(a, b):
        return float((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30))

    assert rel(dx_in, xi.grad) < 1e-4
    assert rel(sums[1], sh.grad) < 1e-4 and rel(sums[2], sc.grad) < 1e-4
    if with_y:
        assert rel(dy.float(), yf.grad) < 4e-3 and rel(sums[0], gt.grad) < 1e-4
        if keep is not None:
            assert float(dy.float()[keep == 0].abs().max()) == 0.0  # dropped elements get no gradient, exactly
    dx2, dy2, sums2 = backward()
    assert torch.equal(dx_in, dx2) and torch.equal(sums[1:], sums2[1:]) and (not with_y or (torch.equal(dy, dy2) and torch.equal(sums, sums2)))
    # row slices over several CTAs per image (what the engine uses): same gradients, sums equal up to the summation order
    scratch = torch.full((B * _lib.DIT_GLM_BWD_SLICES * 3 * C_,), float("nan"), device="cuda")
    dx3, dy3, sums3 = backward(scratch)
    assert torch.equal(dx_in, dx3) and (not with_y or torch.equal(dy, dy3))
    assert rel(sums3[1], sh.grad) < 1e-4 and rel(sums3[2], sc.grad) < 1e-4 and (not with_y or rel(sums3[0], gt.grad) < 1e-4)
    dx4, dy4, sums4 = backward(scratch)
    assert torch.equal(sums3[1:], sums4[1:])


def test_gelu_kernels_match_torch():
    import ctypes as C  # noqa: F401

    from diffusion_models_collection_b200 import _lib

    lib = _lib.load()
    u = (torch.randn(3, 256, 1536, device="cuda") * 2).to(torch.bfloat16)
    dm = torch.randn(3, 256, 1536, device="cuda").to(torch.bfloat16)
    m, du = torch.empty_like(u), torch.empty_like(u)
    _lib.check(lib.dmc_gelu_forward(u.data_ptr(), m.data_ptr(), u.numel(), 0.0, 0, _lib.stream_ptr()), "gelu")
    _lib.check(lib.dmc_gelu_backward(u.data_ptr(), dm.data_ptr(), du.data_ptr(), u.numel(), 0.0, 0, _lib.stream_ptr()), "gelu backward")
    uf = u.float().requires_grad_(True)
    ref = torch.nn.functional.gelu(uf)
    ref.backward(dm.float())
    assert torch.equal(m, ref.detach().to(torch.bfloat16))
    assert float((du.float() - uf.grad).norm() / uf.grad.norm()) < 3e-3
    # + nn.Dropout(0.1): kept elements scaled by 1 / 0.9, the backward kernel regenerates the same mask
    md, dud = torch.empty_like(u), torch.empty_like(u)
    _lib.check(lib.dmc_gelu_forward(u.data_ptr(), md.data_ptr(), u.numel(), 0.1, 777, _lib.stream_ptr()), "gelu drop")
    _lib.check(lib.dmc_gelu_backward(u.data_ptr(), dm.data_ptr(), dud.data_ptr(), u.numel(), 0.1, 777, _lib.stream_ptr()), "gelu drop bwd")
    sure = ref.detach().abs() > 1e-2
    keep = md.float().abs() > 0
    assert abs(float(keep[sure].float().mean()) - 0.9) < 3e-3
    want = (ref.detach() / 0.9).to(torch.bfloat16)
    assert torch.equal(md[keep], want[keep]) and float(md[~keep].abs().max()) == 0.0
    assert float(dud.float()[sure & ~keep].abs().max()) == 0.0
    assert float((dud.float()[keep] - uf.grad[keep] / 0.9).norm() / uf.grad[keep].norm()) < 4e-3
    m3 = torch.empty_like(u)
    _lib.check(lib.dmc_gelu_forward(u.data_ptr(), m3.data_ptr(), u.numel(), 0.1, 778, _lib.stream_ptr()), "gelu drop")
    assert not torch.equal(m3, md)  # another seed, another mask


def test_dit_training_native_glue_matches_the_torch_glue(monkeypatch):
    """the fused glue kernels against the differentiable-PyTorch form of the same step (DMC_DIT_TRAIN_GLUE=torch)"""
    out = {}
    x = torch.randn(4, 3, 32, 32, device="cuda")
    t = torch.randint(0, 1000, (4,), device="cuda")
    y = torch.tensor([0, 3, 10, 7], device="cuda")
    for mode in ("native", "torch"):
        monkeypatch.setenv("DMC_DIT_TRAIN_GLUE", mode)
        net = DiT(**synth.CIFAR_DIT, num_classes=10)
        net.load_state_dict(synth.make_dit_state_dict(synth.CIFAR_DIT, 10, seed=5), strict=True)
        net = net.cuda().eval()
        eps = net(x, t, y)
        (eps - x).pow(2).mean().backward()
        out[mode] = (eps.detach(), {n: p.grad.clone() for n, p in net.named_parameters()})
    assert float((out["native"][0] - out["torch"][0]).norm() / out["torch"][0].norm()) < 5e-3
    tot = sum(float(g.double().norm()) ** 2 for g in out["torch"][1].values()) ** 0.5
    for n, g in out["torch"][1].items():
        if float(g.norm()) > 1e-2 * tot:
            err = float((out["native"][1][n] - g).norm() / g.norm())
            assert err < 3e-2, (n, err)
