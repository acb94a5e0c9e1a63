"""CPU: pins the oracle (oracle/*.py) against fixtures written from the LIVE reference
(tests/golden/make_golden.py).  fp32 CPU vs fp32 CPU of the same torch build => tight tolerances;
integer work (timesteps, label clamp / null token) is bit-exact."""

import os

import numpy as np
import pytest
import torch

from diffusion_models_collection_b200 import synth
from oracle import model_oracle, sched_oracle as so
from tests.golden_cases import DIT_CASES, SMALL_UNET, UNET_CASES, case_inputs


def rel_l2(a, b):
    a = torch.as_tensor(a).double()
    b = torch.as_tensor(b).double()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


@pytest.mark.parametrize("name", list(UNET_CASES))
def test_unet_oracle_matches_reference(golden, name):
    c = UNET_CASES[name]
    cfg = SMALL_UNET if c.get("small") else synth.CIFAR_UNET
    sd = synth.make_unet_state_dict(cfg, c["num_classes"], seed=c["wseed"], null_row_zero=c.get("null_row_zero", True))
    x, t, y = case_inputs(c)
    eps = model_oracle.unet_forward(sd, cfg, x, t, y, num_classes=c["num_classes"])
    assert rel_l2(eps, golden["unet"][name]) < 2e-6


def test_unet_null_label_is_bit_identical_to_none():
    # padding row 0 is all-zero => y=0 == y=None bit for bit
    cfg = SMALL_UNET
    sd = synth.make_unet_state_dict(cfg, 10, seed=4)
    x, t, _ = case_inputs(UNET_CASES["small_cond"])
    a = model_oracle.unet_forward(sd, cfg, x, t, torch.zeros(3, dtype=torch.long), num_classes=10)
    b = model_oracle.unet_forward(sd, cfg, x, t, None, num_classes=10)
    assert torch.equal(a, b)


@pytest.mark.parametrize("name", list(DIT_CASES))
def test_dit_oracle_matches_reference(golden, name):
    c = DIT_CASES[name]
    cfg = synth.CIFAR_DIT
    sd = synth.make_dit_state_dict(cfg, c["num_classes"], seed=c["wseed"])
    x, t, y = case_inputs(c)
    eps = model_oracle.dit_forward(sd, cfg, x, t, y, num_classes=c["num_classes"])
    assert rel_l2(eps, golden["dit"][name]) < 5e-6


def test_param_counts():
    # SURVEY.md section 4 golden values
    n = lambda sd: sum(v.numel() for v in sd.values())  # noqa: E731
    assert n(synth.make_unet_state_dict(None, None)) == 37_064_707
    sd = synth.make_unet_state_dict(None, 10)
    assert n(sd) == 39_626_243 and len(sd) == 357
    assert len(synth.make_unet_state_dict(None, None)) == 334
    assert n(synth.make_dit_state_dict(None, None)) == 32_569_740
    assert n(synth.make_dit_state_dict(None, 10)) == 32_573_964


@pytest.mark.parametrize("sched,T", [("linear", 1000), ("cosine", 1000), ("quadratic", 1000), ("linear20", 20)])
def test_tables_bit_exact(golden, sched, T):
    tb = so.make_tables(T, 1e-4, 0.02, "linear" if sched == "linear20" else sched)
    g = golden["tables"]
    for k, v in tb.items():
        assert np.array_equal(v.numpy(), g[f"{sched}.{k}"]), k


def test_table_known_answers(golden):
    tb = so.make_tables()
    assert float(tb["alphas_cumprod"][0]) == 0.9998999834060669
    assert float(tb["alphas_cumprod"][999]) == pytest.approx(4.035830352222547e-05, rel=1e-6)
    assert float(tb["posterior_variance"][0]) == 0.0
    assert float(tb["posterior_log_variance_clipped"][0]) == pytest.approx(-46.0517, abs=1e-3)
    assert float(tb["posterior_mean_coef2"][0]) == 0.0


def test_ddim_timesteps(golden):
    g = golden["tables"]
    for T, S in ((1000, 50), (1000, 10), (1000, 1000), (1000, 7), (20, 7), (1000, 1)):
        assert np.array_equal(so.ddim_timesteps(T, S).numpy(), g[f"timesteps.{T}.{S}"])
    assert np.array_equal(so.ddim_timesteps(1000, 25).numpy(), g["timesteps.set25"])
    assert so.ddim_timesteps(1000, 50).tolist()[:6] == [999, 979, 958, 938, 917, 897]


def test_quantile_restatement_matches_torch():
    g = torch.Generator().manual_seed(3)
    x = torch.randn(5, 3072, generator=g).abs()
    for q in (0.995, 0.9, 0.5, 0.123):
        lo, hi, w = so.quantile_rank(3072, q)
        srt, _ = torch.sort(x, dim=1)
        mine = so._lerp(srt[:, lo], srt[:, hi], w)
        assert torch.equal(mine, torch.quantile(x, q, dim=1)), q
    assert so.quantile_rank(3072, 0.995) == (3055, 3056, 0.64501953125)


def test_single_steps(golden):
    g = golden["steps"]
    x, eps = torch.from_numpy(g["x"]), torch.from_numpy(g["eps"])
    B = x.shape[0]
    full = lambda v: torch.full((B,), v, dtype=torch.long)  # noqa: E731
    tb = so.make_tables()
    for eta in (0.0, 0.5):
        for (t, tn) in ((999, 979), (510, 489), (20, 0), (0, -1)):
            for clip in (True, False):
                key = f"ddim.eta{eta}.t{t}.n{tn}.clip{int(clip)}"
                noise = torch.from_numpy(g[key + ".noise"]) if eta > 0 else None
                r = so.ddim_step(tb, x, eps, full(t), full(tn), eta=eta, clip_denoised=clip, noise=noise)
                assert np.array_equal(r.numpy(), g[key]), key
    for t in (999, 500, 1, 0):
        for clip in (True, False):
            key = f"ddpm.t{t}.clip{int(clip)}"
            r = so.ddpm_step(tb, x, eps, full(t), torch.from_numpy(g[key + ".noise"]), clip_denoised=clip)
            assert np.array_equal(r.numpy(), g[key]), key
    tq = torch.from_numpy(g["q_sample.t"])
    assert np.array_equal(so.q_sample(tb, x, tq, eps).numpy(), g["q_sample.ddpm"])
    assert np.array_equal(g["q_sample.ddim"], g["q_sample.ddpm"])


def test_sampler_loops(golden):
    g = golden["loops"]
    y = torch.tensor([3, 10])
    tb = so.make_tables()
    ts50, ts10 = so.ddim_timesteps(1000, 50), so.ddim_timesteps(1000, 10)

    def n(key):
        return torch.from_numpy(g[key + ".noise"])

    def eq(key, val):
        assert np.array_equal(val.numpy(), g[key]), key

    eq("ddim50.sample", so.ddim_sample(so.toy_model, tb, ts50, n("ddim50.sample")[0]))
    eq("ddim50.sample_y", so.ddim_sample(so.toy_model, tb, ts50, n("ddim50.sample_y")[0], y))
    eq("ddim50.cfg3", so.ddim_sample_cfg(so.toy_model, tb, ts50, n("ddim50.cfg3")[0], y, 3.0))
    eq("ddim10.sample.traj", so.ddim_sample(so.toy_model, tb, ts10, n("ddim10.sample.traj")[0], trajectory=True))
    eq("ddim10.cfg3.traj", so.ddim_sample_cfg(so.toy_model, tb, ts10, n("ddim10.cfg3.traj")[0], y, 3.0, trajectory=True))
    eq("ddim10.cfg1p5.nothr", so.ddim_sample_cfg(so.toy_model, tb, ts10, n("ddim10.cfg1p5.nothr")[0], y, 1.5, None))
    eq("ddim10.cfg3.p90", so.ddim_sample_cfg(so.toy_model, tb, ts10, n("ddim10.cfg3.p90")[0], y, 3.0, 0.9))
    nz = n("ddim10.eta0p3.sample")
    eq("ddim10.eta0p3.sample", so.ddim_sample(so.toy_model, tb, ts10, nz[0], y, eta=0.3, step_noise=nz[1:]))
    nz = n("ddim10.eta0p3.cfg3")
    eq("ddim10.eta0p3.cfg3", so.ddim_sample_cfg(so.toy_model, tb, ts10, nz[0], y, 3.0, eta=0.3, step_noise=nz[1:]))
    tb20 = so.make_tables(20)
    nz = n("ddpm20.sample.traj")
    eq("ddpm20.sample.traj", so.ddpm_sample(so.toy_model, tb20, nz[0], nz[1:], y, trajectory=True))
    nz = n("ddpm20.cfg3.traj")
    eq("ddpm20.cfg3.traj", so.ddpm_sample_cfg(so.toy_model, tb20, nz[0], nz[1:], y, 3.0, trajectory=True))
    nz = n("ddpm20.cfg2.nothr")
    eq("ddpm20.cfg2.nothr", so.ddpm_sample_cfg(so.toy_model, tb20, nz[0], nz[1:], y, 2.0, None))


def test_oracle_whole_ddim50_run_matches_reference_samples(golden):
    """the oracle's UNet restatement + DDIM loop from the recorded x_T reproduce the FINAL images of the reference's own
    DDIM.sample over 50 steps (fp32 CPU both; the per-forward 2e-6 differences of the restated ops do not grow)"""
    g = golden["samples"]
    sd = synth.make_unet_state_dict(None, None, seed=42)
    tb = so.make_tables()
    ts = so.ddim_timesteps(1000, 50)

    def model(x, t, y=None):
        return model_oracle.unet_forward(sd, synth.CIFAR_UNET, x, t, y, num_classes=None)

    img = so.ddim_sample(model, tb, ts, torch.from_numpy(g["unet.uncond.ddim50.xT"]))
    ref = torch.from_numpy(g["unet.uncond.ddim50"])
    assert float((img - ref).abs().max()) < 1e-3


@pytest.mark.parametrize("name", ["cond_b3", "uncond_b2"])
def test_oracle_training_gradients_match_the_reference(golden, name):
    """BASELINE configs[4]: loss and parameter gradients of autograd through the oracle restatement == the reference's own
    DDPM.p_losses(UNet, ...).backward() (train_golden.npz: per-tensor norm and sum for every tensor, every 1-D tensor in full, 256
    fixed entries of each larger one).  Pins the oracle's BACKWARD too -- e.g. the null row of label_embed (padding_idx=0,
    models/unet.py:183) receives no gradient."""
    from tests.golden_cases import TRAIN_CASES, perturbed_state_dict, sample_index, train_inputs

    c, g = TRAIN_CASES[name], golden["train"]
    sd = {k: v.clone().requires_grad_(True) for k, v in perturbed_state_dict(c["num_classes"]).items()}
    x0, t, y, noise = train_inputs(c)
    tb = so.make_tables()
    eps = model_oracle.unet_forward.__wrapped__(sd, synth.CIFAR_UNET, so.q_sample(tb, x0, t, noise), t, y, c["num_classes"])
    loss = torch.nn.functional.mse_loss(noise, eps)
    assert abs(loss.item() - float(g[name + "/loss"])) < 2e-6
    names = [str(n) for n in g[name + "/names"]]
    assert names == list(sd)  # same tensors, same order as the reference's named_parameters()
    grads = torch.autograd.grad(loss, [sd[n] for n in names])
    norms, sums = g[name + "/norms"], g[name + "/sums"]
    total = float(np.sqrt((norms ** 2).sum()))
    for i, (n, gr) in enumerate(zip(names, grads)):
        assert abs(float(gr.double().norm()) - norms[i]) <= 2e-5 * norms[i] + 1e-7 * total, n
        assert abs(float(gr.double().sum()) - sums[i]) <= 1e-4 * norms[i] * gr.numel() ** 0.5 + 1e-9, n
        flat = gr.reshape(-1)
        if gr.dim() == 1:
            assert rel_l2(flat, g[f"{name}/full/{n}"]) < 2e-5, n
        else:
            want = g[f"{name}/sample/{n}"]
            got = flat[torch.from_numpy(sample_index(flat.numel()))]
            assert float((got - torch.from_numpy(want)).abs().max()) <= 2e-5 * float(np.abs(want).max()) + 1e-6 * norms[i] / gr.numel() ** 0.5, n
    if c["num_classes"]:
        row0 = g[f"{name}/sample/label_embed.weight"]  # (sampled entries; the full check is on the oracle side)
        assert float(grads[names.index("label_embed.weight")][0].abs().max()) == 0.0 and row0 is not None


@pytest.mark.parametrize("name", ["dit_cond_b3", "dit_uncond_b2"])
def test_oracle_dit_training_gradients_match_the_reference(golden, name):
    """the DiT training fixture (train_dit_golden.npz, the reference's own DDPM.p_losses(DiT, ...).backward(), eval mode) against
    autograd through the oracle's restatement of models/dit.py:263-295: pins the oracle's DiT backward as well"""
    from tests.golden_cases import TRAIN_DIT_CASES, sample_index, train_inputs

    c, g = TRAIN_DIT_CASES[name], golden["train_dit"]
    sd = {k: v.clone().requires_grad_(True) for k, v in synth.make_dit_state_dict(synth.CIFAR_DIT, c["num_classes"], seed=c["wseed"]).items()}
    x0, t, y, noise = train_inputs(c)
    tb = so.make_tables()
    fwd = getattr(model_oracle.dit_forward, "__wrapped__", model_oracle.dit_forward)
    eps = fwd(sd, synth.CIFAR_DIT, so.q_sample(tb, x0, t, noise), t, y, c["num_classes"])
    loss = torch.nn.functional.mse_loss(noise, eps)
    assert abs(loss.item() - float(g[name + "/loss"])) < 5e-6
    names = [str(n) for n in g[name + "/names"]]
    assert names == list(sd)
    grads = torch.autograd.grad(loss, [sd[n] for n in names])
    norms = g[name + "/norms"]
    total = float(np.sqrt((norms ** 2).sum()))
    for i, (n, gr) in enumerate(zip(names, grads)):
        assert abs(float(gr.double().norm()) - norms[i]) <= 5e-5 * norms[i] + 1e-7 * total, n
        flat = gr.reshape(-1)
        key = f"{name}/full/{n}"
        if key in g.files:
            assert rel_l2(flat, g[key]) < 5e-5 or norms[i] < 1e-6 * total, n
        else:
            want = g[f"{name}/sample/{n}"]
            got = flat[torch.from_numpy(sample_index(flat.numel()))]
            assert float((got - torch.from_numpy(want)).abs().max()) <= 5e-5 * float(np.abs(want).max()) + 1e-6 * norms[i] / gr.numel() ** 0.5, n
    if c["num_classes"]:
        assert float(grads[names.index("y_embedder.embedding_table.weight")][0].abs().max()) == 0.0  # padding row


def test_config1_b16_oracle_matches_reference_golden():
    """BASELINE configs[0] at its own batch size: the restatement's eps for the 16 images of the reference's first DDIM step (synth
    weights seed 42, checksums stored next to the golden)"""
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "config1_golden.npz"))
    sd = synth.make_unet_state_dict(None, None, seed=42)
    assert np.array_equal(np.array([int(v.contiguous().view(torch.int32).to(torch.int64).sum()) for v in sd.values()]),
                          g["weight_isums"])
    x = torch.from_numpy(g["xT"])
    eps = model_oracle.unet_forward(sd, synth.CIFAR_UNET, x, torch.full((16,), 999, dtype=torch.long), None, num_classes=None)
    ref_eps = torch.from_numpy(g["eps0"])
    err = float((eps - ref_eps).norm() / ref_eps.norm())
    assert err < 2e-6, err
    # and one free-running step of the oracle's sampler lands on the reference's state after step 1
    tb = so.make_tables()
    ts = so.ddim_timesteps(1000, 50)
    x1 = so.ddim_step(tb, x, eps, torch.full((16,), int(ts[0])), torch.full((16,), int(ts[1])))
    assert float((x1 - torch.from_numpy(g["after1"])).abs().max()) < 1e-5
