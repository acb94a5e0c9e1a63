"""GPU: the fused scheduler kernels (csrc/sched.cu) through the C ABI vs the CPU oracle and the golden fixtures
written from the live reference.  Tolerance: BIT-EXACT (the kernels round exactly where the reference's fp32 tensor
expressions round; north_star allows 1e-6 relative)."""

import ctypes as C

import numpy as np
import pytest
import torch

from oracle import sched_oracle as so

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def lib():
    from diffusion_models_collection_b200 import _lib

    return _lib.load()


def _ddim_rows(tb, t, tn, eta):
    """[1, 5] coefficient row from the oracle's CPU tables with the reference's expressions (ddim.py:174-203)."""
    acp = tb["alphas_cumprod"]
    a = acp[t]
    an = acp[tn] if tn >= 0 else torch.ones(())
    sigma = eta * torch.sqrt(torch.clamp((1 - an) / (1 - a) * (1 - a / an), min=0.0))
    dirc = torch.sqrt(torch.clamp(1 - an - sigma ** 2, min=0.0))
    return torch.stack([torch.sqrt(1 - a), torch.sqrt(a), torch.sqrt(an), dirc, sigma]).float().reshape(1, 5)


def _ddpm_rows(tb):
    T = tb["betas"].shape[0]
    mask = (torch.arange(T) != 0).float()
    return torch.stack([torch.sqrt(1.0 / tb["alphas_cumprod"]), tb["sqrt_recipm1_alphas_cumprod"],
                        tb["posterior_mean_coef1"], tb["posterior_mean_coef2"],
                        mask * torch.exp(0.5 * tb["posterior_log_variance_clipped"])], dim=1).contiguous()


def _run_step(lib, ddpm, x, eps_c, eps_u, noise, row, cfg_scale=0.0, clip_mode=1, p=None):
    from diffusion_models_collection_b200 import _lib
    from diffusion_models_collection_b200.diffusion._common import guidance

    dev = "cuda"
    xd, ec = x.to(dev).contiguous(), eps_c.to(dev).contiguous()
    eu = None if eps_u is None else eps_u.to(dev).contiguous()
    nz = None if noise is None else noise.to(dev).contiguous()
    rowd = row.to(dev).contiguous()
    out = torch.full_like(xd, float("nan"))
    B = xd.shape[0]
    g = guidance(cfg_scale, clip_mode, xd.numel() // B, p)
    fn = lib.dmc_ddpm_step if ddpm else lib.dmc_ddim_step
    _lib.check(fn(xd.data_ptr(), ec.data_ptr(), _lib.ptr(eu), _lib.ptr(nz), out.data_ptr(), B, xd.numel() // B,
                  rowd.data_ptr(), C.byref(g), _lib.stream_ptr()), "step")
    torch.cuda.synchronize()
    return out.cpu()


def test_single_steps_bit_exact_vs_golden(lib, golden):
    g = golden["steps"]
    x, eps = torch.from_numpy(g["x"]), torch.from_numpy(g["eps"])
    tb = so.make_tables()
    for eta in (0.0, 0.5):
        for (t, tn) in ((999, 979), (510, 489), (20, 0), (0, -1)):
            for clip in (True, False):
                key = f"ddim.eta{eta}.t{t}.n{tn}.clip{int(clip)}"
                noise = torch.from_numpy(g[key + ".noise"]) if eta > 0 else None
                r = _run_step(lib, False, x, eps, None, noise, _ddim_rows(tb, t, tn, eta), clip_mode=int(clip))
                assert np.array_equal(r.numpy(), g[key]), key
    rows = _ddpm_rows(tb)
    for t in (999, 500, 1, 0):
        for clip in (True, False):
            key = f"ddpm.t{t}.clip{int(clip)}"
            r = _run_step(lib, True, x, eps, None, torch.from_numpy(g[key + ".noise"]), rows[t:t + 1], clip_mode=int(clip))
            assert np.array_equal(r.numpy(), g[key]), key


def test_q_sample_bit_exact(lib, golden):
    from diffusion_models_collection_b200 import _lib

    g = golden["steps"]
    tb = so.make_tables()
    x, eps, t = (torch.from_numpy(g[k]).cuda() for k in ("x", "eps", "q_sample.t"))
    out = torch.empty_like(x)
    sa, s1 = tb["sqrt_alphas_cumprod"].cuda(), tb["sqrt_one_minus_alphas_cumprod"].cuda()
    _lib.check(lib.dmc_q_sample(x.data_ptr(), eps.data_ptr(), t.data_ptr(), sa.data_ptr(), s1.data_ptr(), out.data_ptr(),
                                x.shape[0], x.numel() // x.shape[0], _lib.stream_ptr()), "q_sample")
    assert np.array_equal(out.cpu().numpy(), g["q_sample.ddpm"])


@pytest.mark.parametrize("B,shape", [(1, (3, 32, 32)), (5, (3, 32, 32)), (3, (3, 16, 16)), (2, (1, 7, 9)), (64, (3, 32, 32))])
@pytest.mark.parametrize("ddpm", [False, True])
def test_cfg_dynamic_threshold_bit_exact_vs_oracle(lib, B, shape, ddpm):
    """CFG combine + per-sample quantile threshold + update, ragged / tiny / odd sizes included."""
    g = torch.Generator().manual_seed(B * 131 + shape[1])
    x = torch.randn(B, *shape, generator=g)
    ec = torch.randn(B, *shape, generator=g)
    eu = torch.randn(B, *shape, generator=g)
    nz = torch.randn(B, *shape, generator=g)
    # make one sample all-small (threshold floor s = 1) and one with ties at the quantile rank
    x[0] *= 0.01
    ec[0] *= 0.01
    eu[0] *= 0.01
    tb = so.make_tables()
    n = x[0].numel()
    for t, tn, p, scale in ((979, 958, 0.995, 3.0), (20, 0, 0.9, 1.5), (0, -1, 0.5, 7.5)):
        tt, tnn = torch.full((B,), t), torch.full((B,), tn)
        eps = so.cfg_combine(ec, eu, scale)
        if ddpm:
            x0 = so.dynamic_threshold(so.ddpm_x0(tb, x, eps, tt), p)
            want = so.ddpm_step(tb, x, eps, tt, nz, clip_denoised=False, x0_pred=x0)
            got = _run_step(lib, True, x, ec, eu, nz, _ddpm_rows(tb)[t:t + 1], scale, 2, p)
        else:
            x0 = so.dynamic_threshold(so.ddim_x0(tb, x, eps, tt), p)
            want = so.ddim_step(tb, x, eps, tt, tnn, eta=0.0, clip_denoised=False, x0_pred=x0)
            got = _run_step(lib, False, x, ec, eu, None, _ddim_rows(tb, t, tn, 0.0), scale, 2, p)
        assert torch.equal(got, want), (t, p, float((got - want).abs().max()))


def test_threshold_with_ties(lib):
    """all |x0| equal -> every rank has the same value; and a two-valued input straddling the rank"""
    tb = so.make_tables()
    B, shape = 2, (3, 32, 32)
    x = torch.ones(B, *shape) * 2.5
    x[1, :, :16] = -0.25
    ec = torch.zeros(B, *shape)
    tt, tn = torch.full((B,), 500), torch.full((B,), 489)
    x0 = so.dynamic_threshold(so.ddim_x0(tb, x, ec, tt), 0.995)
    want = so.ddim_step(tb, x, ec, tt, tn, clip_denoised=False, x0_pred=x0)
    got = _run_step(lib, False, x, ec, None, None, _ddim_rows(tb, 500, 489, 0.0), 0.0, 2, 0.995)
    assert torch.equal(got, want)


def _cpu_toy(x, t, y=None):
    """the oracle's toy denoiser evaluated on the CPU (any callable is a legal `model`), so that the whole loop is
    comparable bit for bit with the CPU golden trajectory"""
    return so.toy_model(x.cpu(), t.cpu(), None if y is None else y.cpu()).to(x.device)


def _with_cpu_tables(d, T=1000):
    """inject the schedule and the per-step coefficient rows computed on the CPU with the same expressions (linspace,
    cumprod, exp, log differ in the last bit between the CPU and CUDA back ends; the reference builds its tables on its
    own device, and the goldens were written on the CPU)"""
    tb = so.make_tables(T)
    if hasattr(d, "posterior_mean_coef1"):
        d._coef_cache = _ddpm_rows(tb).cuda()
    else:
        ts = d.inference_timesteps.cpu()
        nxt = torch.cat([ts[1:], torch.full((1,), -1, dtype=ts.dtype)])
        d.alphas_cumprod = tb["alphas_cumprod"]
        d._coef_cache = d._coef_rows(ts, nxt).cuda()
    for k in ("betas", "alphas", "alphas_cumprod"):
        setattr(d, k, tb[k].cuda())
    return d


def test_device_coefficient_rows_match_cpu_rows():
    """the [S, 5] DDIM rows computed on the CUDA device from the SAME alphas_cumprod, with the reference's own torch
    expressions: equal to the CPU rows up to 1 ulp (torch's CUDA sqrt differs from the CPU's in the last bit for a few
    arguments -- the reference run on this device has the same rows as we do; tolerance 2e-7 relative)"""
    from diffusion_models_collection_b200.diffusion import DDIM

    tb = so.make_tables()
    for eta in (0.0, 0.3):
        d = DDIM(1000, 50, eta=eta, device="cuda")
        ts = d.inference_timesteps
        nxt = torch.cat([ts[1:], torch.full((1,), -1, dtype=ts.dtype, device=ts.device)])
        d.alphas_cumprod = tb["alphas_cumprod"].cuda()
        dev_rows = d._coef_rows(ts, nxt).cpu()
        d.alphas_cumprod = tb["alphas_cumprod"]
        cpu_rows = d._coef_rows(ts.cpu(), nxt.cpu())
        rel = ((dev_rows - cpu_rows).abs() / cpu_rows.abs().clamp_min(1e-30)).max()
        assert float(rel) < 2e-7, (eta, float(rel))


def test_sampler_loops_bit_exact_vs_golden(lib, golden):
    from diffusion_models_collection_b200.diffusion import DDIM, DDPM

    g = golden["loops"]
    y = torch.tensor([3, 10]).cuda()

    def n(key):
        return torch.from_numpy(g[key + ".noise"]).cuda()

    def eq(key, val):
        assert np.array_equal(val.cpu().numpy(), g[key]), (key, float(np.abs(val.cpu().numpy() - g[key]).max()))

    shape = (2, 3, 32, 32)
    d50 = _with_cpu_tables(DDIM(1000, 50, device="cuda"))
    d50.progress = False
    eq("ddim50.sample", d50.sample(_cpu_toy, shape, noise=n("ddim50.sample")[0]))
    eq("ddim50.sample_y", d50.sample(_cpu_toy, shape, y, noise=n("ddim50.sample_y")[0]))
    eq("ddim50.cfg3", d50.sample_with_cfg(_cpu_toy, shape, y, 3.0, noise=n("ddim50.cfg3")[0]))
    d10 = _with_cpu_tables(DDIM(1000, 10, device="cuda"))
    d10.progress = False
    eq("ddim10.sample.traj", d10.sample(_cpu_toy, shape, noise=n("ddim10.sample.traj")[0], return_all_timesteps=True))
    eq("ddim10.cfg3.traj", d10.sample_with_cfg(_cpu_toy, shape, y, 3.0, noise=n("ddim10.cfg3.traj")[0],
                                               return_all_timesteps=True))
    eq("ddim10.cfg1p5.nothr", d10.sample_with_cfg(_cpu_toy, shape, y, 1.5, None, noise=n("ddim10.cfg1p5.nothr")[0]))
    eq("ddim10.cfg3.p90", d10.sample_with_cfg(_cpu_toy, shape, y, 3.0, 0.9, noise=n("ddim10.cfg3.p90")[0]))
    de = _with_cpu_tables(DDIM(1000, 10, eta=0.3, device="cuda"))
    de.progress = False
    nz = n("ddim10.eta0p3.sample")
    eq("ddim10.eta0p3.sample", de.sample(_cpu_toy, shape, y, noise=nz[0], step_noise=nz[1:]))
    nz = n("ddim10.eta0p3.cfg3")
    eq("ddim10.eta0p3.cfg3", de.sample_with_cfg(_cpu_toy, shape, y, 3.0, noise=nz[0], step_noise=nz[1:]))
    p20 = _with_cpu_tables(DDPM(20, device="cuda"), 20)
    p20.progress = False
    nz = n("ddpm20.sample.traj")
    eq("ddpm20.sample.traj", p20.sample(_cpu_toy, shape, y, noise=nz[0], step_noise=nz[1:], return_all_timesteps=True))
    nz = n("ddpm20.cfg3.traj")
    eq("ddpm20.cfg3.traj", p20.sample_with_cfg(_cpu_toy, shape, y, 3.0, noise=nz[0], step_noise=nz[1:],
                                               return_all_timesteps=True))
    nz = n("ddpm20.cfg2.nothr")
    eq("ddpm20.cfg2.nothr", p20.sample_with_cfg(_cpu_toy, shape, y, 2.0, None, noise=nz[0], step_noise=nz[1:]))


def test_device_tables_close_to_cpu_tables_and_timesteps_exact(golden):
    """schedules built on the CUDA device (as the reference would there): timestep indices bit-exact, fp32 tables
    within 2 ulp-ish of the CPU goldens (cumprod order differs between back ends)"""
    from diffusion_models_collection_b200.diffusion import DDIM, DDPM

    g = golden["tables"]
    d = DDIM(1000, 50, device="cuda")
    assert np.array_equal(d.inference_timesteps.cpu().numpy(), g["timesteps.1000.50"])
    d.set_inference_steps(25)
    assert np.array_equal(d.inference_timesteps.cpu().numpy(), g["timesteps.set25"])
    p = DDPM(1000, device="cuda")
    # 1 - alphas_cumprod[0] ~ 1e-4 amplifies a 1-ulp difference of the cumprod ~1e3 times in the posterior coefficients
    for k, tol in (("betas", 1e-6), ("alphas_cumprod", 1e-6), ("sqrt_recipm1_alphas_cumprod", 1e-3),
                   ("posterior_mean_coef1", 1e-3), ("posterior_mean_coef2", 1e-3)):
        a, b = getattr(p, k).cpu().double(), torch.from_numpy(g["linear." + k]).double()
        assert float(((a - b).abs() / b.abs().clamp_min(1e-12)).max()) < tol, k


def test_api_errors():
    from diffusion_models_collection_b200.diffusion import DDIM, DDPM

    d = DDIM(1000, 5, device="cuda")
    with pytest.raises(ValueError):
        d.sample_with_cfg(_cpu_toy, (1, 3, 8, 8), None)
    with pytest.raises(ValueError):
        d.sample_with_cfg(_cpu_toy, (1, 3, 8, 8), torch.ones(1, dtype=torch.long), p_threshold=1.5)
    with pytest.raises(ValueError):
        DDPM(10, beta_schedule="nope", device="cuda")
    with pytest.raises(ValueError):
        DDPM(10, device="cuda").p_losses(_cpu_toy, torch.zeros(1, 3, 8, 8).cuda(), torch.zeros(1, dtype=torch.long).cuda(),
                                         loss_type="l3")


@pytest.mark.parametrize("kind", ["ddpm", "ddim_eta"])
def test_sharded_sampling_with_per_step_noise_reproduces_the_single_process_run(kind):
    """DDPM (and DDIM with eta > 0) draw N(0,1) at every step: a rank that samples rows [lo, hi) of a global batch draws the
    GLOBAL per-step tensor and keeps its rows, so the shards of a 3-rank run (emulated one after the other on this GPU,
    each from the same seed like separate processes) concatenate to exactly the 1-rank run."""
    from diffusion_models_collection_b200.diffusion import DDIM, DDPM
    from diffusion_models_collection_b200.sharding import sharded_sample, shard_bounds

    B, shape = 7, (7, 3, 8, 8)
    mk = (lambda: DDPM(12, device="cuda")) if kind == "ddpm" else (lambda: DDIM(1000, 6, eta=0.5, device="cuda"))

    def model(x, t, y=None):  # a per-sample denoiser stand-in on the GPU
        return torch.tanh(x * 0.7 + (t.view(-1, 1, 1, 1).float() / 1000.0))

    d = mk()
    d.progress = False
    torch.manual_seed(11)
    whole = sharded_sample(d, model, shape, rank=0, world=1)
    parts = []
    for r in range(3):
        d = mk()
        d.progress = False
        torch.manual_seed(11)
        lo, hi = shard_bounds(B, r, 3)
        x_T = torch.randn(shape, device="cuda")  # what _global_noise draws on every rank
        d._noise_shard = (B, lo, hi)
        parts.append(d.sample(model, (hi - lo,) + shape[1:], noise=x_T[lo:hi]))
    assert torch.equal(torch.cat(parts, dim=0), whole)


def test_empty_batch_samples_like_the_reference():
    """shape (0, C, H, W): the reference's loops run S no-op steps and return an empty tensor (a trajectory of S empty ones)"""
    from diffusion_models_collection_b200.diffusion import DDIM, DDPM
    from oracle.sched_oracle import toy_model

    d = DDIM(1000, 5, device=torch.device("cuda"))
    d.progress = False
    assert tuple(d.sample(toy_model, (0, 3, 32, 32)).shape) == (0, 3, 32, 32)
    assert tuple(d.sample(toy_model, (0, 3, 32, 32), return_all_timesteps=True).shape) == (5, 0, 3, 32, 32)
    y = torch.zeros(0, dtype=torch.long, device="cuda")
    assert tuple(d.sample_with_cfg(toy_model, (0, 3, 32, 32), y).shape) == (0, 3, 32, 32)
    p = DDPM(7, device=torch.device("cuda"))
    p.progress = False
    assert tuple(p.sample(toy_model, (0, 3, 32, 32)).shape) == (0, 3, 32, 32)
    assert tuple(p.sample_with_cfg(toy_model, (0, 3, 32, 32), y, return_all_timesteps=True).shape) == (7, 0, 3, 32, 32)
