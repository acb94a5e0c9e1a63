"""DDPM drop-in (same constructor, attributes and methods as /root/reference/diffusion/ddpm.py:15-332).

The posterior step (x0 prediction, clamp / dynamic threshold, posterior mean, noise add, CFG combine;
ddpm.py:151-220,289-324) is ONE fused CUDA kernel (csrc/sched.cu via dmc_ddpm_step) fed by a [T, 5] device
coefficient table built once with the reference's own fp32 expressions."""

from __future__ import annotations

import contextlib

import torch
import torch.nn.functional as F

from .. import _lib
from ._common import DiffusionBase, guidance


class DDPM(DiffusionBase):
    def __init__(self, num_timesteps=1000, beta_start=0.0001, beta_end=0.02, beta_schedule="linear", device="cuda"):
        self._init_common(num_timesteps, beta_start, beta_end, beta_schedule, device)
        # ddpm.py:52-71
        self.alphas_cumprod_prev = F.pad(self.alphas_cumprod[:-1], (1, 0), value=1.0)
        self.sqrt_recip_alphas = torch.sqrt(1.0 / self.alphas)
        self.sqrt_recipm1_alphas_cumprod = torch.sqrt(1.0 / self.alphas_cumprod - 1)
        self.posterior_variance = self.betas * (1.0 - self.alphas_cumprod_prev) / (1.0 - self.alphas_cumprod)
        self.posterior_log_variance_clipped = torch.log(torch.clamp(self.posterior_variance, min=1e-20))
        self.posterior_mean_coef1 = self.betas * torch.sqrt(self.alphas_cumprod_prev) / (1.0 - self.alphas_cumprod)
        self.posterior_mean_coef2 = (1.0 - self.alphas_cumprod_prev) * torch.sqrt(self.alphas) / (1.0 - self.alphas_cumprod)
        self._coef_cache = None

    def _coef_table(self):
        """[T, 5] fp32 rows (dmc_ddpm_coef), ddpm.py:170-175,184-185,218-220 vectorised over t."""
        if self._coef_cache is None:
            t = torch.arange(self.num_timesteps, device=self.betas.device)
            mask = (t != 0).float()
            self._coef_cache = torch.stack([
                torch.sqrt(1.0 / self.alphas_cumprod),
                self.sqrt_recipm1_alphas_cumprod,
                self.posterior_mean_coef1,
                self.posterior_mean_coef2,
                mask * torch.exp(0.5 * self.posterior_log_variance_clipped),
            ], dim=1).contiguous()
        return self._coef_cache

    def _seq_tables(self, device):
        """(timesteps int64 [T] = T-1..0, coefficient rows fp32 [T, 5] in that order) on `device`"""
        c = getattr(self, "_seq_cache", None)
        coefs = self._coef_table()
        if c is None or c[0] is not coefs or c[1] != str(device):
            t_seq = torch.arange(self.num_timesteps - 1, -1, -1, device=device, dtype=torch.long)
            c = self._seq_cache = (coefs, str(device), t_seq, coefs.to(device)[t_seq].contiguous())
        return c[2], c[3]

    def _step(self, lib, x, eps_c, eps_u, noise, out, coef_row_ptr, g):
        B = x.shape[0]
        _lib.check(lib.dmc_ddpm_step(x.data_ptr(), eps_c.data_ptr(), _lib.ptr(eps_u), noise.data_ptr(), out.data_ptr(), B,
                                     x.numel() // B, coef_row_ptr, g, _lib.stream_ptr()), "dmc_ddpm_step")

    def p_mean_variance(self, model, x, t, y=None, clip_denoised=True, eps=None, x0_pred=None):
        """ddpm.py:151-195 (kept as plain tensor algebra: it is not on the sampling hot path, p_sample is)."""
        if eps is None:
            eps = model(x, t, y)
        if x0_pred is None:
            x0_pred = (self._extract(torch.sqrt(1.0 / self.alphas_cumprod), t, x.shape) * x
                       - self._extract(self.sqrt_recipm1_alphas_cumprod, t, x.shape) * eps)
        if clip_denoised:
            x0_pred = torch.clamp(x0_pred, -1, 1)
        mean = (self._extract(self.posterior_mean_coef1, t, x.shape) * x0_pred
                + self._extract(self.posterior_mean_coef2, t, x.shape) * x)
        return (mean, self._extract(self.posterior_variance, t, x.shape),
                self._extract(self.posterior_log_variance_clipped, t, x.shape))

    @torch.no_grad()
    def p_sample(self, model, x, t, y=None, clip_denoised=True, eps=None, x0_pred=None):
        """ddpm.py:197-220"""
        self._require_cuda(x, "DDPM.p_sample")
        lib = _lib.load()
        if eps is None:
            eps = model(x, t, y)
        tt = t.to(x.device)
        if x0_pred is not None or not bool((tt == tt[0]).all()):
            mean, _, logvar = self.p_mean_variance(model, x, t, y, clip_denoised, eps=eps, x0_pred=x0_pred)
            noise = torch.randn_like(x)
            mask = (t != 0).float().view(-1, *([1] * (len(x.shape) - 1)))
            return mean + mask * torch.exp(0.5 * logvar) * noise
        x = x.contiguous().float()
        noise = torch.randn_like(x)
        out = torch.empty_like(x)
        coefs = self._coef_table().to(x.device)
        self._step(lib, x, eps.contiguous().float(), None, noise, out, coefs.data_ptr() + 20 * int(tt[0]),
                   guidance(0.0, 1 if clip_denoised else 0))
        return out

    @torch.no_grad()
    def sample(self, model, shape, y=None, return_all_timesteps=False, noise=None, step_noise=None):
        """ddpm.py:222-252.  `noise` / `step_noise` (optional, extensions): x_T and the per-step N(0,1) draws
        (step_noise[k] is used at t = T-1-k) instead of torch.randn / randn_like."""
        lib = _lib.load()
        device = self.device
        img = torch.randn(shape, device=device) if noise is None else noise.to(device).float().clone()
        self._require_cuda(img, "DDPM.sample")
        if img.shape[0] == 0:  # empty batch: the reference's loop runs S no-op steps and returns the empty tensor
            return img.new_empty((len(self._steps_for_empty()),) + tuple(img.shape)).cpu() if return_all_timesteps else img
        B = shape[0]
        coefs = self._coef_table().to(img.device)
        t_batch = torch.empty((B,), device=img.device, dtype=torch.long)
        nxt = torch.empty_like(img)
        g = guidance(0.0, 1)
        if self._graph_ok(model, return_all_timesteps, step_noise):
            t_seq, coef_seq = self._seq_tables(img.device)
            with self._uniform_t(model):
                return self._graph_loop(model, img, None if y is None else y.to(img.device), True, t_seq, coef_seq, g,
                                        cfg=False, draw_noise=True, desc="Sampling")
        imgs = []
        with self._uniform_t(model):
            for k, i in enumerate(self._bar(reversed(range(0, self.num_timesteps)), "Sampling", self.num_timesteps)):
                t_batch.fill_(i)
                eps = model(img, t_batch, y)
                z = self._draw_like(img) if step_noise is None else step_noise[k].to(img.device).float().contiguous()
                self._step(lib, img, eps.contiguous(), None, z, nxt, coefs.data_ptr() + 20 * i, g)
                img, nxt = nxt, img
                if return_all_timesteps:
                    imgs.append(img.cpu())
        if return_all_timesteps:
            return torch.stack(imgs, dim=0)
        return img

    @torch.no_grad()
    def sample_with_cfg(self, model, shape, y, cfg_scale=3.0, p_threshold=0.995, return_all_timesteps=False,
                        noise=None, step_noise=None):
        """ddpm.py:254-332"""
        if y is None:
            raise ValueError("CFG sampling requires class labels y.")
        if p_threshold is not None and not (0.0 < float(p_threshold) < 1.0):
            raise ValueError("p_threshold must be in (0, 1) or None")
        lib = _lib.load()
        device = self.device
        img = torch.randn(shape, device=device) if noise is None else noise.to(device).float().clone()
        self._require_cuda(img, "DDPM.sample_with_cfg")
        if img.shape[0] == 0:  # empty batch: the reference's loop runs S no-op steps and returns the empty tensor
            return img.new_empty((len(self._steps_for_empty()),) + tuple(img.shape)).cpu() if return_all_timesteps else img
        B = shape[0]
        n = img.numel() // B
        coefs = self._coef_table().to(img.device)
        y = y.to(img.device)
        y_uncond = torch.zeros_like(y)
        t_batch = torch.empty((B,), device=img.device, dtype=torch.long)
        nxt = torch.empty_like(img)
        g = guidance(cfg_scale, 2, n, float(p_threshold)) if p_threshold is not None else guidance(cfg_scale, 1)
        if self._graph_ok(model, return_all_timesteps, step_noise) and hasattr(model, "forward_cfg"):
            t_seq, coef_seq = self._seq_tables(img.device)
            with self._uniform_t(model):
                return self._graph_loop(model, img, y, True, t_seq, coef_seq, g, cfg=True, draw_noise=True,
                                        desc=f"DDPM Sampling with CFG scale {cfg_scale}")
        imgs = []
        with self._uniform_t(model):
            for k, i in enumerate(self._bar(reversed(range(0, self.num_timesteps)),
                                            f"DDPM Sampling with CFG scale {cfg_scale}", self.num_timesteps)):
                t_batch.fill_(i)
                eps_c, eps_u = self._eps_pair(model, img, t_batch, y, y_uncond)
                z = self._draw_like(img) if step_noise is None else step_noise[k].to(img.device).float().contiguous()
                self._step(lib, img, eps_c.contiguous(), eps_u.contiguous(), z, nxt, coefs.data_ptr() + 20 * i, g)
                img, nxt = nxt, img
                if return_all_timesteps:
                    imgs.append(img.cpu())
        if return_all_timesteps:
            return torch.stack(imgs, dim=0)
        return img
