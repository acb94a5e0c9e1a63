"""DDIM drop-in (same constructor, attributes and methods as /root/reference/diffusion/ddim.py:13-351).

The per-step tensor program of the reference (x0 prediction, clamp / dynamic threshold, sigma, direction term,
CFG combine: ~86-185 ATen calls and one host sync per step, ddim.py:174-208,300-339) is ONE fused CUDA kernel
(csrc/sched.cu via dmc_ddim_step).  Per-step scalars come from a [S, 5] device table built once with the
reference's own fp32 expressions, so nothing in the loop synchronises with the host."""

from __future__ import annotations

import contextlib

import torch

from .. import _lib
from ._common import DiffusionBase, guidance


class DDIM(DiffusionBase):
    def __init__(self, num_timesteps=1000, num_inference_steps=50, beta_start=0.0001, beta_end=0.02,
                 beta_schedule="linear", eta=0.0, device="cuda"):
        self.num_inference_steps = num_inference_steps
        self.eta = eta
        self._init_common(num_timesteps, beta_start, beta_end, beta_schedule, device)
        self._setup_inference_timesteps()

    def _setup_inference_timesteps(self):
        # ddim.py:71-85: linspace(T-1, 0, S) on the device, round half to even, long
        ts = torch.linspace(self.num_timesteps - 1, 0, self.num_inference_steps, device=self.device)
        self.inference_timesteps = ts.round().long()
        self._coef_cache = None

    def set_inference_steps(self, num_inference_steps):
        self.num_inference_steps = num_inference_steps
        self._setup_inference_timesteps()

    # ---- coefficient table -----------------------------------------------------------------------
    def _coef_rows(self, t, t_next):
        """[len(t), 5] fp32 rows (dmc_ddim_coef) for timestep pairs; t_next == -1 means alpha_next = 1
        (ddim.py:174-203, same expressions vectorised over the steps)."""
        acp = self.alphas_cumprod
        a = acp[t]
        an = torch.where(t_next >= 0, acp[t_next.clamp(min=0)], torch.ones_like(a))
        sigma = self.eta * torch.sqrt(torch.clamp((1 - an) / (1 - a) * (1 - a / an), min=0.0))
        dirc = torch.sqrt(torch.clamp(1 - an - sigma ** 2, min=0.0))
        return torch.stack([torch.sqrt(1 - a), torch.sqrt(a), torch.sqrt(an), dirc, sigma], dim=1).contiguous()

    def _coef_table(self):
        if self._coef_cache is None:
            ts = self.inference_timesteps
            nxt = torch.cat([ts[1:], torch.full((1,), -1, dtype=ts.dtype, device=ts.device)])
            self._coef_cache = self._coef_rows(ts, nxt)
        return self._coef_cache

    def _seq_tables(self, device):
        """(timesteps int64 [S], coefficient rows fp32 [S, 5]) on `device`, one row per sampling step"""
        c = getattr(self, "_seq_cache", None)
        coefs = self._coef_table()
        if c is None or c[0] is not coefs or c[1] != str(device):
            c = self._seq_cache = (coefs, str(device), self.inference_timesteps.to(device).contiguous(),
                                   coefs.to(device).contiguous())
        return c[2], c[3]

    def _step(self, lib, x, eps_c, eps_u, noise, out, coef_row_ptr, g):
        B = x.shape[0]
        _lib.check(lib.dmc_ddim_step(x.data_ptr(), eps_c.data_ptr(), _lib.ptr(eps_u), _lib.ptr(noise), out.data_ptr(), B,
                                     x.numel() // B, coef_row_ptr, g, _lib.stream_ptr()), "dmc_ddim_step")

    # ---- public sampling API -----------------------------------------------------------------------
    @torch.no_grad()
    def p_sample(self, model, x, t, t_next, y=None, clip_denoised=True, eps=None, x0_pred=None):
        """One DDIM update (ddim.py:154-208).  `x0_pred` (the caller's own x0) is honoured like the reference does."""
        self._require_cuda(x, "DDIM.p_sample")
        lib = _lib.load()
        if eps is None:
            eps = model(x, t, y)
        x = x.contiguous().float()
        eps = eps.contiguous().float()
        tt = t.to(self.alphas_cumprod.device)
        tn = t_next.to(self.alphas_cumprod.device)
        rows = self._coef_rows(tt, tn).to(x.device)
        noise = torch.randn_like(x) if self.eta > 0 else None
        if x0_pred is not None or not bool((rows == rows[0]).all()):
            # caller-supplied x0, or per-sample timesteps: not the sampling hot path -> plain tensor algebra on device
            r = rows.view(x.shape[0], 5, *((1,) * (x.dim() - 1)))
            if x0_pred is None:
                x0_pred = (x - r[:, 0] * eps) / r[:, 1]
            if clip_denoised:
                x0_pred = torch.clamp(x0_pred, -1.0, 1.0)
            out = r[:, 2] * x0_pred + r[:, 3] * eps
            if self.eta > 0:
                out = out + r[:, 4] * noise
            return out
        out = torch.empty_like(x)
        self._step(lib, x, eps, None, noise, out, rows.data_ptr(), guidance(0.0, 1 if clip_denoised else 0))
        return out

    @torch.no_grad()
    def sample(self, model, shape, y=None, return_all_timesteps=False, noise=None, step_noise=None):
        """ddim.py:210-249.  `noise` / `step_noise` (optional, extensions): the x_T and the per-step draws
        (eta > 0) to use instead of torch.randn / randn_like."""
        lib = _lib.load()
        device = self.device
        img = torch.randn(shape, device=device) if noise is None else noise.to(device).float().clone()
        self._require_cuda(img, "DDIM.sample")
        if img.shape[0] == 0:  # empty batch: the reference's loop runs S no-op steps and returns the empty tensor
            return img.new_empty((len(self._steps_for_empty()),) + tuple(img.shape)).cpu() if return_all_timesteps else img
        B = shape[0]
        n = img.numel() // B
        coefs = self._coef_table().to(img.device)
        timesteps = self.inference_timesteps.tolist()
        t_batch = torch.empty((B,), device=img.device, dtype=torch.long)
        nxt = torch.empty_like(img)
        g = guidance(0.0, 1)
        if self._graph_ok(model, return_all_timesteps, step_noise):
            t_seq, coef_seq = self._seq_tables(img.device)
            with self._uniform_t(model):
                return self._graph_loop(model, img, None if y is None else y.to(img.device), False, t_seq, coef_seq, g,
                                        cfg=False, draw_noise=self.eta > 0, desc="DDIM Sampling")
        imgs = []
        with self._uniform_t(model):
            for i, t in enumerate(self._bar(timesteps, "DDIM Sampling")):
                t_batch.fill_(t)
                eps = model(img, t_batch, y)
                z = None
                if self.eta > 0:
                    z = self._draw_like(img) if step_noise is None else step_noise[i].to(img.device).float().contiguous()
                self._step(lib, img, eps.contiguous(), None, z, nxt, coefs.data_ptr() + 20 * i, g)
                img, nxt = nxt, img
                if return_all_timesteps:
                    imgs.append(img.cpu())
        if return_all_timesteps:
            return torch.stack(imgs, dim=0)
        return img

    @torch.no_grad()
    def sample_with_cfg(self, model, shape, y, cfg_scale=3.0, p_threshold=0.995, return_all_timesteps=False, noise=None,
                        step_noise=None):
        """ddim.py:251-346: CFG on eps, dynamic thresholding on x0, DDIM update -- one fused kernel per step."""
        if y is None:
            raise ValueError("CFG sampling requires class labels y.")
        if p_threshold is not None and not (0.0 < float(p_threshold) < 1.0):
            raise ValueError("p_threshold must be in (0, 1) or None")
        lib = _lib.load()
        device = self.device
        img = torch.randn(shape, device=device) if noise is None else noise.to(device).float().clone()
        self._require_cuda(img, "DDIM.sample_with_cfg")
        if img.shape[0] == 0:  # empty batch: the reference's loop runs S no-op steps and returns the empty tensor
            return img.new_empty((len(self._steps_for_empty()),) + tuple(img.shape)).cpu() if return_all_timesteps else img
        B = shape[0]
        n = img.numel() // B
        coefs = self._coef_table().to(img.device)
        timesteps = self.inference_timesteps.tolist()
        y = y.to(img.device)
        y_uncond = torch.zeros_like(y)
        t_batch = torch.empty((B,), device=img.device, dtype=torch.long)
        nxt = torch.empty_like(img)
        g = guidance(cfg_scale, 2, n, p_threshold) if p_threshold is not None else guidance(cfg_scale, 1)
        if self._graph_ok(model, return_all_timesteps, step_noise) and hasattr(model, "forward_cfg"):
            t_seq, coef_seq = self._seq_tables(img.device)
            with self._uniform_t(model):
                return self._graph_loop(model, img, y, False, t_seq, coef_seq, g, cfg=True, draw_noise=self.eta > 0,
                                        desc=f"DDIM sampling with CFG scale {cfg_scale}")
        imgs = []
        with self._uniform_t(model):
            for i, t in enumerate(self._bar(timesteps, f"DDIM sampling with CFG scale {cfg_scale}")):
                t_batch.fill_(t)
                eps_c, eps_u = self._eps_pair(model, img, t_batch, y, y_uncond)
                z = None
                if self.eta > 0:
                    z = self._draw_like(img) if step_noise is None else step_noise[i].to(img.device).float().contiguous()
                self._step(lib, img, eps_c.contiguous(), eps_u.contiguous(), z, nxt, coefs.data_ptr() + 20 * i, g)
                img, nxt = nxt, img
                if return_all_timesteps:
                    imgs.append(img.cpu())
        if return_all_timesteps:
            return torch.stack(imgs, dim=0)
        return img
