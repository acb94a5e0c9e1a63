"""Shared host logic of the DDPM / DDIM drop-ins: schedule tables (built with the reference's torch expressions on
``device`` so that they are bit-equal to the reference's on the same device), the training-side helpers
(q_sample / p_losses / _extract) and the glue that hands a sampling step to the fused CUDA kernels."""

from __future__ import annotations

import contextlib
import ctypes as C
import math
import os

import torch
import torch.nn.functional as F

from .. import _lib


def make_betas(num_timesteps, beta_start, beta_end, beta_schedule, device):
    # /root/reference/diffusion/ddpm.py:38-46, :73-82 (same expressions, same op order)
    if beta_schedule == "linear":
        return torch.linspace(beta_start, beta_end, num_timesteps, device=device)
    if beta_schedule == "cosine":
        s = 0.008
        x = torch.linspace(0, num_timesteps, num_timesteps + 1, device=device)
        acp = torch.cos(((x / num_timesteps) + s) / (1 + s) * torch.pi * 0.5) ** 2
        acp = acp / acp[0]
        return torch.clip(1 - (acp[1:] / acp[:-1]), 0.0001, 0.9999)
    if beta_schedule == "quadratic":
        return torch.linspace(beta_start ** 0.5, beta_end ** 0.5, num_timesteps, device=device) ** 2
    raise ValueError(f"Unknown beta schedule: {beta_schedule}")


def quantile_rank(n, q):
    """(lower rank, upper rank, lerp weight) of torch.quantile for an fp32 row of length n (fp32 rank arithmetic)."""
    rank = (torch.tensor(float(q), dtype=torch.float32) * (n - 1)).item()
    lo = int(math.floor(rank))
    hi = int(math.ceil(rank))
    w = float(torch.tensor(rank, dtype=torch.float32) - torch.tensor(float(lo), dtype=torch.float32))
    return lo, hi, w


def guidance(cfg_scale, clip_mode, n_per_sample=0, p_threshold=None):
    g = _lib.Guidance()
    g.cfg_scale = float(cfg_scale)
    g.clip_mode = int(clip_mode)
    if clip_mode == 2:
        g.q_lo, g.q_hi, g.q_weight = quantile_rank(n_per_sample, p_threshold)
    return g


class DiffusionBase:
    """Attributes and training-side methods shared by DDPM and DDIM (reference: diffusion/ddpm.py:27-149,
    diffusion/ddim.py:27-152)."""

    progress = True  # show the reference's tqdm bars
    # Replay ONE captured CUDA graph ("advance counter -> denoiser forward -> fused scheduler step") per sampling step
    # when the denoiser is one of the native models; DMC_CUDA_GRAPH=0 keeps the launch-by-launch loop.
    use_cuda_graph = os.environ.get("DMC_CUDA_GRAPH", "1") != "0"

    def _init_common(self, num_timesteps, beta_start, beta_end, beta_schedule, device):
        self.num_timesteps = num_timesteps
        self.device = device
        self.betas = make_betas(num_timesteps, beta_start, beta_end, beta_schedule, device)
        self.alphas = 1.0 - self.betas
        self.alphas_cumprod = torch.cumprod(self.alphas, dim=0)
        self.sqrt_alphas_cumprod = torch.sqrt(self.alphas_cumprod)
        self.sqrt_one_minus_alphas_cumprod = torch.sqrt(1.0 - self.alphas_cumprod)

    # ---- training side -------------------------------------------------------------------------
    def _extract(self, a, t, x_shape):
        out = a.to(t.device)[t]
        return out.reshape(t.shape[0], *((1,) * (len(x_shape) - 1)))

    def q_sample(self, x_start, t, noise=None):
        if noise is None:
            noise = torch.randn_like(x_start)
        if x_start.is_cuda and x_start.dtype == torch.float32:
            lib = _lib.load()
            x0 = x_start.contiguous()
            nz = noise.contiguous()
            tt = t.to(device=x0.device, dtype=torch.long).contiguous()
            sa = self.sqrt_alphas_cumprod.to(x0.device)
            s1 = self.sqrt_one_minus_alphas_cumprod.to(x0.device)
            out = torch.empty_like(x0)
            B = x0.shape[0]
            _lib.check(lib.dmc_q_sample(x0.data_ptr(), nz.data_ptr(), tt.data_ptr(), sa.data_ptr(), s1.data_ptr(),
                                        out.data_ptr(), B, x0.numel() // B, _lib.stream_ptr()), "dmc_q_sample")
            return out
        raise _lib.DmcError("q_sample needs fp32 CUDA tensors: the B200 hot path has no CPU fallback")

    def p_losses(self, model, x_start, t, y=None, noise=None, loss_type="l2"):
        if loss_type not in ("l1", "l2", "huber"):
            raise ValueError(f"Unknown loss type: {loss_type}")
        if noise is None:
            noise = torch.randn_like(x_start)
        x_noisy = self.q_sample(x_start, t, noise)
        predicted = model(x_noisy, t, y)
        if loss_type == "l1":
            return F.l1_loss(noise, predicted)
        if loss_type == "l2":
            return F.mse_loss(noise, predicted)
        return F.smooth_l1_loss(noise, predicted)

    # ---- sampling glue ---------------------------------------------------------------------------
    def _bar(self, it, desc, total=None):
        if not self.progress:
            return it
        from tqdm import tqdm

        return tqdm(it, desc=desc, total=total)

    # ---- per-step noise ---------------------------------------------------------------------------
    _noise_shard = None  # (global batch, lo, hi) when this process samples one shard of a larger batch (sharding.py)

    def _draw_like(self, img):
        """The per-step N(0,1) draw (DDPM every step, DDIM when eta > 0; ddpm.py:216, ddim.py:205).  On a shard of a
        larger batch every rank draws the GLOBAL tensor and keeps its own rows: the generator advances exactly as in the
        single-process run, so an N-rank run reproduces it image for image."""
        sh = self._noise_shard
        if sh is None:
            return torch.randn_like(img)
        total, lo, hi = sh
        return torch.randn((total,) + tuple(img.shape[1:]), device=img.device, dtype=img.dtype)[lo:hi]

    def _steps_for_empty(self):
        """the steps a sampling loop would run (length of the trajectory returned with return_all_timesteps)"""
        ts = getattr(self, "inference_timesteps", None)
        return range(len(ts)) if ts is not None else range(self.num_timesteps)

    # ---- whole-loop CUDA graph ------------------------------------------------------------------
    def _graph_ok(self, model, return_all_timesteps, step_noise):
        return (self.use_cuda_graph and not return_all_timesteps and step_noise is None
                and getattr(model, "graph_capturable", False))

    def _graph_loop(self, model, img, y, ddpm, t_seq, coef_seq, g, cfg, draw_noise, desc):
        """Runs `len(t_seq)` sampling steps by replaying one captured graph.  Step k reads its timestep t_seq[k] and its
        coefficient row coef_seq[k] on the DEVICE (dmc_advance / dmc_*_step_at), so nothing in the loop touches the
        host; x_t is updated in place.  Per-step noise (DDPM, DDIM eta > 0) is torch.randn_like inside the graph: the
        same Philox stream, in the same order, as the launch-by-launch loop and the reference (ddpm.py:216)."""
        lib = _lib.load()
        B = img.shape[0]
        n = img.numel() // B
        S = int(t_seq.numel())
        # Refresh the packed bf16 operands BEFORE the cache lookup: a replay runs no Python of the model, so weights changed
        # in place since the capture (optimizer step, EMA update, load_state_dict) would otherwise never be re-packed and
        # the graph would keep sampling from the weights of the first call.  The full (non-training) refresh also rebuilds
        # the Upsample phase weights a training step skips.  In-place re-packs keep plans and graphs valid; a storage
        # change replaces model._packed and invalidates the entry below.
        ensure = getattr(model, "_ensure_packed", None)
        if callable(ensure):
            with torch.cuda.device(img.device):
                ensure(img.device)
        key = (self._model_token(model), bool(ddpm), bool(cfg), tuple(img.shape), y is not None, float(g.cfg_scale), int(g.clip_mode),
               int(g.q_lo), int(g.q_hi), float(g.q_weight), bool(draw_noise), S, id(coef_seq), id(t_seq), str(img.device),
               self._noise_shard)
        ent = getattr(self, "_graph_cache", None)
        if ent is not None and (ent["key"] != key or ent["packed"] is not getattr(model, "_packed", None)):
            ent = self._graph_cache = None
        if ent is None:
            dev = img.device
            x = torch.empty_like(img)
            t_batch = torch.zeros((B,), device=dev, dtype=torch.long)
            counter = torch.zeros((2,), device=dev, dtype=torch.int32)
            ybuf = torch.zeros((B,), device=dev, dtype=torch.long) if y is not None else None
            yzero = torch.zeros_like(ybuf) if cfg else None
            step_at = lib.dmc_ddpm_step_at if ddpm else lib.dmc_ddim_step_at

            def model_call():
                if cfg:
                    return self._eps_pair(model, x, t_batch, ybuf, yzero)
                return model(x, t_batch, ybuf), None

            def one_step():
                _lib.check(lib.dmc_advance(counter.data_ptr(), t_seq.data_ptr(), t_batch.data_ptr(), B,
                                           _lib.stream_ptr()), "dmc_advance")
                eps_c, eps_u = model_call()
                z = self._draw_like(x) if draw_noise else None
                _lib.check(step_at(x.data_ptr(), eps_c.data_ptr(), _lib.ptr(eps_u), _lib.ptr(z), x.data_ptr(), B, n,
                                   coef_seq.data_ptr(), counter.data_ptr() + 4, g, _lib.stream_ptr()), "dmc_step_at")

            x.zero_()
            side = torch.cuda.Stream(device=dev)
            side.wait_stream(torch.cuda.current_stream(dev))
            with torch.cuda.stream(side):
                model_call()  # builds the plans / sets kernel attributes outside the capture; draws no random numbers
            torch.cuda.current_stream(dev).wait_stream(side)
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                one_step()
            packed = getattr(model, "_packed", None)
            ent = self._graph_cache = dict(key=key, packed=packed, graph=graph, x=x, t_batch=t_batch, counter=counter,
                                           ybuf=ybuf, yzero=yzero, t_seq=t_seq, coef_seq=coef_seq, g=g)
        ent["x"].copy_(img)
        if y is not None:
            ent["ybuf"].copy_(y)
        ent["counter"].zero_()
        for _ in self._bar(range(S), desc, S):
            ent["graph"].replay()
        return ent["x"].clone()

    @staticmethod
    def _model_token(model):
        """identity of a denoiser for the graph cache: a token object the model owns (created on first use and dropped with
        the model), not id(model) -- CPython reuses ids after garbage collection, which could replay a stale graph"""
        tok = getattr(model, "_dmc_graph_token", None)
        if tok is None:
            tok = object()
            try:
                object.__setattr__(model, "_dmc_graph_token", tok)
            except Exception:  # a callable without a __dict__: fall back to the object itself (kept alive by the cache key)
                return model
        return tok

    @staticmethod
    def _uniform_t(model):
        """context in which the native denoisers may assume all timesteps of a call are equal (sampling loops)"""
        ctx = getattr(model, "uniform_timesteps", None)
        return ctx() if callable(ctx) else contextlib.nullcontext()

    @staticmethod
    def _require_cuda(t, what):
        if not (isinstance(t, torch.Tensor) and t.is_cuda):
            raise _lib.DmcError(f"{what}: the sampler runs on a CUDA device only (no CPU fallback); got device "
                                f"{getattr(t, 'device', None)}")

    @staticmethod
    def _eps_pair(model, img, t_batch, y, y_uncond):
        """(eps_cond, eps_uncond): ONE 2B-image forward when the model is the native denoiser (identical weights, only
        the label differs, nothing in the model mixes samples); otherwise two calls in the reference's order."""
        fwd = getattr(model, "forward_cfg", None)
        if fwd is not None:
            return fwd(img, t_batch, y)
        return model(img, t_batch, y), model(img, t_batch, y_uncond)
