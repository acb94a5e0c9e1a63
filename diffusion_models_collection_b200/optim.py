"""FusedAdamW: the optimizer side of the reference's training iteration (utils/trainer.py:256-262) in two native launches.

The reference does, per optimizer step, ``clip_grad_norm_(model.parameters(), 1.0)``, ``AdamW.step()`` and (rank 0) an EMA update
that loops over the state dict -- in PyTorch that is a foreach norm, a foreach scale, the fused AdamW kernels and 2 x 357 small EMA
kernels.  ``FusedAdamW`` is a ``torch.optim.Optimizer`` with ``torch.optim.AdamW``'s hyper-parameters, state keys (``step``,
``exp_avg``, ``exp_avg_sq``) and update rule that runs ONE gradient-norm pass and ONE update pass over all parameters
(csrc/optim.cu), with gradient clipping (``max_grad_norm``) and the EMA copy (``ema_params``) folded into the update pass:

    opt = FusedAdamW(model.parameters(), lr=2e-4, weight_decay=1e-4, max_grad_norm=1.0,
                     ema_params=ema_model.parameters(), ema_decay=0.9999)
    loss.backward(); opt.step(); opt.zero_grad()

There is no CPU path: parameters and gradients must be contiguous fp32 CUDA tensors.  A parameter whose ``.grad`` is None is
left alone for that step (as in torch); bias correction uses one step count per parameter group.
"""

from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib

_CHUNK = 16384


class FusedAdamW(torch.optim.Optimizer):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2, max_grad_norm=None, ema_params=None,
                 ema_decay=0.9999):
        if lr < 0 or eps < 0 or not (0 <= betas[0] < 1 and 0 <= betas[1] < 1) or weight_decay < 0:
            raise ValueError("FusedAdamW: invalid hyper-parameter")
        self._tables = None
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay))
        self.max_grad_norm = float(max_grad_norm) if max_grad_norm else 0.0
        self.ema_decay = float(ema_decay)
        self._ema = list(ema_params) if ema_params is not None else None
        flat = [p for g in self.param_groups for p in g["params"]]
        if self._ema is not None and len(self._ema) != len(flat):
            raise ValueError("FusedAdamW: ema_params must pair one to one with params")
        self.last_grad_norm = None  # device tensor [1] (what clip_grad_norm_ returns), set by step()
        self._tables = None

    # ------------------------------------------------------------------------------------------------------------------
    def _build_tables(self, device):
        """static part of the work list: the chunk table and the item columns that never change (p, m, v, ema, n)"""
        ema_of = {}
        if self._ema is not None:
            for p, e in zip([p for g in self.param_groups for p in g["params"]], self._ema):
                ema_of[id(p)] = e
        groups = []
        for g in self.param_groups:
            ps = [p for p in g["params"] if p.requires_grad]
            for p in ps:
                if not (p.is_cuda and p.dtype == torch.float32 and p.is_contiguous()):
                    raise _lib.DmcError("FusedAdamW: contiguous fp32 CUDA parameters only -- the B200 path has no CPU fallback")
                st = self.state[p]
                if not st:
                    st["step"] = torch.zeros((), dtype=torch.float32)
                    st["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                    st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
            items = np.zeros((len(ps), 6), dtype=np.int64)
            chunks = []
            for i, p in enumerate(ps):
                st = self.state[p]
                e = ema_of.get(id(p))
                if e is not None and not (e.is_cuda and e.dtype == torch.float32 and e.is_contiguous() and e.shape == p.shape):
                    raise _lib.DmcError("FusedAdamW: EMA tensors must be contiguous fp32 CUDA tensors of the parameter's shape")
                items[i] = (p.data_ptr(), 0, st["exp_avg"].data_ptr(), st["exp_avg_sq"].data_ptr(),
                            e.data_ptr() if e is not None else 0, p.numel())
                chunks += [(i, s) for s in range(0, p.numel(), _CHUNK)]
            ch = np.zeros((len(chunks), 2), dtype=np.int64)
            for j, (i, s) in enumerate(chunks):
                ch[j] = (i, s)  # (item, start): int32 item + int32 pad share the first 8 bytes (little endian)
            groups.append(dict(params=ps, items_host=[torch.from_numpy(items.copy()).pin_memory() for _ in range(4)],
                               copied=[None] * 4, turn=0,
                               items_dev=torch.empty((len(ps), 6), dtype=torch.int64, device=device),
                               chunks_dev=torch.from_numpy(ch).to(device), n_chunks=len(chunks),
                               partial=torch.empty(max(len(chunks), 1), dtype=torch.float32, device=device), group=g))
        self._tables = dict(device=device, groups=groups,
                            norms=torch.zeros(len(groups) + 1, dtype=torch.float32, device=device))

    def load_state_dict(self, state_dict):
        """(torch.optim.AdamW checkpoints load too: same state keys.)  Loaded moments are new tensors: rebuild the pointer tables."""
        super().load_state_dict(state_dict)
        for st in self.state.values():  # torch's fused AdamW keeps `step` on the device: ours is a host counter
            if "step" in st:
                st["step"] = torch.as_tensor(float(st["step"]), dtype=torch.float32)
        self._tables = None

    def add_param_group(self, param_group):
        super().add_param_group(param_group)
        self._tables = None

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        lib = _lib.load()
        first = next(p for g in self.param_groups for p in g["params"])
        if self._tables is None or self._tables["device"] != first.device:
            self._build_tables(first.device)
        tb = self._tables
        st_ptr = _lib.stream_ptr()
        with torch.cuda.device(tb["device"]):
            # gradient pointers change every step (fresh tensors from autograd): refresh that column of the item tables
            for gi, t in enumerate(tb["groups"]):
                grads = []
                for p in t["params"]:
                    g = p.grad
                    if g is None:  # left alone this step, as torch.optim.AdamW does (NULL gradient pointer)
                        grads.append(0)
                        continue
                    if not (g.is_cuda and g.dtype == torch.float32 and g.is_contiguous()):
                        g = p.grad = g.contiguous().float()
                    grads.append(g.data_ptr())
                    self.state[p]["step"] += 1
                # four rotating pinned staging buffers: the host may run a whole step ahead of the device, so a buffer is
                # rewritten only after the asynchronous copy that read it has completed
                k = t["turn"] = (t["turn"] + 1) % 4
                if t["copied"][k] is not None:
                    t["copied"][k].synchronize()
                host = t["items_host"][k]
                host[:, 1] = torch.tensor(grads, dtype=torch.int64)
                t["items_dev"].copy_(host, non_blocking=True)
                ev = torch.cuda.Event()
                ev.record()
                t["copied"][k] = ev
            clip = self.max_grad_norm > 0
            if clip:
                # global norm = sqrt of the sum over the groups' squared norms
                for gi, t in enumerate(tb["groups"]):
                    _lib.check(lib.dmc_opt_grad_norm(t["items_dev"].data_ptr(), t["chunks_dev"].data_ptr(), t["n_chunks"],
                                                     t["partial"].data_ptr(), tb["norms"][gi:].data_ptr(), st_ptr), "dmc_opt_grad_norm")
                n = len(tb["groups"])
                if n == 1:
                    total = tb["norms"][0:1]
                else:
                    tb["norms"][n] = tb["norms"][:n].square().sum().sqrt()
                    total = tb["norms"][n:n + 1]
                self.last_grad_norm = total
            for t in tb["groups"]:
                g = t["group"]
                step = max(1.0, max(float(self.state[p]["step"]) for p in t["params"]))
                h = _lib.AdamWDesc()
                h.lr, h.beta1, h.beta2, h.eps, h.weight_decay = g["lr"], g["betas"][0], g["betas"][1], g["eps"], g["weight_decay"]
                h.bias_correction1, h.bias_correction2 = 1.0 - g["betas"][0] ** step, 1.0 - g["betas"][1] ** step
                h.max_norm = self.max_grad_norm if clip else 0.0
                h.ema_decay = self.ema_decay if self._ema is not None else 0.0
                _lib.check(lib.dmc_opt_adamw_step(t["items_dev"].data_ptr(), t["chunks_dev"].data_ptr(), t["n_chunks"], C.byref(h),
                                                  self.last_grad_norm.data_ptr() if clip else None, st_ptr), "dmc_opt_adamw_step")
                # The kernel wrote the parameters (and EMA tensors) through raw device pointers, which autograd's version
                # counters do not see.  The native UNet / DiT re-pack their bf16 GEMM operands when a parameter version
                # changes (models/unet.py:_ensure_packed), so every tensor the launch touched is bumped here -- without it the
                # forward would keep training on the weights packed before the first step.
                touched = [p for p in t["params"] if p.grad is not None]
                torch.autograd.graph.increment_version(touched)
                if self._ema is not None:
                    ema_of = t.get("ema_list")
                    if ema_of is None:
                        pairs = dict(zip([id(p) for g_ in self.param_groups for p in g_["params"]], self._ema))
                        ema_of = t["ema_list"] = {id(p): pairs[id(p)] for p in t["params"] if id(p) in pairs}
                    torch.autograd.graph.increment_version([ema_of[id(p)] for p in touched if id(p) in ema_of])
        return loss
