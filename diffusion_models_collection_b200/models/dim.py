"""``DiM`` drop-in (reference: models/dim.py:208-340) -- the variant the reference itself runs when ``mamba_ssm`` is not
installed (models/dim.py:103-117, 135-138): every "Mamba" block is ``nn.MultiheadAttention(hidden, num_heads=8)`` behind an
adaLN-modulated LayerNorm, followed by an adaLN-modulated feed-forward block.  That is a DiT block whose six modulation
vectors come from two 3-chunk linears instead of one 6-chunk linear, and whose LayerNorms carry an affine (weight, bias):

    h = (LN(x) * w + b) * (1 + scale) + shift  =  LN(x) * (1 + scale') + shift'
        scale' = w * scale + (w - 1),  shift' = b * scale + b + shift          -- both still LINEAR in SiLU(c)

so the LayerNorm affine folds into the adaLN weights at pack time and the whole model runs on the DiT launch plan
(models/dit.py: dit_cond, patch_embed, ln_modulate, the tcgen05 GEMMs and attention) with 8 heads.  Parameter names, shapes and
the constructor are the reference DiM's, so its checkpoints load with ``strict=True``.  True Mamba (selective scan) is outside
the hot path: a checkpoint trained WITH mamba_ssm has ``mamba.in_proj.weight``-style keys and
fails to load here, loudly.  Head dim = hidden / 8 must be 64 (tcgen05 attention, hidden 512) or 32 (CUDA-core flash kernel,
hidden 256); the reference default hidden 768 (head dim 96) constructs and loads but raises on its first forward."""

from __future__ import annotations

import math
from typing import Tuple

import torch
import torch.nn as nn

from .dit import DiT
from .unet import _register


class DiM(DiT):
    _train_supported = False

    def __init__(self, img_size: Tuple[int, int] = (32, 32), patch_size=2, in_channels=3, hidden_size=768, depth=12,
                 state_size=16, mlp_ratio=4.0, num_classes=None, dropout=0.1):
        self.state_size = state_size
        super().__init__(img_size=img_size, patch_size=patch_size, in_channels=in_channels, hidden_size=hidden_size, depth=depth,
                         num_heads=8, mlp_ratio=mlp_ratio, num_classes=num_classes, dropout=dropout)

    def _init_parameters(self):
        """models/dim.py:283-305: Xavier-uniform linears with zero bias (nn.MultiheadAttention: Xavier in_proj, zero biases),
        LayerNorm 1 / 0, pos-emb N(0, 0.02^2), zero-init adaLN and final layer; patch conv and label table keep PyTorch defaults"""
        hs, p, c = self.hidden_size, self.patch_size, self.in_channels
        hid = int(hs * self.mlp_ratio)

        def xavier(name, cout, cin, zero=False):
            w = torch.zeros(cout, cin)
            if not zero:
                nn.init.xavier_uniform_(w)
            _register(self, name + ".weight", w)
            _register(self, name + ".bias", torch.zeros(cout))

        def norm(name):
            _register(self, name + ".weight", torch.ones(hs))
            _register(self, name + ".bias", torch.zeros(hs))

        _register(self, "pos_embed", torch.randn(1, self.h_tokens * self.w_tokens, hs) * 0.02)
        bound = 1.0 / math.sqrt(c * p * p)
        _register(self, "x_embedder.proj.weight", torch.empty(hs, c, p, p).uniform_(-bound, bound))
        _register(self, "x_embedder.proj.bias", torch.empty(hs).uniform_(-bound, bound))
        xavier("t_embedder.mlp.0", hs, 256)
        xavier("t_embedder.mlp.2", hs, hs)
        if self.num_classes is not None:
            w = torch.randn(self.num_classes + 1, hs)
            w[0].zero_()
            _register(self, "y_embedder.embedding_table.weight", w)
        for i in range(self.depth):
            b = f"blocks.{i}"
            norm(b + ".mamba_block.norm")
            w = torch.empty(3 * hs, hs)
            nn.init.xavier_uniform_(w)
            _register(self, b + ".mamba_block.mamba.in_proj_weight", w)
            _register(self, b + ".mamba_block.mamba.in_proj_bias", torch.zeros(3 * hs))
            xavier(b + ".mamba_block.mamba.out_proj", hs, hs)
            xavier(b + ".mamba_block.adaLN_modulation.1", 3 * hs, hs, zero=True)
            norm(b + ".ff_block.norm")
            xavier(b + ".ff_block.mlp.0", hid, hs)
            xavier(b + ".ff_block.mlp.3", hs, hid)
            xavier(b + ".ff_block.adaLN_modulation.1", 3 * hs, hs, zero=True)
        norm("final_layer.norm_final")
        xavier("final_layer.linear", p * p * c, hs, zero=True)
        xavier("final_layer.adaLN_modulation.1", 2 * hs, hs, zero=True)

    @staticmethod
    def _fold_ln_affine(w_mod, b_mod, ln_w, ln_b, chunks):
        """adaLN linear [chunks * hs, hs] (chunk 0 = shift, 1 = scale, 2 = gate if present) with the LayerNorm affine folded in"""
        hs = ln_w.numel()
        W, B = w_mod.clone().view(chunks, hs, -1), b_mod.clone().view(chunks, hs)
        W[0] = w_mod.view(chunks, hs, -1)[0] + ln_b[:, None] * w_mod.view(chunks, hs, -1)[1]
        B[0] = b_mod.view(chunks, hs)[0] + ln_b * b_mod.view(chunks, hs)[1] + ln_b
        W[1] = ln_w[:, None] * w_mod.view(chunks, hs, -1)[1]
        B[1] = ln_w * b_mod.view(chunks, hs)[1] + ln_w - 1.0
        return W.reshape(chunks * hs, -1), B.reshape(chunks * hs)

    def _canonical_state(self, sd):
        out = {k: sd[k] for k in ("pos_embed", "x_embedder.proj.weight", "x_embedder.proj.bias", "t_embedder.mlp.0.weight",
                                  "t_embedder.mlp.0.bias", "t_embedder.mlp.2.weight", "t_embedder.mlp.2.bias",
                                  "final_layer.linear.weight", "final_layer.linear.bias")}
        if "y_embedder.embedding_table.weight" in sd:
            out["y_embedder.embedding_table.weight"] = sd["y_embedder.embedding_table.weight"]
        for i in range(self.depth):
            b, m, f = f"blocks.{i}", f"blocks.{i}.mamba_block", f"blocks.{i}.ff_block"
            out[b + ".attn.in_proj_weight"], out[b + ".attn.in_proj_bias"] = sd[m + ".mamba.in_proj_weight"], sd[m + ".mamba.in_proj_bias"]
            out[b + ".attn.out_proj.weight"], out[b + ".attn.out_proj.bias"] = sd[m + ".mamba.out_proj.weight"], sd[m + ".mamba.out_proj.bias"]
            for j in ("0", "3"):
                out[f"{b}.mlp.{j}.weight"], out[f"{b}.mlp.{j}.bias"] = sd[f"{f}.mlp.{j}.weight"], sd[f"{f}.mlp.{j}.bias"]
            wa, ba = self._fold_ln_affine(sd[m + ".adaLN_modulation.1.weight"], sd[m + ".adaLN_modulation.1.bias"],
                                          sd[m + ".norm.weight"], sd[m + ".norm.bias"], 3)
            wf, bf = self._fold_ln_affine(sd[f + ".adaLN_modulation.1.weight"], sd[f + ".adaLN_modulation.1.bias"],
                                          sd[f + ".norm.weight"], sd[f + ".norm.bias"], 3)
            out[b + ".adaLN_modulation.1.weight"] = torch.cat([wa, wf], dim=0).contiguous()
            out[b + ".adaLN_modulation.1.bias"] = torch.cat([ba, bf], dim=0).contiguous()
        w, bb = self._fold_ln_affine(sd["final_layer.adaLN_modulation.1.weight"], sd["final_layer.adaLN_modulation.1.bias"],
                                     sd["final_layer.norm_final.weight"], sd["final_layer.norm_final.bias"], 2)
        out["final_layer.adaLN_modulation.1.weight"], out["final_layer.adaLN_modulation.1.bias"] = w.contiguous(), bb.contiguous()
        return out
