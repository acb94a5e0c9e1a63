"""DiT denoiser drop-in: constructor, attributes and parameter names of /root/reference/models/dit.py:154-295.

Parameters are registered with the reference's names/shapes (state_dict contract).  The native
sm_100a forward (tcgen05 GEMMs with fused bias/GELU/gate-residual epilogues, LN+modulate, patchify/unpatchify)
is wired in `forward`."""

from __future__ import annotations

import math
from typing import Tuple

import torch
import torch.nn as nn

from .. import _lib
from .unet import _register


class DiT(nn.Module):
    def __init__(self, img_size: Tuple[int, int] = (32, 32), patch_size=2, in_channels=3, hidden_size=768, depth=12,
                 num_heads=12, mlp_ratio=4.0, num_classes=None, dropout=0.1):
        super().__init__()
        img_h, img_w = (img_size, img_size) if isinstance(img_size, int) else img_size
        self.img_size = (img_h, img_w)
        self.patch_size = patch_size
        self.in_channels = in_channels
        self.out_channels = in_channels
        self.hidden_size = hidden_size
        self.depth = depth
        self.num_heads = num_heads
        self.mlp_ratio = mlp_ratio
        self.num_classes = num_classes
        self.dropout = dropout
        self.h_tokens = img_h // patch_size
        self.w_tokens = img_w // patch_size
        self._init_parameters()

    def _init_parameters(self):
        """models/dit.py:233-247: Xavier-uniform linears with zero bias, pos-emb N(0, 0.02^2), zero-init adaLN and
        final layer; conv patch embed and label table keep PyTorch defaults."""
        hs, p, c = self.hidden_size, self.patch_size, self.in_channels
        hid = int(hs * self.mlp_ratio)

        def xavier(name, cout, cin, zero=False):
            w = torch.zeros(cout, cin)
            if not zero:
                nn.init.xavier_uniform_(w)
            _register(self, name + ".weight", w)
            _register(self, name + ".bias", torch.zeros(cout))

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
            w = torch.empty(3 * hs, hs)
            nn.init.xavier_uniform_(w)
            _register(self, b + ".attn.in_proj_weight", w)
            _register(self, b + ".attn.in_proj_bias", torch.zeros(3 * hs))
            xavier(b + ".attn.out_proj", hs, hs)
            xavier(b + ".mlp.0", hid, hs)
            xavier(b + ".mlp.3", hs, hid)
            xavier(b + ".adaLN_modulation.1", 6 * hs, hs, zero=True)
        xavier("final_layer.linear", p * p * c, hs, zero=True)
        xavier("final_layer.adaLN_modulation.1", 2 * hs, hs, zero=True)

    def forward(self, x, t, y=None):
        if not (isinstance(x, torch.Tensor) and x.is_cuda):
            raise _lib.DmcError("DiT.forward: CUDA tensors only -- the B200 hot path has no CPU / PyTorch fallback")
        raise NotImplementedError("native DiT forward: not wired yet in this round (UNet path first, SURVEY.md section 7.1 step 9)")
