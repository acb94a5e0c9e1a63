"""CPU ORACLE (test infrastructure only) -- fp32 restatement of the reference denoisers.

This file is NOT part of the product path.  Only ``tests/``, ``__graft_entry__.smoke()`` and
``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may import it, and only as the checker
(or as the timed CPU baseline), never as a fallback for the CUDA kernels.

It restates, in plain functional PyTorch fp32 on the CPU, what the reference computes:

* ``unet_forward``  <- /root/reference/models/unet.py:243-292 (UNet.forward), :62-72 (ResidualBlock),
  :84-99 (AttentionBlock), :108-109 (Downsample), :118-120 (Upsample), :18-25 (TimeEmbedding).
* ``dit_forward``   <- /root/reference/models/dit.py:263-295 (DiT.forward), :111-132 (DiTBlock),
  :146-151 (FinalLayer), :23-27 (PatchEmbed), :42-55 (TimestepEmbedder), :249-261 (unpatchify).

The arithmetic itself lives in PyTorch/ATen (third party; the reference pins only ``torch>=2.0.0``,
requirements.txt:1; effective pin = this image's torch 2.11.0+cu128).  Parity pin: the reference has
no tests or golden vectors, so this oracle is pinned against the *live reference* imported from
/root/reference in the build container (tests/golden/make_golden.py writes the fixtures,
tests/test_oracle_golden.py replays them on any box).
"""

from __future__ import annotations

import math

import torch
import torch.nn.functional as F

from diffusion_models_collection_b200.synth import unet_block_structure


def _unet_time_embedding(t, dim):
    # models/unet.py:18-25 -- divisor (half-1), order [sin | cos], t promoted by the multiply.
    half = dim // 2
    k = math.log(10000) / (half - 1)
    freqs = torch.exp(torch.arange(half, device=t.device) * -k)
    arg = t[:, None] * freqs[None, :]
    return torch.cat((arg.sin(), arg.cos()), dim=-1)


def _resblock(sd, p, x, t_emb, y_emb):
    # models/unet.py:62-72
    h = F.group_norm(x, 8, sd[p + ".conv1.0.weight"], sd[p + ".conv1.0.bias"], eps=1e-5)
    h = F.conv2d(F.silu(h), sd[p + ".conv1.2.weight"], sd[p + ".conv1.2.bias"], padding=1)
    h = h + F.linear(F.silu(t_emb), sd[p + ".time_mlp.1.weight"], sd[p + ".time_mlp.1.bias"])[:, :, None, None]
    if (p + ".label_proj.1.weight") in sd and y_emb is not None:
        h = h + F.linear(F.silu(y_emb), sd[p + ".label_proj.1.weight"])[:, :, None, None]
    h = F.group_norm(h, 8, sd[p + ".conv2.0.weight"], sd[p + ".conv2.0.bias"], eps=1e-5)
    h = F.conv2d(F.silu(h), sd[p + ".conv2.3.weight"], sd[p + ".conv2.3.bias"], padding=1)
    if (p + ".shortcut.weight") in sd:
        x = F.conv2d(x, sd[p + ".shortcut.weight"], sd[p + ".shortcut.bias"])
    return h + x


def _attnblock(sd, p, x, heads=4):
    # models/unet.py:84-99 ; qkv channel order is [q|k|v][head][dim]
    B, C, H, W = x.shape
    h = F.group_norm(x, 8, sd[p + ".norm.weight"], sd[p + ".norm.bias"], eps=1e-5)
    qkv = F.conv2d(h, sd[p + ".qkv.weight"], sd[p + ".qkv.bias"])
    qkv = qkv.reshape(B, 3, heads, C // heads, H * W).permute(1, 0, 2, 4, 3)
    q, k, v = qkv[0], qkv[1], qkv[2]
    a = torch.softmax(torch.matmul(q, k.transpose(-2, -1)) / math.sqrt(C // heads), dim=-1)
    h = torch.matmul(a, v).permute(0, 1, 3, 2).reshape(B, C, H, W)
    return x + F.conv2d(h, sd[p + ".proj.weight"], sd[p + ".proj.bias"])


def _run_entry(sd, prefix, layers, h, t_emb, y_emb):
    for j, l in enumerate(layers):
        p = f"{prefix}.{j}"
        if l[0] == "res":
            h = _resblock(sd, p, h, t_emb, y_emb)
        elif l[0] == "attn":
            h = _attnblock(sd, p, h)
        elif l[0] == "down":
            h = F.conv2d(h, sd[p + ".conv.weight"], sd[p + ".conv.bias"], stride=2, padding=1)
        elif l[0] == "up":
            h = F.interpolate(h, scale_factor=2, mode="nearest")
            h = F.conv2d(h, sd[p + ".conv.weight"], sd[p + ".conv.bias"], padding=1)
    return h


@torch.no_grad()
def unet_forward(sd, cfg, x, t, y=None, num_classes=None):
    """eps = UNet(x, t, y) in fp32 on the tensors' device (CPU in tests)."""
    mc = cfg["model_channels"]
    e = _unet_time_embedding(t, mc)
    e = F.linear(e, sd["time_embed.1.weight"], sd["time_embed.1.bias"])
    t_emb = F.linear(F.silu(e), sd["time_embed.3.weight"], sd["time_embed.3.bias"])
    if num_classes is not None and y is not None:
        # unet.py:257-258; the table is nn.Embedding(..., padding_idx=0) (unet.py:183): same forward, and under autograd the null
        # row 0 receives no gradient
        y_emb = F.embedding(torch.clamp(y, 0, num_classes), sd["label_embed.weight"], padding_idx=0)
    else:
        y_emb = None
    down, middle, up, _ = unet_block_structure(cfg)
    h = F.conv2d(x, sd["input_conv.weight"], sd["input_conv.bias"], padding=1)
    hs = [h]
    for i, layers in enumerate(down):
        h = _run_entry(sd, f"down_blocks.{i}", layers, h, t_emb, y_emb)
        hs.append(h)
    h = _run_entry(sd, "middle_block", middle, h, t_emb, y_emb)
    for i, layers in enumerate(up):
        h = torch.cat([h, hs.pop()], dim=1)
        h = _run_entry(sd, f"up_blocks.{i}", layers, h, t_emb, y_emb)
    h = F.silu(F.group_norm(h, 8, sd["output.0.weight"], sd["output.0.bias"], eps=1e-5))
    return F.conv2d(h, sd["output.2.weight"], sd["output.2.bias"], padding=1)


def _dit_time_embedding(t, dim=256, max_period=10000):
    # models/dit.py:42-50 -- divisor half, order [cos | sin], t.float()
    half = dim // 2
    freqs = torch.exp(-math.log(max_period) * torch.arange(0, half, dtype=torch.float32) / half).to(t.device)
    args = t[:, None].float() * freqs[None]
    return torch.cat([torch.cos(args), torch.sin(args)], dim=-1)


def _ln(x):
    return F.layer_norm(x, (x.shape[-1],), eps=1e-6)


@torch.no_grad()
def dit_forward(sd, cfg, x, t, y=None, num_classes=None):
    """eps = DiT(x, t, y) in fp32, eval mode (no dropout)."""
    hs = cfg["hidden_size"]
    nh = cfg["num_heads"]
    p = cfg["patch_size"]
    B, C, H, W = x.shape
    hh, ww = H // p, W // p
    tok = F.conv2d(x, sd["x_embedder.proj.weight"], sd["x_embedder.proj.bias"], stride=p)
    tok = tok.flatten(2).transpose(1, 2) + sd["pos_embed"]
    c = F.linear(_dit_time_embedding(t), sd["t_embedder.mlp.0.weight"], sd["t_embedder.mlp.0.bias"])
    c = F.linear(F.silu(c), sd["t_embedder.mlp.2.weight"], sd["t_embedder.mlp.2.bias"])
    if num_classes is not None and y is not None:
        # padding_idx = 0 (models/dit.py:62): same forward, and under autograd the null row receives no gradient
        c = c + F.embedding(torch.clamp(y, 0, num_classes), sd["y_embedder.embedding_table.weight"], padding_idx=0)
    sc = F.silu(c)
    hd = hs // nh
    for i in range(cfg["depth"]):
        b = f"blocks.{i}"
        mod = F.linear(sc, sd[b + ".adaLN_modulation.1.weight"], sd[b + ".adaLN_modulation.1.bias"])
        sh_a, sc_a, g_a, sh_m, sc_m, g_m = mod.chunk(6, dim=-1)
        h = _ln(tok) * (1 + sc_a.unsqueeze(1)) + sh_a.unsqueeze(1)
        qkv = F.linear(h, sd[b + ".attn.in_proj_weight"], sd[b + ".attn.in_proj_bias"])
        q, k, v = qkv.chunk(3, dim=-1)
        q = q.reshape(B, -1, nh, hd).transpose(1, 2)
        k = k.reshape(B, -1, nh, hd).transpose(1, 2)
        v = v.reshape(B, -1, nh, hd).transpose(1, 2)
        a = torch.softmax(torch.matmul(q, k.transpose(-2, -1)) / math.sqrt(hd), dim=-1)
        h = torch.matmul(a, v).transpose(1, 2).reshape(B, -1, hs)
        h = F.linear(h, sd[b + ".attn.out_proj.weight"], sd[b + ".attn.out_proj.bias"])
        tok = tok + g_a.unsqueeze(1) * h
        h = _ln(tok) * (1 + sc_m.unsqueeze(1)) + sh_m.unsqueeze(1)
        h = F.gelu(F.linear(h, sd[b + ".mlp.0.weight"], sd[b + ".mlp.0.bias"]))
        h = F.linear(h, sd[b + ".mlp.3.weight"], sd[b + ".mlp.3.bias"])
        tok = tok + g_m.unsqueeze(1) * h
    mod = F.linear(sc, sd["final_layer.adaLN_modulation.1.weight"], sd["final_layer.adaLN_modulation.1.bias"])
    shift, scale = mod.chunk(2, dim=-1)
    h = _ln(tok) * (1 + scale.unsqueeze(1)) + shift.unsqueeze(1)
    h = F.linear(h, sd["final_layer.linear.weight"], sd["final_layer.linear.bias"])
    h = h.reshape(B, hh, ww, p, p, C)
    h = torch.einsum("nhwpqc->nchpwq", h)
    return h.reshape(B, C, hh * p, ww * p)


@torch.no_grad()
def dim_forward(sd, cfg, x, t, y=None, num_classes=None):
    """eps = DiM(x, t, y) in fp32, eval mode, for the variant the reference runs without mamba_ssm (models/dim.py:103-117:
    nn.MultiheadAttention with 8 heads in place of Mamba).  Follows models/dim.py:119-141 (MambaBlock.forward), :153-164
    (FeedForward.forward), :188-193 (FinalLayer.forward), :306-340 (unpatchify, DiM.forward); the LayerNorms are affine."""
    hs, nh, p = cfg["hidden_size"], 8, cfg["patch_size"]
    B, C, H, W = x.shape
    hh, ww = H // p, W // p
    hd = hs // nh

    def ln(v, name):
        return F.layer_norm(v, (hs,), sd[name + ".weight"], sd[name + ".bias"], eps=1e-6)

    tok = F.conv2d(x, sd["x_embedder.proj.weight"], sd["x_embedder.proj.bias"], stride=p)
    tok = tok.flatten(2).transpose(1, 2) + sd["pos_embed"]
    c = F.linear(_dit_time_embedding(t), sd["t_embedder.mlp.0.weight"], sd["t_embedder.mlp.0.bias"])
    c = F.linear(F.silu(c), sd["t_embedder.mlp.2.weight"], sd["t_embedder.mlp.2.bias"])
    if num_classes is not None and y is not None:
        c = c + F.embedding(torch.clamp(y, 0, num_classes), sd["y_embedder.embedding_table.weight"], padding_idx=0)
    sc = F.silu(c)
    for i in range(cfg["depth"]):
        m, f = f"blocks.{i}.mamba_block", f"blocks.{i}.ff_block"
        shift, scale, gate = F.linear(sc, sd[m + ".adaLN_modulation.1.weight"], sd[m + ".adaLN_modulation.1.bias"]).chunk(3, dim=-1)
        h = ln(tok, m + ".norm") * (1 + scale.unsqueeze(1)) + shift.unsqueeze(1)
        q, k, v = F.linear(h, sd[m + ".mamba.in_proj_weight"], sd[m + ".mamba.in_proj_bias"]).chunk(3, dim=-1)
        q = q.reshape(B, -1, nh, hd).transpose(1, 2)
        k = k.reshape(B, -1, nh, hd).transpose(1, 2)
        v = v.reshape(B, -1, nh, hd).transpose(1, 2)
        a = torch.softmax(torch.matmul(q, k.transpose(-2, -1)) / math.sqrt(hd), dim=-1)
        h = torch.matmul(a, v).transpose(1, 2).reshape(B, -1, hs)
        h = F.linear(h, sd[m + ".mamba.out_proj.weight"], sd[m + ".mamba.out_proj.bias"])
        tok = tok + gate.unsqueeze(1) * h
        shift, scale, gate = F.linear(sc, sd[f + ".adaLN_modulation.1.weight"], sd[f + ".adaLN_modulation.1.bias"]).chunk(3, dim=-1)
        h = ln(tok, f + ".norm") * (1 + scale.unsqueeze(1)) + shift.unsqueeze(1)
        h = F.gelu(F.linear(h, sd[f + ".mlp.0.weight"], sd[f + ".mlp.0.bias"]))
        h = F.linear(h, sd[f + ".mlp.3.weight"], sd[f + ".mlp.3.bias"])
        tok = tok + gate.unsqueeze(1) * h
    shift, scale = F.linear(sc, sd["final_layer.adaLN_modulation.1.weight"], sd["final_layer.adaLN_modulation.1.bias"]).chunk(2, dim=-1)
    h = ln(tok, "final_layer.norm_final") * (1 + scale.unsqueeze(1)) + shift.unsqueeze(1)
    h = F.linear(h, sd["final_layer.linear.weight"], sd["final_layer.linear.bias"])
    h = h.reshape(B, hh, ww, p, p, C)
    h = torch.einsum("nhwpqc->nchpwq", h)
    return h.reshape(B, C, hh * p, ww * p)
