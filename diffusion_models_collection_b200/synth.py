"""Deterministic synthetic weights for the UNet / DiT denoisers.

There is no network on the build or GPU boxes, so every benchmark, smoke run and parity
fixture uses random-init weights.  The reference's own default init is not a good parity
probe (GroupNorm gamma=1/beta=0, zero conv biases in places, DiT output identically zero,
SURVEY.md section 8c traps (i)/(ii)), so this module draws *every* tensor from a seeded CPU generator
with the reference's parameter names and shapes:

* UNet keys follow /root/reference/models/unet.py:163-241 (``time_embed.{1,3}``, ``label_embed``,
  ``input_conv``, ``down_blocks.N.M.*``, ``middle_block.{0,1,2}.*``, ``up_blocks.N.M.*``, ``output.{0,2}``).
* DiT keys follow /root/reference/models/dit.py:196-231.

The same function is used (a) to load the *reference* modules when golden fixtures are generated
(``strict=True`` there proves the key/shape contract), (b) by the CPU oracle and (c) by the CUDA
product path, so all three see bit-identical fp32 weights.
"""

from __future__ import annotations

import math
from collections import OrderedDict

import torch

CIFAR_UNET = dict(
    image_size=(32, 32), in_channels=3, model_channels=128, out_channels=3, num_res_blocks=2,
    attention_resolutions=(16, 8), dropout=0.1, channel_mult=(1, 2, 2, 2), use_attention=True,
)

CIFAR_DIT = dict(
    img_size=(32, 32), patch_size=2, in_channels=3, hidden_size=384, depth=12, num_heads=6,
    mlp_ratio=4.0, dropout=0.1,
)


def _conv(g, sd, name, cout, cin, k, bias_scale=0.05):
    fan_in = cin * k * k
    bound = 1.0 / math.sqrt(fan_in)
    sd[name + ".weight"] = (torch.rand(cout, cin, k, k, generator=g) * 2 - 1) * bound
    sd[name + ".bias"] = torch.randn(cout, generator=g) * bias_scale


def _linear(g, sd, name, cout, cin, bias=True, bias_scale=0.05):
    bound = 1.0 / math.sqrt(cin)
    sd[name + ".weight"] = (torch.rand(cout, cin, generator=g) * 2 - 1) * bound
    if bias:
        sd[name + ".bias"] = torch.randn(cout, generator=g) * bias_scale


def _gn(g, sd, name, c):
    sd[name + ".weight"] = 1.0 + 0.2 * torch.randn(c, generator=g)
    sd[name + ".bias"] = 0.1 * torch.randn(c, generator=g)


def unet_block_structure(cfg):
    """Replays the constructor loops of /root/reference/models/unet.py:188-234.

    Returns (down, middle, up, out_ch): lists of entries; every entry is a list of layer tuples
    ``('res', cin, cout)``, ``('attn', ch)``, ``('down', ch)``, ``('up', ch)``.
    """
    mc = cfg["model_channels"]
    mult = tuple(cfg["channel_mult"])
    nrb = cfg["num_res_blocks"]
    attn_res = tuple(cfg["attention_resolutions"])
    use_attn = cfg.get("use_attention", True)
    res = list(cfg["image_size"]) if not isinstance(cfg["image_size"], int) else [cfg["image_size"]] * 2
    ch = mc
    chans = [ch]
    down = []
    for level, m in enumerate(mult):
        out_ch = mc * m
        for _ in range(nrb):
            layers = [("res", ch, out_ch)]
            ch = out_ch
            if use_attn and (res[0] in attn_res or res[1] in attn_res):
                layers.append(("attn", ch))
            down.append(layers)
            chans.append(ch)
        if level != len(mult) - 1:
            down.append([("down", ch)])
            chans.append(ch)
            res = [res[0] // 2, res[1] // 2]
    middle = [("res", ch, ch), ("attn", ch) if use_attn else ("identity",), ("res", ch, ch)]
    up = []
    for level, m in enumerate(reversed(mult)):
        for i in range(nrb + 1):
            ich = chans.pop()
            layers = [("res", ch + ich, mc * m)]
            ch = mc * m
            if use_attn and (res[0] in attn_res or res[1] in attn_res):
                layers.append(("attn", ch))
            if level != len(mult) - 1 and i == nrb:
                layers.append(("up", ch))
                res = [res[0] * 2, res[1] * 2]
            up.append(layers)
    return down, middle, up, ch


def make_unet_state_dict(cfg=None, num_classes=None, seed=0, null_row_zero=True):
    """fp32 CPU state_dict with the reference UNet's keys (357 tensors cond / 334 uncond at CIFAR config)."""
    cfg = dict(CIFAR_UNET if cfg is None else cfg)
    g = torch.Generator(device="cpu")
    g.manual_seed(int(seed))
    sd = OrderedDict()
    mc = cfg["model_channels"]
    temb = mc * 4
    _linear(g, sd, "time_embed.1", temb, mc)
    _linear(g, sd, "time_embed.3", temb, temb)
    if num_classes is not None:
        w = torch.randn(num_classes + 1, temb, generator=g)
        if null_row_zero:
            w[0].zero_()  # padding_idx=0, /root/reference/models/unet.py:183
        sd["label_embed.weight"] = w
    _conv(g, sd, "input_conv", mc, cfg["in_channels"], 3)

    def res(prefix, cin, cout):
        _gn(g, sd, prefix + ".conv1.0", cin)
        _conv(g, sd, prefix + ".conv1.2", cout, cin, 3)
        _linear(g, sd, prefix + ".time_mlp.1", cout, temb)
        if num_classes is not None:
            _linear(g, sd, prefix + ".label_proj.1", cout, temb, bias=False)
        _gn(g, sd, prefix + ".conv2.0", cout)
        _conv(g, sd, prefix + ".conv2.3", cout, cout, 3)
        if cin != cout:
            _conv(g, sd, prefix + ".shortcut", cout, cin, 1)

    def attn(prefix, ch):
        _gn(g, sd, prefix + ".norm", ch)
        _conv(g, sd, prefix + ".qkv", 3 * ch, ch, 1)
        _conv(g, sd, prefix + ".proj", ch, ch, 1)

    def entry(prefix, layers):
        for j, l in enumerate(layers):
            p = f"{prefix}.{j}"
            if l[0] == "res":
                res(p, l[1], l[2])
            elif l[0] == "attn":
                attn(p, l[1])
            elif l[0] in ("down", "up"):
                _conv(g, sd, p + ".conv", l[1], l[1], 3)

    down, middle, up, out_ch = unet_block_structure(cfg)
    for i, layers in enumerate(down):
        entry(f"down_blocks.{i}", layers)
    entry("middle_block", middle)
    for i, layers in enumerate(up):
        entry(f"up_blocks.{i}", layers)
    _gn(g, sd, "output.0", out_ch)
    _conv(g, sd, "output.2", cfg["out_channels"], out_ch, 3)
    return sd


def make_dit_state_dict(cfg=None, num_classes=None, seed=0, null_row_zero=True):
    """fp32 CPU state_dict with the reference DiT's keys; the zero-init tensors are re-randomised
    (SURVEY.md section 8c trap (i): default init makes the output identically zero)."""
    cfg = dict(CIFAR_DIT if cfg is None else cfg)
    g = torch.Generator(device="cpu")
    g.manual_seed(int(seed))
    sd = OrderedDict()
    hs = cfg["hidden_size"]
    p = cfg["patch_size"]
    c = cfg["in_channels"]
    img = cfg["img_size"]
    ih, iw = (img, img) if isinstance(img, int) else img
    ntok = (ih // p) * (iw // p)
    hid = int(hs * cfg["mlp_ratio"])
    sd["pos_embed"] = 0.02 * torch.randn(1, ntok, hs, generator=g)
    _conv(g, sd, "x_embedder.proj", hs, c, p)
    _linear(g, sd, "t_embedder.mlp.0", hs, 256)
    _linear(g, sd, "t_embedder.mlp.2", hs, hs)
    if num_classes is not None:
        w = torch.randn(num_classes + 1, hs, generator=g)
        if null_row_zero:
            w[0].zero_()
        sd["y_embedder.embedding_table.weight"] = w
    for i in range(cfg["depth"]):
        b = f"blocks.{i}"
        bound = 1.0 / math.sqrt(hs)
        sd[b + ".attn.in_proj_weight"] = (torch.rand(3 * hs, hs, generator=g) * 2 - 1) * bound
        sd[b + ".attn.in_proj_bias"] = 0.05 * torch.randn(3 * hs, generator=g)
        _linear(g, sd, b + ".attn.out_proj", hs, hs)
        _linear(g, sd, b + ".mlp.0", hid, hs)
        _linear(g, sd, b + ".mlp.3", hs, hid)
        sd[b + ".adaLN_modulation.1.weight"] = 0.02 * torch.randn(6 * hs, hs, generator=g)
        sd[b + ".adaLN_modulation.1.bias"] = 0.02 * torch.randn(6 * hs, generator=g)
    sd["final_layer.linear.weight"] = 0.05 * torch.randn(p * p * c, hs, generator=g)
    sd["final_layer.linear.bias"] = 0.02 * torch.randn(p * p * c, generator=g)
    sd["final_layer.adaLN_modulation.1.weight"] = 0.02 * torch.randn(2 * hs, hs, generator=g)
    sd["final_layer.adaLN_modulation.1.bias"] = 0.02 * torch.randn(2 * hs, generator=g)
    return sd


CIFAR_DIM = dict(img_size=(32, 32), patch_size=2, in_channels=3, hidden_size=512, depth=4, state_size=16, mlp_ratio=4.0, dropout=0.1)


def make_dim_state_dict(cfg=None, num_classes=None, seed=0, null_row_zero=True):
    """fp32 CPU state_dict with the keys of the reference DiM as it is built WITHOUT mamba_ssm (models/dim.py:103-117: the
    nn.MultiheadAttention variant); zero-init tensors re-randomised, LayerNorm affines off their trivial 1 / 0."""
    cfg = dict(CIFAR_DIM if cfg is None else cfg)
    g = torch.Generator(device="cpu")
    g.manual_seed(int(seed))
    sd = OrderedDict()
    hs, p, c = cfg["hidden_size"], cfg["patch_size"], cfg["in_channels"]
    img = cfg["img_size"]
    ih, iw = (img, img) if isinstance(img, int) else img
    hid = int(hs * cfg["mlp_ratio"])
    sd["pos_embed"] = 0.02 * torch.randn(1, (ih // p) * (iw // p), hs, generator=g)
    _conv(g, sd, "x_embedder.proj", hs, c, p)
    _linear(g, sd, "t_embedder.mlp.0", hs, 256)
    _linear(g, sd, "t_embedder.mlp.2", hs, hs)
    if num_classes is not None:
        w = torch.randn(num_classes + 1, hs, generator=g)
        if null_row_zero:
            w[0].zero_()
        sd["y_embedder.embedding_table.weight"] = w

    def norm(name):
        sd[name + ".weight"] = 1.0 + 0.2 * torch.randn(hs, generator=g)
        sd[name + ".bias"] = 0.1 * torch.randn(hs, generator=g)

    for i in range(cfg["depth"]):
        m, f = f"blocks.{i}.mamba_block", f"blocks.{i}.ff_block"
        norm(m + ".norm")
        sd[m + ".mamba.in_proj_weight"] = (torch.rand(3 * hs, hs, generator=g) * 2 - 1) / math.sqrt(hs)
        sd[m + ".mamba.in_proj_bias"] = 0.05 * torch.randn(3 * hs, generator=g)
        _linear(g, sd, m + ".mamba.out_proj", hs, hs)
        sd[m + ".adaLN_modulation.1.weight"] = 0.02 * torch.randn(3 * hs, hs, generator=g)
        sd[m + ".adaLN_modulation.1.bias"] = 0.02 * torch.randn(3 * hs, generator=g)
        norm(f + ".norm")
        _linear(g, sd, f + ".mlp.0", hid, hs)
        _linear(g, sd, f + ".mlp.3", hs, hid)
        sd[f + ".adaLN_modulation.1.weight"] = 0.02 * torch.randn(3 * hs, hs, generator=g)
        sd[f + ".adaLN_modulation.1.bias"] = 0.02 * torch.randn(3 * hs, generator=g)
    norm("final_layer.norm_final")
    sd["final_layer.linear.weight"] = 0.05 * torch.randn(p * p * c, hs, generator=g)
    sd["final_layer.linear.bias"] = 0.02 * torch.randn(p * p * c, generator=g)
    sd["final_layer.adaLN_modulation.1.weight"] = 0.02 * torch.randn(2 * hs, hs, generator=g)
    sd["final_layer.adaLN_modulation.1.bias"] = 0.02 * torch.randn(2 * hs, generator=g)
    return sd
